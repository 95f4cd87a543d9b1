"""Sharding logic on CPU, including a world_size-2 gloo run (no data-path collective in the forward)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vmg_b200 import shard


def test_assign_partitions_units():
    for n, w in ((72, 8), (18, 4), (5, 8), (0, 2)):
        got = sorted(u for r in range(w) for u in shard.assign(n, r, w))
        assert got == list(range(n))
        sizes = [len(shard.assign(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_tile_starts_match_reference_enumeration():
    # tools/Tester.py:107-143 with wins=[128,128], overlap 20 on 180x320, and T=100 windows of 50 overlap 25
    assert shard.tile_starts(180, 128, 20) == [0, 52]
    assert shard.tile_starts(320, 128, 20) == [0, 108, 192]
    assert shard.tile_starts(100, 50, 25) == [0, 25, 50]
    assert shard.tile_starts(540, 128, 20) == [0, 108, 216, 324, 412]
    assert shard.tile_starts(64, 128, 20) == [0]


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = {u: float(u) * 2.0 for u in shard.assign(7, rank, world)}     # stand-in for per-unit results
    allr = shard.gather_objects(mine)
    merged = {}
    for d in allr:
        merged.update(d)
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                               # the bench's max-over-ranks reduction
    q.put((rank, merged, t.item()))
    dist.destroy_process_group()


def test_two_rank_gloo_gather():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, merged, mx in res:
        assert merged == {u: float(u) * 2.0 for u in range(7)}
        assert mx == 4.0
