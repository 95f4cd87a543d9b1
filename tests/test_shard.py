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


# ---- Tester-route sharding (BASELINE configs[2]): units over ranks, tiles shipped to the clip owner, reference blend order
def test_unit_enumeration_and_block_assignment():
    import vmg_b200.tester as vt
    t = vt.Tester.__new__(vt.Tester)
    t.test_spatial, t.overlapped_spatial_length, t.test_num_frames, t.overlapped_num_frames = [128, 128], 20, 50, 25
    units = shard.enumerate_units(t, 4, 100, 180, 320)
    assert len(units) == 72 and units[0] == (0, 0, 0, 0) and units[17] == (0, 50, 52, 192) and units[18] == (1, 0, 0, 0)
    for world in (1, 2, 4, 8, 3, 5):
        blocks = [shard.assign_blocks(72, r, world) for r in range(world)]
        assert sorted(u for b in blocks for u in b) == list(range(72))
        assert all(b == list(range(b[0], b[-1] + 1)) for b in blocks) and max(map(len, blocks)) - min(map(len, blocks)) <= 1
    t.overlapped_spatial_length = None
    assert shard.enumerate_units(t, 2, 100, 180, 320) == [(c, w, None, None) for c in range(2) for w in (0, 25, 50)]


def _tester_worker(rank, world, port, q):
    import numpy as np
    from oracle import mint_eval_goldens as mint
    from tests import common, emu_ops
    import vmg_b200.tester as vt
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    shard._ops = emu_ops
    vt._ops = emu_ops
    gold = torch.load(os.path.join(common.GOLDEN, 'eval', 'evalpath.pt'))['cases'][0]
    T, H, W, nf, ovf, wins, ov = gold['args']

    class UnitModel(mint.FakeModel):
        """Output depends on the GLOBAL unit index, which `decay_to` restores on every rank (stand-in for the stateful
        weight decay of the real model: unit n must see the state after n forwards of a sequential run)."""
        def parameters(self):
            return iter([torch.zeros(1)])

        def modules(self):
            return iter([self])

    m = UnitModel()
    t = vt.Tester({'scale': 4, 'dataset': {'name': 'REDS4', 'num_frames': nf, 'overlapped_mode': ovf, 'wins': wins,
                                           'overlapped_spatial_length': ov}}, model=m)
    real_decay = shard.decay_to

    def fake_decay(model, target, done):
        model.calls = target
        return target
    shard.decay_to = fake_decay
    x = emu_ops.frames_u8_to_f32(gold['frames'])
    out = shard.run_tester_route_sharded(t, [x, x.flip(1).contiguous()], rank, world, exact_decay=True)
    shard.decay_to = real_decay
    q.put((rank, {c: v.numpy() for c, v in out.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_tester_route_matches_reference_golden():
    """2 clips x 3 windows x 6 tiles over 2 ranks (and over 3: a clip spanning two ranks ships tiles to its owner):
    the blended, quantised frames must equal what the unmodified reference Tester produced sequentially."""
    import numpy as np
    from tests import common
    gold = torch.load(os.path.join(common.GOLDEN, 'eval', 'evalpath.pt'))['cases'][0]
    ctx = mp.get_context('spawn')
    for world in (2, 3):
        q = ctx.Queue()
        port = 31500 + (os.getpid() + world) % 2000
        procs = [ctx.Process(target=_tester_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = dict()
        for _ in procs:
            rank, out = q.get(timeout=90)
            for c, v in out.items():
                assert c not in res
                res[c] = v
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert sorted(res) == [0, 1]
        # clip 0 is the golden clip: its 18 units are the first 18 forwards of the sequential reference run
        assert np.array_equal(res[0], gold['u8'].numpy())
