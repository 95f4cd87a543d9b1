"""Work-unit sharding across GPUs (one process per GPU, torch.distributed).

The VMG forward has no exchange step: clips -- and, under the reference's eval semantics
(tools/Tester.py:107-177), temporal windows and spatial tiles of one clip -- are independent forwards
(every reduction in the model is per sample or finer).  So units are partitioned statically
and the only communication is the optional gather of results.

The reference's forward is STATEFUL (in-place Morph-FC weight decay, function.py:765-781): a sequential
reference run uses weights W0*gamma^(n+1) for its n-th forward.  `decay_index` reproduces that for a
sharded run: before unit `n` a replica fast-forwards its weights to exactly `n` decays.
"""
import torch
import torch.distributed as dist

from . import ops as _ops


def assign(n_units, rank, world):
    """Static round-robin: unit u -> rank u % world (keeps per-rank counts within 1 of each other)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError('bad rank/world')
    return list(range(rank, n_units, world))


def tile_starts(length, win, overlap):
    """Start offsets of the reference's overlapped tiling (tools/Tester.py:107-143):
    range(0, L-win, win-ov) + [max(0, L-win)]."""
    if length <= win:
        return [0]
    return list(range(0, length - win, win - overlap)) + [max(0, length - win)]


def gather_objects(local, group=None):
    """All ranks receive the list of every rank's `local` picklable object (gloo or nccl)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [local]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local, group=group)
    return out


def decay_to(model, target_calls, done_calls):
    """Advance the in-place weight decay of every TAB from `done_calls` to `target_calls` forwards
    (W <- W * gamma^(target-done)) so that unit `target_calls` sees the weights the sequential
    reference would have used.  Pure parameter bookkeeping; uses the same elementwise kernel."""
    n = target_calls - done_calls
    if n < 0:
        raise ValueError('cannot undo decay; reload the state_dict')
    for mod in model.modules():
        if hasattr(mod, 'gamma_h') and hasattr(mod, 'mlp_h'):
            for lin, g in ((mod.mlp_h[0], mod.gamma_h), (mod.mlp_w[0], mod.gamma_w)):
                for _ in range(n):
                    _ops.decay_weights_(lin.weight.data, g)
    return target_calls


@torch.no_grad()
def run_sharded(model, units, rank, world, exact_decay=False):
    """Runs `model` on this rank's share of `units` (list of [B,T,3,H,W] tensors); returns
    {unit_index: output}.  With exact_decay the result of unit n equals the n-th forward of one
    sequential reference run."""
    out = {}
    done = 0
    for u in assign(len(units), rank, world):
        if exact_decay:
            done = decay_to(model, u, done)
        x = units[u]
        model.num_out_frames = x.shape[1]
        out[u] = model(x.to(next(model.parameters()).device))
        done += 1
    return out


# ---------------------------------------------------------------------------------------------
# Tester-route sharding (BASELINE configs[2] / [4]): units = (clip, temporal window, spatial tile)
# ---------------------------------------------------------------------------------------------
def enumerate_units(tester, n_clips, T, H, W):
    """Units of `tools/Tester.py:test_clips` -> `test_image` for n_clips clips of [T,3,H,W], in the reference's loop order
    (clip, window, tile row, tile column): list of (clip, t_idx, h_idx, w_idx).  REDS4 eval route: 4 x 3 x 2 x 3 = 72."""
    t_list = tester.window_list(T)
    if tester.overlapped_spatial_length is None:
        return [(c, t, None, None) for c in range(n_clips) for t in t_list]
    h_list, w_list = tester.tile_lists(H, W)
    return [(c, t, h, w) for c in range(n_clips) for t in t_list for h in h_list for w in w_list]


def assign_blocks(n_units, rank, world):
    """Contiguous blocks (unit u -> rank u * world // n_units): consecutive units belong to the same clip, so most
    finished tiles are blended where they were produced and only the tiles of a clip that spans two ranks move."""
    return [u for u in range(n_units) if u * world // n_units == rank]


@torch.no_grad()
def run_tester_route_sharded(tester, clips, rank=0, world=1, exact_decay=True, group=None, state=None):
    """clips: list of device fp32 [1,T,3,H,W] tensors (replicated on every rank).  Every rank forwards its block of
    units; the rank that produced the FIRST unit of a clip owns the clip: it receives the other ranks' finished tiles
    over NCCL send/recv (NVLink), blends every window in the reference's order (`vmgb_blend_accumulate`, bit-identical
    sums), and quantises.  Returns {clip: uint8 [T,sH,sW,3] device tensor} for the clips this rank owns.
    exact_decay: unit n runs on the weights the n-th forward of one sequential reference run would see.
    state: optional {'base': first unit index of this call in that sequential run, 'done': forwards (= decays) this
    replica has executed so far}; updated in place so that consecutive calls continue one long reference run."""
    ops = _ops
    n_clips = len(clips)
    _, T, C, H, W = clips[0].shape
    s = tester.scale
    units = enumerate_units(tester, n_clips, T, H, W)
    n = len(units)
    producer = [u * world // n for u in range(n)]
    owner = {}
    for u, (c, *_r) in enumerate(units):
        owner.setdefault(c, producer[u])
    t_list = tester.window_list(T)
    tiled = tester.overlapped_spatial_length is not None
    h_list, w_list = tester.tile_lists(H, W) if tiled else ([None], [None])
    nf = min(tester.test_num_frames, T)
    dev = clips[0].device
    state = state if state is not None else {'base': 0, 'done': 0}
    base, done = state.get('base', 0), state.get('done', 0)
    pending, keep = [], []
    patches = {}
    # pass 1: produce (and ship) this rank's units, post the receives of the units this rank owns but does not produce
    for u, (c, t_idx, h_idx, w_idx) in enumerate(units):
        p, o = producer[u], owner[c]
        if rank == p:
            if exact_decay:
                done = decay_to(tester.model, base + u, done)
            x = clips[c][:, t_idx:t_idx + tester.test_num_frames]
            if tiled:
                x = x[..., h_idx:h_idx + tester.test_spatial[0], w_idx:w_idx + tester.test_spatial[1]]
            y = tester._forward(x)
            done += 1
            if o == rank:
                patches[u] = y
            else:
                pending.append(dist.isend(y, o, group=group))
                keep.append(y)
        elif rank == o:
            ph = (min(tester.test_spatial[0], H) if tiled else H) * s
            pw = (min(tester.test_spatial[1], W) if tiled else W) * s
            buf = torch.empty((1, nf, C, ph, pw), device=dev, dtype=torch.float32)
            pending.append(dist.irecv(buf, p, group=group))
            patches[u] = buf
    state['done'] = done
    state['base'] = base + n
    for req in pending:
        req.wait()
    # pass 2 (owners): blend in the reference's loop order
    out = {}
    u = 0
    for c in range(n_clips):
        mine = owner[c] == rank
        E = cnt = None
        if mine:
            E = torch.zeros((1, T, C, H * s, W * s), device=dev, dtype=torch.float32)
            cnt = torch.zeros((T, H * s, W * s), device=dev, dtype=torch.float32)
        for t_idx in t_list:
            if tiled:
                Es = cs = None
                if mine:
                    Es = torch.zeros((1, nf, C, H * s, W * s), device=dev, dtype=torch.float32)
                    cs = torch.zeros((nf, H * s, W * s), device=dev, dtype=torch.float32)
                for h_idx in h_list:
                    for w_idx in w_list:
                        if mine:
                            y = patches.pop(u)
                            zh, zw = tester.tile_masks(h_idx, w_idx, h_list, w_list, y.shape[-2], y.shape[-1])
                            ops.blend_accumulate_(Es, cs, y, 0, h_idx * s, w_idx * s, (0, 0), zh, zw)
                        u += 1
                win = ops.blend_normalize_(Es, cs) if mine else None
            else:
                win = patches.pop(u) if mine else None
                u += 1
            if mine:
                ops.blend_accumulate_(E, cnt, win, t_idx, 0, 0, tester.window_masks(t_idx, t_list, nf), (0, 0), (0, 0))
        if mine:
            out[c] = ops.blend_finalize_u8(E, cnt)
    del keep
    return out
