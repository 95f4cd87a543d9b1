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
    from . import ops
    n = target_calls - done_calls
    if n < 0:
        raise ValueError('cannot undo decay; reload the state_dict')
    for mod in model.modules():
        if hasattr(mod, 'gamma_h') and hasattr(mod, 'mlp_h'):
            for lin, g in ((mod.mlp_h[0], mod.gamma_h), (mod.mlp_w[0], mod.gamma_w)):
                for _ in range(n):
                    ops.decay_weights_(lin.weight.data, g)
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
