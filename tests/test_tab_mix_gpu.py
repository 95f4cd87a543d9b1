"""`-m gpu`: the fused Morph-FC mixing half of a TAB (csrc/tab_mix.cu, function.py:743-805 + 1212-1215) against
  (1) the oracle's restatement of Enhanced_MorphFCs_decay (oracle/vmg_oracle.py:morphfc_decay) in strict fp32 on the
      same bf16-rounded inputs and weights, executed on the device;
  (2) the unfused kernel sequence of vmg_b200/model.py (layernorm, morph permutes, linears, channel sums, mix, gated
      projection), which the fused path replaces.
Shapes: whole tiles, ragged borders (H % 8 != 0, W % 16 != 0: zero padding after the norm, sums over valid pixels only),
B = 2 (per-sample means with time-major images), the headline frame size."""
import pytest
import torch

from oracle import vmg_oracle
from vmg_b200 import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * 2 - 1) * scale


def make(C, seed):
    sd = {}
    for i, name in enumerate(('mlp_h.0', 'mlp_w.0', 'mlp_c.0', 'proj')):
        sd['p.' + name + '.weight'] = rnd(C, C, seed=seed + i, scale=3.0 * C ** -0.5).bfloat16().float().to(DEV)
        sd['p.' + name + '.bias'] = rnd(C, seed=seed + 10 + i, scale=0.3).to(DEV)
    sd['p.gamma_h'] = torch.ones(C, C, device=DEV)
    sd['p.gamma_w'] = torch.ones(C, C, device=DEV)
    sd['p.reweight.fc1.weight'] = rnd(C // 4, C, seed=seed + 20, scale=2.0).to(DEV)
    sd['p.reweight.fc1.bias'] = rnd(C // 4, seed=seed + 21, scale=0.5).to(DEV)
    sd['p.reweight.fc2.weight'] = rnd(3 * C, C // 4, seed=seed + 22, scale=1.0).to(DEV)
    sd['p.reweight.fc2.bias'] = rnd(3 * C, seed=seed + 23, scale=0.5).to(DEV)
    sd['n.weight'] = (1 + rnd(C, seed=seed + 30, scale=0.3)).to(DEV)
    sd['n.bias'] = rnd(C, seed=seed + 31, scale=0.3).to(DEV)
    return sd


CASES = [
    # T, B, H, W, C
    (3, 1, 24, 32, 112),
    (2, 2, 20, 40, 112),       # ragged: 2.5 tile rows, 2.5 tile columns; two batch samples (time-major images)
    (1, 1, 8, 16, 112),        # a single tile
    (2, 1, 184, 320, 112),     # the headline frame size: 920 tiles on 148 persistent blocks
    (2, 1, 36, 24, 64),        # C = 64: one K block
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: 'x'.join(map(str, c)))
def test_tab_mix_against_oracle_and_unfused(case):
    T, B, H, W, C = case
    sd = make(C, seed=C + H)
    m_scaling = 0.7
    x = rnd(T * B, H, W, C, seed=1, scale=2.0).bfloat16().to(DEV)           # time-major: image i = t*B + b
    pcs = tuple(ops.pack_conv(sd[f'p.{n}.weight'], sd[f'p.{n}.bias'], 1, tc=True) for n in ('mlp_h.0', 'mlp_w.0', 'mlp_c.0', 'proj'))
    assert ops.tab_mix_supported(x, C, 8, 8)
    # ---- oracle (fp32, reference layout [B,T,H,W,C])
    xf = x.float().view(T, B, H, W, C).permute(1, 0, 2, 3, 4).contiguous()
    xn = vmg_oracle.layer_norm(sd, 'n', xf)
    sdo = {k: v.clone() for k, v in sd.items()}
    ref = xf + vmg_oracle.morphfc_decay(sdo, 'p', xn, 8, 8, 'vanilla') * m_scaling
    ref = ref.permute(1, 0, 2, 3, 4).reshape(T * B, H, W, C)
    # reference mean of h + w + c (recomputed the same way)
    # ---- fused
    n0 = ops.launch_count()
    acc = ops.tab_mix(x, sd['n.weight'], sd['n.bias'], 1e-5, pcs, 0, B)
    mean = ops.fixsum_finalize(acc, 1.0 / (T * H * W))
    a = ops.linear(mean, ops.pack_conv(sd['p.reweight.fc1.weight'], sd['p.reweight.fc1.bias']), act=ops.ACT_GELU, use_tc=False)
    a = ops.softmax3(ops.linear(a, ops.pack_conv(sd['p.reweight.fc2.weight'], sd['p.reweight.fc2.bias']), use_tc=False))
    out = ops.tab_mix(x, sd['n.weight'], sd['n.bias'], 1e-5, pcs, 1, B, a=a, m_scaling=m_scaling)
    torch.cuda.synchronize()
    # ---- unfused kernel sequence (vmg_b200/model.py:_tab before the fusion)
    xnk = ops.layernorm(x, sd['n.weight'], sd['n.bias'])
    br = []
    for axis, pc in ((0, pcs[0]), (1, pcs[1])):
        A = ops.morph_gather(xnk, 8, axis)
        Y = torch.empty_strided(A.shape, A.stride(), device=DEV, dtype=A.dtype)
        ops.linear(A, pc, act=ops.ACT_RELU, out_scale=1.0 / C, out=Y)
        br.append(ops.morph_scatter(Y, T * B, H, W, C, 8, axis))
    cb = ops.linear(xnk, pcs[2], act=ops.ACT_RELU, out_scale=1.0 / C)
    mean_u = ops.channel_sum(br[0], br[1], cb, T, B, H * W, 1.0 / (T * H * W))
    err_mean = (mean - mean_u).abs().max().item() / mean_u.abs().max().item()
    assert err_mean <= 5e-3, f'phase-0 mean vs unfused channel_sum: relative {err_mean}'
    au = ops.linear(mean_u, ops.pack_conv(sd['p.reweight.fc1.weight'], sd['p.reweight.fc1.bias']), act=ops.ACT_GELU, use_tc=False)
    au = ops.softmax3(ops.linear(au, ops.pack_conv(sd['p.reweight.fc2.weight'], sd['p.reweight.fc2.bias']), use_tc=False))
    mix = ops.mix3(br[0], br[1], cb, au, T, B, H * W)
    out_u = ops.linear(mix, pcs[3], epi=ops.EPI_GATE, res=x, aux=xnk, res_scale=m_scaling)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs()
    err_u = (out_u.float() - ref).abs()
    bound = 0.03 + ref.abs() * 2.0 ** -7
    assert torch.isfinite(out.float()).all()
    assert (err <= bound).all(), f'fused vs oracle: max-abs {err.max().item():.4f} (unfused path: {err_u.max().item():.4f})'
    assert err.mean().item() <= 1.5 * err_u.mean().item() + 1e-4, (err.mean().item(), err_u.mean().item())
    assert ops.launch_count() - n0 >= 2


def test_tab_mix_is_run_to_run_identical():
    T, B, H, W, C = 2, 2, 40, 48, 112
    sd = make(C, seed=7)
    x = rnd(T * B, H, W, C, seed=2, scale=2.0).bfloat16().to(DEV)
    pcs = tuple(ops.pack_conv(sd[f'p.{n}.weight'], sd[f'p.{n}.bias'], 1, tc=True) for n in ('mlp_h.0', 'mlp_w.0', 'mlp_c.0', 'proj'))
    accs = [ops.tab_mix(x, sd['n.weight'], sd['n.bias'], 1e-5, pcs, 0, B).clone() for _ in range(4)]
    assert all(torch.equal(accs[0], t) for t in accs[1:]), 'fixed-point channel sums must not depend on the order of the atomics'


def test_tab_mix_unsupported_shapes():
    x = torch.zeros(1, 16, 16, 144, device=DEV, dtype=torch.bfloat16)
    assert not ops.tab_mix_supported(x, 144, 8, 8)
    assert not ops.tab_mix_supported(x[..., :112].contiguous(), 112, 16, 16)
    L = ops.lib()
    assert L.vmgb_tab_mix(None, None, None, 1e-5, None, None, None, None, 112, None, None, None, None, 0, None, None, None, 1.0, 1, 1, 8, 8, 112, 8, None) == -1
