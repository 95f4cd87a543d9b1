"""`-m gpu`: the persistent residual-block chain kernel (csrc/resblock_chain.cu, trajectory.py:16-52,165-221) against
  (1) the same chain issued as one tcgen05 launch per convolution (required BIT-IDENTICAL: same MMA order, same epilogue
      arithmetic -- any cross-tile dependency race shows up as a difference), repeated to catch timing-dependent races;
  (2) cuDNN in strict fp32 on the same bf16-rounded operands (tests/emu_ops.py on the device)."""
import pytest
import torch

from tests import emu_ops as emu
from vmg_b200 import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * 2 - 1) * scale


def make_chain(C, n_blocks, seed=0):
    layers, ws = [], []
    for i in range(2 * n_blocks):
        w = rnd(C, C, 3, 3, seed=seed + 2 * i, scale=1.5 * (C * 9) ** -0.5).bfloat16().float()
        b = rnd(C, seed=seed + 2 * i + 1, scale=0.2)
        ws.append((w, b))
        layers.append(ops.pack_conv(w.to(DEV), b.to(DEV), 1, tc=True))
    return ops.pack_resblock_chain(layers), ws


CASES = [
    # n, H, W, C, blocks
    (2, 184, 320, 112, 15),     # the headline trajectory step: 480 tile pairs on 72 clusters, 30 layers
    (1, 40, 48, 112, 3),        # 9 pairs: one per cluster, every layer waits for its neighbours
    (2, 128, 128, 64, 4),       # one K block
    (1, 64, 112, 128, 2),       # C = 128, W = 112
    (3, 24, 16, 112, 2),        # a single pair per tile row, ragged last tile row
    (4, 180, 320, 112, 2),      # more tile rows than clusters / 3: equal raster ranges instead of row-aligned ownership
    (2, 184, 320, 144, 15),     # the few-levels trajectory step: C = 144 as two output-channel slices of 80 (37 tile clusters x 2 slices)
    (1, 40, 48, 144, 3),        # two slices, one pair per cluster
    (2, 72, 80, 160, 2),        # C = 160: both slices full
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: 'x'.join(map(str, c)))
def test_chain_bit_identical_to_per_layer_launches(case):
    n, H, W, C, nb = case
    pc, ws = make_chain(C, nb, seed=C + nb)
    x = rnd(n, H, W, C, seed=1).bfloat16().to(DEV)
    assert ops.resblock_chain_supported(pc, x)
    # reference 1: one launch per convolution
    y = x.clone()
    for k in range(nb):
        # variant 13 = the CTA-pair weight-stationary kernel (the chain's MMA order: K block major, tap minor); the auto
        # policy would send the small cases to the per-tap kernel, whose accumulation order differs by an ulp
        t = ops.conv2d(y, pc.layers[2 * k], act=ops.ACT_RELU, tc_variant=13)
        y = ops.conv2d(t, pc.layers[2 * k + 1], epi=ops.EPI_RESADD, res=y, res_scale=0.1, tc_variant=13)
    ref = y
    # reference 2: cuDNN fp32 on the device, bf16 rounding between layers as the kernels do
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        yr = x
        for k in range(nb):
            w1, b1 = ws[2 * k]
            w2, b2 = ws[2 * k + 1]
            t = emu.conv2d(yr, emu.pack_conv(w1.to(DEV), b1.to(DEV)), act=ops.ACT_RELU)
            yr = emu.conv2d(t, emu.pack_conv(w2.to(DEV), b2.to(DEV)), epi=ops.EPI_RESADD, res=yr, res_scale=0.1)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    tol = 0.02 * (1 + nb / 4)                    # independent bf16 roundings of 2*nb intermediate tensors
    err_ref = (ref.float() - yr.float()).abs().max().item()
    assert err_ref <= tol, f'per-layer launches vs cuDNN: {err_ref}'
    flags = ops.chain_flags(x)
    scratch = (torch.empty_like(x), torch.empty_like(x))
    n0 = ops.launch_count()
    for rep in range(6):                       # alternating flag arrays, as consecutive trajectory steps do
        out = torch.empty_like(x)
        ops.resblock_chain(x.clone(), pc, 0.1, out, flags[rep & 1], flags[(rep + 1) & 1], scratch)
        torch.cuda.synchronize()
        err = (out.float() - yr.float()).abs().max().item()
        assert err <= tol, f'rep {rep}: chain vs cuDNN {err} (per-layer launches vs cuDNN {err_ref})'
        assert torch.equal(out, ref), f'rep {rep}: max-abs vs per-layer launches {(out.float() - ref.float()).abs().max().item()}'
    assert ops.launch_count() - n0 == 6, 'one kernel per chain call'


def test_chain_unsupported_shapes_take_the_per_layer_path():
    pc, _ = make_chain(192, 1, seed=5)             # C = 192: more than two slices -> not a chain shape
    x = rnd(1, 32, 32, 192, seed=2).bfloat16().to(DEV)
    assert not ops.resblock_chain_supported(pc, x)
    out = torch.empty_like(x)
    y = ops.resblock_chain(x.clone(), pc, 0.1, out, None, None)
    t = ops.conv2d(x, pc.layers[0], act=ops.ACT_RELU)
    ref = ops.conv2d(t, pc.layers[1], epi=ops.EPI_RESADD, res=x, res_scale=0.1)
    assert torch.equal(y, ref)
    L = ops.lib()
    assert L.vmgb_resblock_chain(None, None, None, None, None, 112, None, None, None, 1, 16, 16, 112, 1, 0.1, None) == -1
