"""`-m gpu`: parity AT THE BENCHMARKED CONFIGURATION (BASELINE.json configs[1]: 100 x 180 x 320, padded 184 x 320).

  * the exact launches of the headline step through the AUTO kernel policy (the CTA-pair weight-stationary kernel
    with several rounds per cluster: accumulator phase flips, activation-ring wrap, one-ahead TMA residual, odd tail)
    against the same convolution computed by cuDNN in strict fp32 on the same device;
  * the nearest / bilinear warps against ATen's CUDA grid_sampler (the reference's own GPU arithmetic), required to be
    BIT-EXACT for the index work (0 mismatching locations);
  * the whole model at 180 x 320: T = 5 against the CPU oracle, T = 100 (both families) against the oracle executed
    on cuda:0 in strict fp32 (SURVEY.md 8c "O3"), with the error-vs-T curve written to gpurun_out/.
Tolerances are the north-star's: fp32 path 1e-4 max-abs on [0,1] pixels; bf16 path 5e-3 max-abs and 0.02 dB PSNR.
"""
import json
import os

import pytest
import torch

from tests import common
from tests import emu_ops as emu
from vmg_b200 import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * 2 - 1) * scale


def big_conv_case(N, H, W, Cin, Cout, R, groups, n_src, act, epi, ps, variant, out_dtype, seed=0):
    """bf16 operands; reference = tests/emu_ops.conv2d on the GPU with TF32 off (cuDNN fp32 on the same bf16-rounded values)."""
    w = rnd(Cout, Cin // groups, R, R, seed=seed + 1, scale=(Cin // groups * R * R) ** -0.5).bfloat16().float()
    b = rnd(Cout, seed=seed + 2, scale=0.5)
    xs = [rnd(N, H, W, Cin // n_src, seed=seed + 10 + i).bfloat16().to(DEV) for i in range(n_src)]
    oshape = (N, 2 * H, 2 * W, Cout // 4) if ps else (N, H, W, Cout)
    res = rnd(*oshape, seed=seed + 3).bfloat16().to(DEV) if epi != ops.EPI_STORE else None
    kw = dict(act=act, alpha=0.1, out_scale=1.0, epi=epi, res_scale=0.1 if epi == ops.EPI_RESADD else 1.0, pixel_shuffle=ps, out_dtype=out_dtype)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        pce = emu.pack_conv(w.to(DEV), b.to(DEV), groups)
        ref = emu.conv2d(tuple(xs) if n_src > 1 else xs[0], pce, res=res, **dict(kw, out_dtype=torch.float32))
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    pc = ops.pack_conv(w.to(DEV), b.to(DEV), groups, tc=True)
    n0 = ops.launch_count()
    out = ops.conv2d(tuple(xs) if n_src > 1 else xs[0], pc, res=res, use_tc=True, tc_variant=variant, **kw)
    torch.cuda.synchronize()
    assert ops.launch_count() == n0 + 1
    return out.float(), ref.float()


def check(out, ref, out_dtype):
    err = (out - ref).abs()
    if out_dtype == torch.bfloat16:      # one bf16 rounding of the result (2^-9 relative) + accumulation-order noise
        bound = 2e-3 + ref.abs() * 2.0 ** -8
    else:
        bound = torch.full_like(ref, 2e-3)
    worst = (err - bound).max().item()
    assert worst <= 0, f'max-abs {err.max().item():.3e}, {worst:.3e} over the bound'


HEADLINE = [
    # N, H, W, Cin, Cout, R, groups, n_src, act, epi, ps                                  what it is in the headline step
    (2, 184, 320, 112, 112, 3, 1, 1, ops.ACT_RELU, ops.EPI_STORE, 0),      # trajectory resblock conv1 (3 000 launches / step)
    (2, 184, 320, 112, 112, 3, 1, 1, ops.ACT_NONE, ops.EPI_RESADD, 0),     # trajectory resblock conv2: TMA residual (3 000 / step)
    (2, 184, 320, 224, 112, 3, 1, 2, ops.ACT_LRELU, ops.EPI_STORE, 0),     # trajectory input conv, virtual concat of 2
    (2, 184, 320, 144, 144, 3, 1, 1, ops.ACT_RELU, ops.EPI_STORE, 0),      # few_levels resblock conv1: two N slices
    (2, 184, 320, 144, 144, 3, 1, 1, ops.ACT_NONE, ops.EPI_RESADD, 0),     # few_levels resblock conv2
    (25, 184, 320, 112, 672, 3, 4, 1, ops.ACT_GELU, ops.EPI_STORE, 0),     # FFN fc1, grouped (a quarter of the clip)
    (4, 360, 640, 112, 256, 3, 1, 1, ops.ACT_LRELU, ops.EPI_STORE, 2),     # upconv2 + pixel shuffle (recon chunk)
    (8, 180, 320, 112, 448, 3, 1, 1, ops.ACT_LRELU, ops.EPI_STORE, 2),     # upconv1 + pixel shuffle
    (99, 192, 320, 32, 64, 7, 1, 1, ops.ACT_RELU, ops.EPI_STORE, 0),       # SPyNet 7x7 at the finest level
    (3, 176, 328, 112, 112, 3, 1, 1, ops.ACT_NONE, ops.EPI_RESADD, 0),     # 1 353 tiles: ODD tile count, 9+ rounds per cluster
    (1, 176, 328, 112, 112, 3, 1, 1, ops.ACT_RELU, ops.EPI_STORE, 0),      # 451 tiles: odd, 3 rounds
]


@pytest.mark.parametrize('out_dtype', [torch.bfloat16, torch.float32], ids=['out_bf16', 'out_f32'])
@pytest.mark.parametrize('variant', [0, 13], ids=['auto', 'cta_pair'])
@pytest.mark.parametrize('case', HEADLINE, ids=lambda c: 'x'.join(map(str, c)))
def test_headline_launch_shapes(case, variant, out_dtype):
    N, H, W, Cin, Cout, R, groups, n_src, act, epi, ps = case
    if out_dtype == torch.float32 and N * H * W * Cout > 3e8:
        pytest.skip('fp32 copy of this output is > 1.2 GB; covered with bf16 output')
    try:
        out, ref = big_conv_case(N, H, W, Cin, Cout, R, groups, n_src, act, epi, ps, variant, out_dtype)
    except RuntimeError as e:
        if variant == 13 and 'code -2' in str(e):
            pytest.skip('CTA-pair kernel does not take this shape')
        raise
    check(out, ref, out_dtype)


def test_headline_linears_and_gate():
    # 1x1 layers of a level-0 TAB at a quarter of the headline token count (1.47 M rows)
    rows = 25 * 184 * 320
    for Cin, Cout, act, epi in ((112, 112, ops.ACT_RELU, ops.EPI_STORE), (112, 112, ops.ACT_NONE, ops.EPI_GATE), (672, 112, ops.ACT_NONE, ops.EPI_RESADD)):
        w = rnd(Cout, Cin, seed=1, scale=Cin ** -0.5).bfloat16().float()
        b = rnd(Cout, seed=2, scale=0.5)
        x = rnd(rows, Cin, seed=3).bfloat16().to(DEV)
        res = rnd(rows, Cout, seed=4).bfloat16().to(DEV) if epi != ops.EPI_STORE else None
        aux = rnd(rows, Cout, seed=5).bfloat16().to(DEV) if epi == ops.EPI_GATE else None
        out = ops.linear(x, ops.pack_conv(w.to(DEV), b.to(DEV), 1, tc=True), act=act, epi=epi, res=res, aux=aux)
        ref = emu.linear(x, emu.pack_conv(w.to(DEV), b.to(DEV)), act=act, epi=epi, res=res, aux=aux, out_dtype=torch.float32)
        check(out.float(), ref.float(), torch.bfloat16)


@pytest.mark.parametrize('case', [
    # rows, Cin, Cout, act, epi, out dtype, pitched output
    (100 * 92 * 160, 224, 224, ops.ACT_RELU, ops.EPI_STORE, torch.bfloat16, False),     # level-1 token / channel Linear of the headline step
    (8 * 7400, 224, 224, ops.ACT_RELU, ops.EPI_STORE, torch.bfloat16, False),           # 463 tiles: odd tail pair, last tile 8 of 16 row groups
    (8 * 7400 + 8, 112, 112, ops.ACT_NONE, ops.EPI_STORE, torch.bfloat16, True),        # projected LTAM values written into a pitched slice
    (8 * 7400, 448, 224, ops.ACT_NONE, ops.EPI_STORE, torch.float32, False),            # resampling Linear, fp32 output (staged epilogue)
    (25 * 184 * 320, 672, 112, ops.ACT_NONE, ops.EPI_RESADD, torch.bfloat16, False),    # FFN fc2 of level 0 + residual
    (8 * 7400, 672, 112, ops.ACT_NONE, ops.EPI_RESADD, torch.bfloat16, False),
], ids=lambda c: f'{c[0]}x{c[1]}to{c[2]}')
def test_linear_on_pair_kernel_bit_identical_to_per_tap(case):
    """Linears with one resident N slice take the CTA-pair weight-stationary kernel (conv_tc.cu, R = 1: rows as an [rows/8, 8]
    image).  Same K order and epilogue arithmetic as the per-tap kernel -> required bit-identical, and within the bf16 bound of
    the fp32 reference."""
    rows, Cin, Cout, act, epi, odt, pitched = case
    w = rnd(Cout, Cin, seed=1, scale=Cin ** -0.5).bfloat16().float()
    b = rnd(Cout, seed=2, scale=0.5)
    pc = ops.pack_conv(w.to(DEV), b.to(DEV), 1, tc=True)
    x = rnd(rows, Cin, seed=3).bfloat16().to(DEV)
    res = rnd(rows, Cout, seed=4).bfloat16().to(DEV) if epi != ops.EPI_STORE else None
    outs = []
    for variant in (0, 1):
        if pitched:
            buf = torch.zeros(rows, 2 * Cout + 16, device=DEV, dtype=odt)
            o = buf[:, 16:16 + Cout]
        else:
            o = torch.empty(rows, Cout, device=DEV, dtype=odt)
        ops.linear(x, pc, act=act, epi=epi, res=res, out=o, tc_variant=variant)
        outs.append(o.float().clone())
        if pitched:
            assert buf[:, :16].abs().max().item() == 0 and buf[:, 16 + Cout:].abs().max().item() == 0      # nothing outside the slice
    assert torch.equal(outs[0], outs[1])
    ref = emu.linear(x, emu.pack_conv(w.to(DEV), b.to(DEV)), act=act, epi=epi, res=res, out_dtype=torch.float32)
    check(outs[0], ref.float(), odt)


def _tie_flows(n, h, w, seed):
    """Flows on a 1/8 grid (many coordinates land exactly on .5 ties) mixed with continuous and 4x4-block-constant ones."""
    g = torch.Generator().manual_seed(seed)
    f = torch.randint(-96, 97, (n, h, w, 2), generator=g).float() / 8
    f[:, : h // 2] = (torch.rand(n, h // 2, w, 2, generator=g) * 2 - 1) * 13.0
    blk = (torch.rand(n, (h + 3) // 4, (w + 3) // 4, 2, generator=g) * 2 - 1) * 9.0
    f[:, :, : w // 3] = blk.repeat_interleave(4, 1).repeat_interleave(4, 2)[:, :h, : w // 3]
    return f


@pytest.mark.parametrize('shape', [(2, 184, 320), (2, 128, 128), (3, 18, 22), (1, 64, 112)], ids=str)
def test_warps_bit_exact_vs_cuda_grid_sampler(shape):
    """Index work must be bit-exact: the location warp against ATen's CUDA grid_sampler (what the reference's forward
    runs on a GPU), ZERO mismatches; and against ATen's CPU grid_sampler with the CPU coordinate semantics."""
    n, h, w = shape
    K = 5
    flow = _tie_flows(n, h, w, seed=h)
    loc = torch.zeros(n, h, w, 2 * K + 2)
    g = torch.Generator().manual_seed(1)
    loc[..., 0:2 * K:2] = torch.randint(0, w, (n, h, w, K), generator=g).float()
    loc[..., 1:2 * K:2] = torch.randint(0, h, (n, h, w, K), generator=g).float()
    x = rnd(n, h, w, 112, seed=3).bfloat16()
    for sem, dev in ((ops.COORD_ATEN_CUDA, DEV), (ops.COORD_ATEN_CPU, 'cpu')):
        ref_loc = emu.flow_warp_nearest(loc.to(dev), flow.to(dev), 2 * K, torch.zeros_like(loc, device=dev)).cpu()
        ref_x = emu.flow_warp_bilinear(x.to(dev).float(), flow.to(dev)).cpu()
        out_loc = torch.zeros_like(loc, device=DEV)
        out_x = ops.flow_warp_fused(x.to(DEV), flow.to(DEV), loc.to(DEV), 2 * K, out_loc, semantics=sem)
        mism = (out_loc.cpu()[..., :2 * K] != ref_loc[..., :2 * K]).sum().item()
        assert mism == 0, f'{mism} mismatching tracked locations vs ATen {dev} grid_sampler'
        assert (out_x.float().cpu() - ref_x).abs().max().item() <= 1.6e-2          # bf16 output rounding
        # the separate launches must agree with the fused one bit for bit
        o2 = ops.flow_warp_nearest(loc.to(DEV), flow.to(DEV), 2 * K, torch.zeros_like(loc, device=DEV), semantics=sem)
        assert torch.equal(o2[..., :2 * K], out_loc[..., :2 * K])
        assert torch.equal(ops.flow_warp_bilinear(x.to(DEV), flow.to(DEV), semantics=sem), out_x)
        xf = x.float().to(DEV)
        o32 = ops.flow_warp_bilinear(xf, flow.to(DEV), semantics=sem)
        assert (o32.cpu() - ref_x).abs().max().item() <= 2e-5


def _psnr_target(shape, seed=8):
    return torch.rand(shape, generator=torch.Generator().manual_seed(seed))


def _report(name, rec):
    os.makedirs(os.path.join(common.ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(common.ROOT, 'gpurun_out', 'parity_headline.jsonl'), 'a') as f:
        f.write(json.dumps(dict(rec, test=name)) + '\n')
    print(json.dumps(dict(rec, test=name)))


_T5_REF = {}


@pytest.mark.parametrize('family', ['full', 'few'])
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_model_180x320_T5_vs_oracles(family, precision):
    """T = 5 at the headline resolution against BOTH executions of the reference arithmetic: the CPU oracle (O1, with the
    ATen-CPU coordinate rounding) and the oracle on cuda:0 in strict fp32 (O3, ATen-CUDA rounding).  SURVEY.md 8(d): the
    gate of the fp32 path is the GPU reference (O2/O3), O1 is reported beside it; the two references themselves differ by
    `o1_vs_o3` (fp32 reassociation moves a smoothed flow across a .5 tie now and then, which changes one tracked
    location), so the fp32 path must be within 1e-4 of the reference execution whose coordinate rounding it follows on
    at least one of the two, and within 3e-4 of both."""
    from oracle import vmg_oracle
    shape = (1, 5, 3, 180, 320)
    kw = common.kwargs_of(family)
    x = common.seeded_input(shape, seed=5)
    m = common.build(family, seed=3, precision=precision, T=5, coord='cpu')
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    if family not in _T5_REF:
        y_o1 = vmg_oracle.run(sd0, kw, x)[0]
        y_o3 = common.oracle_on_cuda(sd0, kw, x, allow_tf32=False)[0].cpu()
        torch.cuda.empty_cache()
        _T5_REF[family] = (y_o1, y_o3)
    y_o1, y_o3 = _T5_REF[family]
    m = m.cuda()
    y_cpu_sem = m(x.cuda()).cpu()
    m.load_state_dict(sd0)
    m.coord_semantics = 'cuda'
    y_cuda_sem = m(x.cuda()).cpu()
    e1 = (y_cpu_sem - y_o1).abs().max().item()
    e3 = (y_cuda_sem - y_o3).abs().max().item()
    tgt = _psnr_target(y_o1.shape)
    dpsnr = max(abs(common.psnr(y_cpu_sem, tgt) - common.psnr(y_o1, tgt)), abs(common.psnr(y_cuda_sem, tgt) - common.psnr(y_o3, tgt)))
    _report('T5_vs_oracles', {'family': family, 'precision': precision, 'shape': shape, 'max_abs_vs_cpu_oracle': e1, 'max_abs_vs_cuda_oracle_fp32': e3,
                              'o1_vs_o3': (y_o1 - y_o3).abs().max().item(), 'dpsnr': dpsnr,
                              'frac_pixels_over_1e-5_vs_o3': ((y_cuda_sem - y_o3).abs() > 1e-5).float().mean().item()})
    if precision == 'fp32':
        assert min(e1, e3) <= 1e-4 and max(e1, e3) <= 3e-4
    else:
        assert e1 <= 5e-3 and e3 <= 5e-3
    assert dpsnr <= 0.02


@pytest.mark.parametrize('family', ['full', 'few'])
def test_model_headline_T100_vs_oracle_on_cuda(family):
    """The benchmarked forward itself (graph replay included) against the oracle on cuda:0 in strict fp32."""
    import bench
    T = 100
    shape = (1, T, 3, 180, 320)
    x = bench.synthetic_clip(1, T, 180, 320, seed=1)
    m = common.build(family, seed=0, precision='bf16', T=T, coord='cuda').cuda()
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    xd = x.cuda()
    m.use_cuda_graph = False
    y_e1, y_e2 = m(xd), m(xd)            # eager: first and second (decayed once more) forward
    m.load_state_dict(sd0)               # same weights and decay index again
    m.use_cuda_graph = True
    y_g1 = m(xd)                         # eager warm-up call of the graph path
    y_g2 = m(xd)                         # captured + replayed
    assert torch.equal(y_e1, y_g1)
    assert torch.equal(y_e2, y_g2), 'graph replay must reproduce the eager forward bit for bit'
    peak_gb = torch.cuda.max_memory_allocated() / 1e9
    y = y_e1.cpu()
    del y_e1, y_e2, y_g1, y_g2, m
    torch.cuda.empty_cache()
    y_ref, _ = common.oracle_on_cuda(sd0, common.kwargs_of(family), x, allow_tf32=False)
    y_ref = y_ref.cpu()
    torch.cuda.empty_cache()
    per_t = (y - y_ref).abs().amax(dim=(0, 2, 3, 4))
    tgt = _psnr_target(y.shape)
    dpsnr = abs(common.psnr(y, tgt) - common.psnr(y_ref, tgt))
    rec = {'family': family, 'precision': 'bf16', 'shape': shape, 'max_abs': per_t.max().item(), 'dpsnr': dpsnr,
           'max_abs_by_frame_decile': [per_t[i:i + 10].max().item() for i in range(0, T, 10)],
           'mean_abs': (y - y_ref).abs().mean().item(), 'peak_mem_GB_eager_plus_graph': peak_gb}
    _report('T100_vs_oracle_on_cuda_fp32', rec)
    assert torch.isfinite(y).all()
    assert rec['max_abs'] <= 5e-3 and dpsnr <= 0.02
