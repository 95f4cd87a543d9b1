"""`-m gpu`: the device-side evaluation path (csrc/evalpath.cu through vmg_b200/tester.py, metrics.py, frame_io.py)
against the golden vectors minted from the unmodified reference Tester / metrics and against oracle/eval_oracle.py.
Byte / integer / index work is required to be BIT-EXACT; float64 metrics agree to 1e-10."""
import os

import numpy as np
import pytest
import torch

from oracle import eval_oracle as eo
from oracle import mint_eval_goldens as mint
from oracle import vmg_oracle
from tests import common
from vmg_b200 import frame_io, metrics, ops
from vmg_b200 import tester as vt

pytestmark = pytest.mark.gpu
GOLD = torch.load(os.path.join(common.GOLDEN, 'eval', 'evalpath.pt'))


def _cfg(nf, ovf, wins, ov, name='REDS4', **ds):
    d = {'name': name, 'num_frames': nf, 'overlapped_mode': ovf, 'wins': wins, 'overlapped_spatial_length': ov}
    d.update(ds)
    return {'scale': 4, 'dataset': d}


class FakeDev(mint.FakeModel):
    def parameters(self):
        return iter([torch.zeros(1, device='cuda')])


@pytest.mark.parametrize('case', GOLD['cases'], ids=lambda c: 'x'.join(map(str, c['args'][:5])))
def test_tester_route_bit_exact_vs_reference_golden(case):
    T, H, W, nf, ovf, wins, ov = case['args']
    t = vt.Tester(_cfg(nf, ovf, wins, ov), model=FakeDev())
    n0 = ops.launch_count()
    out = t.evaluate(case['frames'].numpy())
    assert ops.launch_count() > n0
    assert out.dtype == np.uint8 and np.array_equal(out, case['u8'].numpy())


def test_frame_io_bit_exact_and_pinned_roundtrip():
    rng = np.random.RandomState(3)
    frames = rng.randint(0, 256, size=(5, 37, 53, 3)).astype(np.uint8)
    x = ops.frames_u8_to_f32(torch.from_numpy(frames).cuda())
    assert torch.equal(x.cpu(), eo.frames_to_input(frames))
    xb = ops.frames_u8_to_f32(torch.from_numpy(frames).cuda(), reverse_channels=True)
    assert torch.equal(xb.cpu(), eo.frames_to_input(frames[..., ::-1].copy()))
    # quantisation incl. exact .5 ties (half to even), out-of-range values and the clamp
    y = torch.rand(1, 4, 3, 20, 28, generator=torch.Generator().manual_seed(5)) * 1.4 - 0.2
    y[0, 0, 0, 0, :8] = torch.tensor([0.5 / 255, 1.5 / 255, 2.5 / 255, 254.5 / 255, -1.0, 2.0, 0.0, 1.0])
    assert np.array_equal(ops.frames_f32_to_u8(y.cuda()).cpu().numpy(), eo.quantize(y))
    up = frame_io.FrameUploader()
    for _ in range(3):                                                   # double-buffered pinned staging
        xd = up.upload(frames)
        assert torch.equal(xd.cpu(), eo.frames_to_input(frames))
    host = frame_io.download_frames(xd)
    torch.cuda.synchronize()
    assert np.array_equal(host.numpy(), frames)                          # u8 -> /255 -> *255 round trip is the identity
    host_bgr = frame_io.download_frames(xd, bgr=True)
    torch.cuda.synchronize()
    assert np.array_equal(host_bgr.numpy(), frames[..., ::-1])


def test_blend_kernels_vs_oracle_multi_contributor_order():
    # 4 contributors in the corner overlaps: the fp32 sum depends on the order -> must follow the reference loop order
    g = torch.Generator().manual_seed(2)
    x = torch.rand(2, 3, 3, 40, 44, generator=g)
    ref = eo.test_image(mint.FakeModel(), x, [16, 16], 12, 4)
    t = vt.Tester(_cfg(3, 'small', [16, 16], 12), model=FakeDev())
    out = t.test_image(x.cuda())
    assert torch.equal(out.cpu(), ref)
    E = torch.rand(1, 2, 3, 8, 12, generator=g).cuda()
    cnt = torch.full((2, 8, 12), 2.0, device='cuda')
    u8 = ops.blend_finalize_u8(E, cnt)
    assert np.array_equal(u8.cpu().numpy(), eo.quantize(E.cpu() / 2.0))
    assert np.array_equal(ops.blend_finalize_u8(E).cpu().numpy(), eo.quantize(E.cpu()))


def test_metrics_vs_reference_golden_and_oracle():
    m = GOLD['metrics']
    a, b = m['a'].cuda(), m['b'].cuda()
    for i in range(3):
        assert metrics.calculate_psnr(a[i], b[i]) == m['psnr'][i]                          # bit-exact
        assert metrics.structural_similarity(a[i], b[i]) == pytest.approx(m['ssim'][i], abs=1e-10)
    assert metrics.calculate_psnr(a[0], b[0], border=2) == m['psnr_border2']
    assert metrics.calculate_psnr(a[0], a[0]) == float('inf')
    # 720p-sized frames, luma path, ragged SSIM tiles
    rng = np.random.RandomState(9)
    o = rng.randint(0, 256, size=(2, 187, 253, 3)).astype(np.uint8)
    tg = np.clip(o.astype(np.int32) + rng.randint(-20, 21, size=o.shape), 0, 255).astype(np.uint8)
    res = metrics.evaluate_frames(torch.from_numpy(o).cuda(), torch.from_numpy(tg).cuda())
    for i in range(2):
        oy, ty = eo.rgb2ycbcr_y(o[i]), eo.rgb2ycbcr_y(tg[i])
        assert res['psnr'][i] == eo.calculate_psnr(o[i], tg[i])
        assert res['psnr_y'][i] == pytest.approx(eo.calculate_psnr(oy, ty), abs=1e-9)
        assert res['ssim'][i] == pytest.approx(eo.structural_similarity(o[i], tg[i]), abs=1e-10)
        assert res['ssim_y'][i] == pytest.approx(eo.structural_similarity(oy, ty), abs=1e-10)
    y = metrics.rgb2ycbcr_y(torch.from_numpy(o).cuda()).cpu().numpy()
    assert np.abs(y - eo.rgb2ycbcr_y(o)).max() <= 1e-12


def test_augment_and_self_ensemble():
    x = GOLD['augment']['x']
    for h, v, r, ref in GOLD['augment']['modes']:
        assert torch.equal(ops.augment(x.cuda(), h, v, r).cpu(), ref)
    # an equivariant "model" (nearest x4) must come back unchanged from the x8 ensemble
    f = lambda z: z.repeat_interleave(4, -1).repeat_interleave(4, -2)
    xs = torch.rand(1, 2, 3, 10, 14, generator=torch.Generator().manual_seed(1)).cuda()
    y = vt.self_ensemble_x8(f, xs)
    assert (y - f(xs)).abs().max().item() <= 1e-6
    # Tester.evaluate with data_enhance (the reference re-applies the same sequence as the inverse)
    frames = GOLD['cases'][2]['frames'].numpy()
    T, H, W, nf, ovf, wins, ov = GOLD['cases'][2]['args']
    t = vt.Tester(_cfg(nf, ovf, wins, ov, data_enhance=True, use_hflip=True, use_vflip=True, use_rot=False), model=FakeDev())
    xin = eo.augment(eo.frames_to_input(frames), True, True, False)
    ref = eo.quantize(eo.augment(eo.test_clips(mint.FakeModel(), xin, nf, ovf, wins, ov, 4), True, True, False))
    assert np.array_equal(t.evaluate(frames), ref)


def test_tester_route_with_the_real_model_vs_cpu_oracle():
    """Windows x tiles through the real VMG (few_levels, strict fp32): 8 stateful forwards in the reference's order
    (each one decays the Morph-FC weights), blended and quantised on the device, against the same route with the
    CPU oracle as the model."""
    T, H, W, nf = 4, 72, 80, 3
    m = common.build('few', seed=2, precision='fp32', T=nf)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    kw = common.kwargs_of('few')

    class OracleModel:
        def __init__(self):
            self.sd = sd

        def __call__(self, x):
            y, self.sd = vmg_oracle.run(self.sd, kw, x.contiguous())
            return y
    frames = (common.seeded_input((T, H, W, 3), seed=4) * 255).to(torch.uint8).numpy()
    ref_f = eo.test_clips(OracleModel(), eo.frames_to_input(frames), nf, 2, [64, 64], 20, 4)
    t = vt.Tester(_cfg(nf, 'small', [64, 64], 20), model=m.cuda())
    x = ops.frames_u8_to_f32(torch.from_numpy(frames).cuda())
    out_f = t.test_clips(x)
    err = (out_f.cpu() - ref_f).abs().max().item()
    print(f'tester route, real model fp32: max-abs {err:.3e}')
    assert err <= 1e-4
    u8 = ops.frames_f32_to_u8(out_f).cpu().numpy()
    diff = np.abs(u8.astype(np.int32) - eo.quantize(ref_f).astype(np.int32))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3          # quantisation flips only where fp32 differs by ~1e-7 at a .5 tie
