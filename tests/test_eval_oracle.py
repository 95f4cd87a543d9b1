"""CPU suite for the evaluation path: pins oracle/eval_oracle.py against the golden vectors minted
from the UNMODIFIED reference (`tools.Tester.Tester`, `utils.metrics`) and -- in the build container -- against the
reference itself; then checks the HOST logic of vmg_b200/tester.py (tile / window enumeration, mask extents, loop
order) against the same goldens with the kernels replaced by their torch doubles (tests/emu_ops.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import eval_oracle as eo
from oracle import mint_eval_goldens as mint
from oracle import ref_loader
from tests import common
from tests import emu_ops as emu

GOLD = torch.load(os.path.join(common.GOLDEN, 'eval', 'evalpath.pt'))


def _cfg(nf, ovf, wins, ov, name='REDS4'):
    return {'scale': 4, 'dataset': {'name': name, 'num_frames': nf, 'overlapped_mode': ovf, 'wins': wins,
                                    'overlapped_spatial_length': ov}}


@pytest.mark.parametrize('case', GOLD['cases'], ids=lambda c: 'x'.join(map(str, c['args'][:5])))
def test_oracle_tester_route_matches_reference_golden(case):
    T, H, W, nf, ovf, wins, ov = case['args']
    x = eo.frames_to_input(case['frames'].numpy())
    y = eo.test_clips(mint.FakeModel(), x, nf, ovf, wins, ov, 4)
    assert np.array_equal(eo.quantize(y), case['u8'].numpy())


@pytest.mark.parametrize('case', GOLD['cases'], ids=lambda c: 'x'.join(map(str, c['args'][:5])))
def test_product_tester_host_logic_matches_reference_golden(case, monkeypatch):
    import vmg_b200.tester as vt
    monkeypatch.setattr(vt, '_ops', emu)
    T, H, W, nf, ovf, wins, ov = case['args']

    class M(mint.FakeModel):
        def parameters(self):
            return iter([torch.zeros(1)])
    t = vt.Tester(_cfg(nf, ovf, wins, ov), model=M())
    t._to_device_u8 = lambda f: torch.from_numpy(np.ascontiguousarray(f)) if isinstance(f, np.ndarray) else f
    out = t.evaluate_device(case['frames'].numpy())
    assert torch.equal(out, case['u8'])
    with pytest.raises(RuntimeError):
        t.test_clips(torch.zeros(1, nf - 1, 3, H, W))


def test_oracle_metrics_and_augment_match_reference_golden():
    m = GOLD['metrics']
    a, b = m['a'].numpy(), m['b'].numpy()
    for i in range(3):
        assert eo.calculate_psnr(a[i], b[i]) == m['psnr'][i]                      # integer SSE: bit-exact
        assert eo.structural_similarity(a[i], b[i]) == pytest.approx(m['ssim'][i], abs=1e-12)
    assert eo.calculate_psnr(a[0], b[0], border=2) == m['psnr_border2']
    assert eo.calculate_psnr(a[0], a[0]) == float('inf')
    x = GOLD['augment']['x']
    for h, v, r, ref in GOLD['augment']['modes']:
        assert torch.equal(eo.augment(x, h, v, r), ref)
        assert torch.equal(emu.augment(x, h, v, r), ref)
    # published known answers of the BT.601 studio-swing luma
    assert eo.rgb2ycbcr_y(np.zeros((1, 1, 3), np.uint8))[0, 0] == 16.0
    assert eo.rgb2ycbcr_y(np.full((1, 1, 3), 255, np.uint8))[0, 0] == pytest.approx(235.0, abs=1e-12)
    k = eo.gaussian_kernel()
    assert k.sum() == pytest.approx(1.0, abs=1e-15) and k[5] == k.max()
    from vmg_b200 import metrics as vm
    assert np.allclose(vm.gaussian_window(), k, rtol=0, atol=1e-15)


def test_overlap_modes_and_python_floor_division():
    assert [eo.overlapped_num_frames(50, m) for m in ('small', 'mid', 'large', 7)] == [2, 25, 49, 7]
    import vmg_b200.tester as vt
    t = vt.Tester.__new__(vt.Tester)
    t.overlapped_num_frames, t.overlapped_spatial_length = 25, 20
    # -25//2 = -13: the earlier window drops its last 13 frames, the later one its first 12 (tools/Tester.py:166-170)
    assert t.window_masks(0, [0, 25, 50], 50) == (0, 13) and t.window_masks(25, [0, 25, 50], 50) == (12, 13)
    assert t.window_masks(50, [0, 25, 50], 50) == (12, 0)
    assert t.tile_masks(0, 108, [0, 52], [0, 108, 192], 512, 512) == ((0, 10), (10, 10))
    # REDS4 eval route (configs/VMG-few_levels-eval-REDS4.yml): 3 windows x 2 x 3 tiles per 100 x 180 x 320 clip
    t.test_spatial, t.test_num_frames = [128, 128], 50
    assert t.tile_lists(180, 320) == ([0, 52], [0, 108, 192]) and t.window_list(100) == [0, 25, 50]
    assert t.tile_lists(540, 960) == ([0, 108, 216, 324, 412], [0, 108, 216, 324, 432, 540, 648, 756, 832])


@pytest.mark.skipif(not ref_loader.available(), reason='/root/reference not present (GPU box)')
def test_oracle_matches_live_reference():
    Tester, M = mint.load_reference_tester()
    rng = np.random.RandomState(11)
    for (T, H, W, nf, ovf, wins, ov) in [(8, 26, 34, 4, 3, [12, 16], 4), (5, 16, 20, 5, 2, None, None)]:
        frames = rng.randint(0, 256, size=(T, H, W, 3)).astype(np.uint8)
        ref = mint.bare_tester(Tester, nf, ovf, wins, ov).evaluate(frames)
        mine = eo.quantize(eo.test_clips(mint.FakeModel(), eo.frames_to_input(frames), nf, ovf, wins, ov, 4))
        assert np.array_equal(ref, mine)
    a = rng.randint(0, 256, size=(33, 47, 3)).astype(np.uint8)
    b = rng.randint(0, 256, size=(33, 47, 3)).astype(np.uint8)
    assert eo.calculate_psnr(a, b) == M.calculate_psnr(a, b)
    assert eo.structural_similarity(a, b) == pytest.approx(float(M.structural_similarity(a, b)), abs=1e-12)
    ya, yb = eo.rgb2ycbcr_y(a), eo.rgb2ycbcr_y(b)
    assert eo.structural_similarity(ya, yb) == pytest.approx(float(M.structural_similarity(ya, yb)), abs=1e-12)
