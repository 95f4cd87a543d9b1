"""TEST INFRASTRUCTURE (build container only) -- mint golden vectors for the evaluation path from the
UNMODIFIED reference (`tools.Tester.Tester`, `utils.metrics`):   python -m oracle.mint_eval_goldens

The reference `Tester` is instantiated WITHOUT its constructor (which builds a model and reads a checkpoint) and
driven with `fake_model`, a deterministic stand-in whose arithmetic is exact elementwise fp32 (identical on CPU
and GPU), so that the fixture pins tiling, masking, accumulation order, normalisation and quantisation -- not the
network.  Output: tests/golden/eval/evalpath.pt.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, ref_shims            # noqa: E402


class FakeModel:
    """y = nearest_x4(x) * a_k + b_k with call-dependent dyadic (a_k, b_k): every call produces different values, so a
    wrong accumulation order or mask shows up; works on any device."""

    def __init__(self):
        self.calls = 0
        self.num_out_frames = None

    def __call__(self, x):
        k = self.calls
        self.calls += 1
        a = 0.5 + 0.125 * (k % 3)
        b = 0.0625 * (k % 5) - 0.125
        y = x.repeat_interleave(4, -1).repeat_interleave(4, -2)
        return y * a + b


def load_reference_tester():
    ref_shims.install()
    for name in ('fvcore', 'fvcore.nn', 'skimage', 'skimage.metrics', 'skimage.color'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['fvcore.nn'].FlopCountAnalysis = None
    sys.modules['fvcore.nn'].parameter_count_table = None
    if not hasattr(sys.modules['skimage.metrics'], 'peak_signal_noise_ratio'):
        sys.modules['skimage.metrics'].peak_signal_noise_ratio = None
    if ref_loader.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REF_ROOT)
    import tools.Tester as T
    import utils.metrics as M
    return T.Tester, M


def bare_tester(Tester, num_frames, ov_frames, wins, ov, scale=4, name='REDS4'):
    t = Tester.__new__(Tester)
    t.scale, t.test_num_frames, t.overlapped_num_frames = scale, num_frames, ov_frames
    t.test_spatial, t.overlapped_spatial_length = wins, ov
    t.dataset_name, t.data_enhance = name, False
    t.hflip = t.vflip = t.rot = False
    t.device = torch.device('cpu')
    t.model = FakeModel()
    return t


CASES = [
    # (T, H, W, num_frames, ov_frames, wins, ov)
    (9, 30, 38, 4, 2, [16, 16], 6),      # 3 windows x (2 x 3) tiles, ragged last tile, ov//2 = 3
    (7, 20, 24, 5, 3, [16, 16], 5),      # odd overlaps: -5//2 = -3 vs 5//2 = 2, -3//2 = -2 vs 3//2 = 1
    (6, 16, 16, 3, 1, None, None),       # temporal windows only
]


def main():
    Tester, M = load_reference_tester()
    out = {'cases': []}
    rng = np.random.RandomState(7)
    for (T, H, W, nf, ovf, wins, ov) in CASES:
        frames = rng.randint(0, 256, size=(T, H, W, 3)).astype(np.uint8)
        t = bare_tester(Tester, nf, ovf, wins, ov)
        u8 = t.evaluate(frames)                                            # tools/Tester.py:217-252 (test_clips route)
        out['cases'].append({'args': (T, H, W, nf, ovf, wins, ov), 'frames': torch.from_numpy(frames), 'u8': torch.from_numpy(u8)})
    a = rng.randint(0, 256, size=(3, 40, 52, 3)).astype(np.uint8)
    b = np.clip(a.astype(np.int32) + rng.randint(-9, 10, size=a.shape), 0, 255).astype(np.uint8)
    out['metrics'] = {'a': torch.from_numpy(a), 'b': torch.from_numpy(b),
                      'psnr': [M.calculate_psnr(a[i], b[i]) for i in range(3)],
                      'ssim': [float(M.structural_similarity(a[i], b[i])) for i in range(3)],
                      'psnr_border2': M.calculate_psnr(a[0], b[0], border=2)}
    aug = bare_tester(Tester, 4, 2, None, None)
    x = torch.from_numpy(rng.rand(1, 2, 3, 6, 10).astype(np.float32))
    out['augment'] = {'x': x, 'modes': [(h, v, r, aug.augment(x, h, v, r)) for h in (False, True) for v in (False, True) for r in (False, True)]}
    torch.save(out, os.path.join(ROOT, 'tests', 'golden', 'eval', 'evalpath.pt'))
    print('minted tests/golden/eval/evalpath.pt', os.path.getsize(os.path.join(ROOT, 'tests', 'golden', 'eval', 'evalpath.pt')), 'bytes')


if __name__ == '__main__':
    main()
