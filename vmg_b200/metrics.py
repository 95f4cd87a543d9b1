"""Device-side evaluation metrics -- mirrors `utils/metrics.py` of the reference (calculate_psnr :11-28,
structural_similarity :33-70) and the luma conversion used by `tools/test_reds4.py:194-218`
(`skimage.color.rgb2ycbcr(...)[..., 0]`), for frames that already live on the GPU.

Inputs are what the reference feeds them: uint8 [H,W,3] RGB frames (or float64 [H,W] luma planes), also batched
over a leading frame axis.  PSNR on uint8 is bit-identical to the reference (the squared-error sum is an exact
integer); luma and SSIM are float64 like numpy / cv2 and agree to ~1e-12 (summation order differs).
"""
import math

import torch

from . import ops as _ops


def gaussian_window(ksize=11, sigma=1.5):
    """cv2.getGaussianKernel(ksize, sigma) for sigma > 0: exp(-(i-(k-1)/2)^2 / (2 sigma^2)) normalised to sum 1."""
    c = (ksize - 1) * 0.5
    scale2x = -0.5 / (sigma * sigma)
    w = [math.exp(scale2x * (i - c) * (i - c)) for i in range(ksize)]
    s = 1.0 / sum(w)
    return [v * s for v in w]


_WIN = gaussian_window()


def _frames(img):
    return img if img.dim() == 4 else img.unsqueeze(0)


def calculate_psnr(img1, img2, border=0):
    """utils/metrics.py:11-28.  uint8 [H,W,3] / [H,W] or float64 planes, range [0,255]; returns a python float
    (inf for identical images)."""
    if img1.shape != img2.shape:
        raise ValueError('Input images must have the same dimensions.')
    if border:
        h, w = img1.shape[:2]
        img1 = img1[border:h - border, border:w - border]
        img2 = img2[border:h - border, border:w - border]
    img1, img2 = img1.contiguous(), img2.contiguous()
    n = img1.numel()
    if img1.dtype == torch.uint8:
        sse = float(_ops.sse_u8(img1, img2))
    else:
        sse = float(_ops.sse_f64(img1.double().contiguous(), img2.double().contiguous()).item())
    mse = sse / n
    if mse == 0:
        return float('inf')
    return 20 * math.log10(255.0 / math.sqrt(mse))


def rgb2ycbcr_y(img):
    """Y channel of skimage.color.rgb2ycbcr for uint8 RGB [..., 3] -> float64 [...] in [16, 235]."""
    return _ops.rgb_to_y(img.contiguous())


def structural_similarity(img1, img2):
    """utils/metrics.py:33-70.  uint8 [H,W,3] (mean over channels) or float64 / uint8 [H,W]; python float."""
    if img1.shape != img2.shape:
        raise ValueError('Input images must have the same dimensions.')
    if img1.dim() == 2 and img1.dtype == torch.uint8:
        img1, img2 = img1.double(), img2.double()
    a, b = (img1.unsqueeze(0).contiguous(), img2.unsqueeze(0).contiguous())
    return float(_ops.ssim(a, b, _WIN)[0].item())


def evaluate_frames(outputs, targets):
    """The per-frame loop of tools/test_reds4.py:194-218 for a whole clip in four launches per metric:
    outputs / targets uint8 [T,H,W,3] on the device -> dict of float64 tensors [T] (psnr, psnr_y, ssim, ssim_y)."""
    assert outputs.shape == targets.shape and outputs.dtype == torch.uint8 and targets.dtype == torch.uint8
    T = outputs.shape[0]
    outputs, targets = outputs.contiguous(), targets.contiguous()
    oy, ty = rgb2ycbcr_y(outputs), rgb2ycbcr_y(targets)
    res = {'psnr': [calculate_psnr(outputs[t], targets[t]) for t in range(T)],
           'psnr_y': [calculate_psnr(oy[t], ty[t]) for t in range(T)]}
    res['ssim'] = _ops.ssim(outputs, targets, _WIN).tolist()
    res['ssim_y'] = _ops.ssim(oy, ty, _WIN).tolist()
    return res
