"""TEST INFRASTRUCTURE -- CPU restatement (numpy / torch-CPU) of the reference code either side of the model
forward: `tools/Tester.py` (tiling, window blending, quantisation, augmentation), `utils/metrics.py`
(PSNR, SSIM) and the luma conversion of `tools/test_reds4.py`.  The checker for vmg_b200/tester.py,
vmg_b200/metrics.py and csrc/evalpath.cu; never imported by the product.

Pinning: `tests/test_eval_oracle.py` runs every function here against the UNMODIFIED reference classes
(`tools.Tester.Tester` methods called on a bare instance, `utils.metrics`) in the build container and against
`tests/golden/eval/evalpath.pt`, minted from them by `oracle/mint_eval_goldens.py`.
Third-party arithmetic that is absent here: `skimage.color.rgb2ycbcr` / `skimage.metrics.peak_signal_noise_ratio`
(scikit-image, imported by tools/test_reds4.py:15 and tools/Tester.py:20; no version pinned by the reference) --
restated from the published algorithm (ITU-R BT.601 studio-swing matrix on uint8 / 255), pinned only by its
known-answer values (black -> 16, white -> 235) -> "parity unpinned" for the luma conversion.
"""
import math

import numpy as np
import torch


# ---------------------------------------------------------------------------------------------
# tools/Tester.py
# ---------------------------------------------------------------------------------------------
def overlapped_num_frames(num_frames, mode):
    """tools/Tester.py:50-60"""
    if mode == 'small':
        return 2
    if mode == 'mid':
        return num_frames // 2
    if mode == 'large':
        return num_frames - 1
    if not isinstance(mode, str):
        return int(mode)
    raise Exception('choose right mode of testing.')


def test_image(model, inputs, wins, ov, scale):
    """tools/Tester.py:107-143 (torch CPU, same statements)."""
    B, T, C, H, W = inputs.shape
    stride_h, stride_w = wins[0] - ov, wins[1] - ov
    h_idx_list = list(range(0, H - wins[0], stride_h)) + [max(0, H - wins[0])]
    w_idx_list = list(range(0, W - wins[1], stride_w)) + [max(0, W - wins[1])]
    E = inputs.new_zeros(B, T, C, H * scale, W * scale)
    Wm = torch.zeros_like(E)
    for h_idx in h_idx_list:
        for w_idx in w_idx_list:
            out_patch = model(inputs[..., h_idx:h_idx + wins[0], w_idx:w_idx + wins[1]]).clone()
            mask = torch.ones_like(out_patch)
            if h_idx < h_idx_list[-1]:
                out_patch[..., -ov // 2:, :] *= 0
                mask[..., -ov // 2:, :] *= 0
            if w_idx < w_idx_list[-1]:
                out_patch[..., :, -ov // 2:] *= 0
                mask[..., :, -ov // 2:] *= 0
            if h_idx > h_idx_list[0]:
                out_patch[..., :ov // 2, :] *= 0
                mask[..., :ov // 2, :] *= 0
            if w_idx > w_idx_list[0]:
                out_patch[..., :, :ov // 2] *= 0
                mask[..., :, :ov // 2] *= 0
            E[..., h_idx * scale:(h_idx + wins[0]) * scale, w_idx * scale:(w_idx + wins[1]) * scale].add_(out_patch)
            Wm[..., h_idx * scale:(h_idx + wins[0]) * scale, w_idx * scale:(w_idx + wins[1]) * scale].add_(mask)
    return E.div_(Wm)


def test_clips(model, inputs, num_frames, ov_frames, wins, ov, scale):
    """tools/Tester.py:146-177"""
    B, T, C, H, W = inputs.shape
    E = inputs.new_zeros(B, T, C, H * scale, W * scale)
    N = inputs.new_zeros(B, T, 1, 1, 1)
    stride = num_frames - ov_frames
    t_idx_list = list(range(0, T - num_frames, stride)) + [max(0, T - num_frames)]
    for t_idx in t_idx_list:
        clips = inputs[:, t_idx:t_idx + num_frames]
        out = model(clips).clone() if ov is None else test_image(model, clips, wins, ov, scale)
        n_clips = inputs.new_ones(B, num_frames, 1, 1, 1)
        if ov_frames > 0:
            if t_idx < t_idx_list[-1]:
                out[:, -ov_frames // 2:] *= 0
                n_clips[:, -ov_frames // 2:] *= 0
            if t_idx > t_idx_list[0]:
                out[:, :ov_frames // 2] *= 0
                n_clips[:, :ov_frames // 2] *= 0
        E[:, t_idx:t_idx + num_frames].add_(out)
        N[:, t_idx:t_idx + num_frames].add_(n_clips)       # raises when T < num_frames, like the reference
    return E.div_(N)


def frames_to_input(frames_u8):
    """tools/Tester.py:219-221: uint8 [T,H,W,C] -> fp32 [1,T,C,H,W] / 255"""
    x = frames_u8.astype(np.float32) / 255.
    return torch.from_numpy(x).permute(0, 3, 1, 2).contiguous().unsqueeze(0)


def quantize(outputs):
    """tools/Tester.py:249-250: [1,T,C,H,W] fp32 -> uint8 [T,H,W,C]"""
    o = outputs.cpu().squeeze().clamp(0, 1).numpy()
    return np.round(np.ascontiguousarray(o.squeeze().transpose(0, 2, 3, 1)) * 255.0).astype(np.uint8)


def augment(img_clip, hflip=True, vflip=True, rot90=True):
    """tools/Tester.py:387-415 (and :417-445, which is the same sequence)."""
    B, D, C, H, W = img_clip.shape
    assert B == 1
    x = img_clip.numpy()
    if hflip:
        x = x[..., ::-1]
    if vflip:
        x = x[..., ::-1, :]
    if rot90:
        x = x.transpose(0, 1, 2, 4, 3)
    return torch.from_numpy(np.ascontiguousarray(x)).float()


# ---------------------------------------------------------------------------------------------
# utils/metrics.py, tools/test_reds4.py
# ---------------------------------------------------------------------------------------------
def calculate_psnr(img1, img2, border=0):
    """utils/metrics.py:11-28"""
    h, w = img1.shape[:2]
    img1 = img1[border:h - border, border:w - border].astype(np.float64)
    img2 = img2[border:h - border, border:w - border].astype(np.float64)
    mse = np.mean((img1 - img2) ** 2)
    if mse == 0:
        return float('inf')
    return 20 * math.log10(255.0 / math.sqrt(mse))


def gaussian_kernel(ksize=11, sigma=1.5):
    """cv2.getGaussianKernel (utils/metrics.py:56): normalised exp(-(i - (k-1)/2)^2 / (2 sigma^2))"""
    i = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-0.5 / (sigma * sigma) * i * i)
    return k / k.sum()


def _filter_valid(img, window):
    """cv2.filter2D(img, -1, window)[5:-5, 5:-5]: correlation, only the samples that do not touch the border."""
    v = np.lib.stride_tricks.sliding_window_view(img, window.shape)
    return np.einsum('ijkl,kl->ij', v, window)


def structural_similarity(img1, img2):
    """utils/metrics.py:33-70"""
    if img1.ndim == 3:
        return np.array([structural_similarity(img1[..., i], img2[..., i]) for i in range(img1.shape[2])]).mean()
    C1, C2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    img1, img2 = img1.astype(np.float64), img2.astype(np.float64)
    k = gaussian_kernel(11, 1.5)
    window = np.outer(k, k)
    mu1, mu2 = _filter_valid(img1, window), _filter_valid(img2, window)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 ** 2, mu2 ** 2, mu1 * mu2
    s1 = _filter_valid(img1 ** 2, window) - mu1_sq
    s2 = _filter_valid(img2 ** 2, window) - mu2_sq
    s12 = _filter_valid(img1 * img2, window) - mu1_mu2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return ssim_map.mean()


def rgb2ycbcr_y(img_u8):
    """skimage.color.rgb2ycbcr(img)[..., 0] for uint8 RGB (tools/test_reds4.py:208-209): float64,
    Y = 16 + 65.481 R' + 128.553 G' + 24.966 B' with R'G'B' = uint8 * (1/255)."""
    f = img_u8.astype(np.float64) * (1.0 / 255.0)
    return 16.0 + f[..., 0] * 65.481 + f[..., 1] * 128.553 + f[..., 2] * 24.966
