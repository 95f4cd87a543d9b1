"""Tester -- device-side mirror of the reference's evaluation driver `tools/Tester.py` for the steps either
side of the model forward: uint8 frames in, overlapped temporal windows and spatial
tiles through the model, mask blending, clamp + quantisation, uint8 frames out.

Same names, arguments and semantics as the reference class (tools/Tester.py:36-98 config handling,
:107-143 test_image, :146-177 test_clips, :180-215 test_clips_max, :217-252 evaluate, :285-309 evaluate_fps,
:387-445 augment / augment_inverse).  Everything stays on the GPU: the reference slices tiles in Python, keeps
ones_like masks and round-trips every clip through numpy; here tiles are blended by `vmgb_blend_accumulate`
in the reference's loop order (bit-identical fp32 sums), normalised by a count plane and quantised by
`vmgb_blend_finalize_u8` -- one device->host copy of uint8 frames at the end.
"""
import time

import numpy as np
import torch

from . import ops as _ops


def _starts(length, win, overlap):
    """tools/Tester.py:112-113,155: list(range(0, L-win, win-ov)) + [max(0, L-win)]"""
    return list(range(0, length - win, win - overlap)) + [max(0, length - win)]


def _zeroed(n, sl):
    """How many leading / trailing entries of an axis of length n the Python slice `sl` covers (the reference zeroes
    `x[..., -ov//2:]` and `x[..., :ov//2]`; note -ov//2 == (-ov)//2 rounds away from zero, and -0//2 == 0 selects all)."""
    a, b, _ = sl.indices(n)
    return max(0, b - a)


class Tester:
    def __init__(self, config, model=None, precision='bf16'):
        """config: the dict of an eval YAML (configs/*-eval-*.yml of the reference) -- only the keys the reference's
        Tester reads on this path are used.  model: a ready vmg_b200.VMG (else built with create_model)."""
        ds = config['dataset']
        self.config = config
        self.hflip = ds.get('use_hflip', False)
        self.vflip = ds.get('use_vflip', False)
        self.rot = ds.get('use_rot', False)
        self.data_enhance = ds.get('data_enhance', False)
        self.dataset_name = ds.get('name')
        self.scale = config['scale']
        self.test_num_frames = ds['num_frames']
        mode = ds['overlapped_mode']
        if mode == 'small':
            self.overlapped_num_frames = 2
        elif mode == 'mid':
            self.overlapped_num_frames = ds['num_frames'] // 2
        elif mode == 'large':
            self.overlapped_num_frames = ds['num_frames'] - 1
        elif not isinstance(mode, str):
            self.overlapped_num_frames = int(mode)
        else:
            raise Exception('choose right mode of testing.')
        self.test_spatial = ds.get('wins')
        self.overlapped_spatial_length = ds.get('overlapped_spatial_length')
        if model is None:
            from .config import create_model
            model = create_model(config, precision=precision).eval().cuda()
        self.model = model
        self.device = next(model.parameters()).device
        self.t_idx_list = None

    # ------------------------------------------------------------------------------------------
    # unit enumeration (shared with vmg_b200.shard for multi-GPU runs)
    # ------------------------------------------------------------------------------------------
    def tile_lists(self, H, W):
        ov = self.overlapped_spatial_length
        return _starts(H, self.test_spatial[0], ov), _starts(W, self.test_spatial[1], ov)

    def window_list(self, T):
        stride = self.test_num_frames - self.overlapped_num_frames
        return _starts_t(T, self.test_num_frames, stride)

    def tile_masks(self, h_idx, w_idx, h_list, w_list, ph, pw):
        """(zero_h, zero_w) = leading / trailing rows and columns of the OUTPUT patch [.., ph, pw] the reference zeroes."""
        ov = self.overlapped_spatial_length
        tail, head = slice(-ov // 2, None), slice(None, ov // 2)
        zh = [0, 0]
        zw = [0, 0]
        if h_idx < h_list[-1]:
            zh[1] = _zeroed(ph, tail)
        if w_idx < w_list[-1]:
            zw[1] = _zeroed(pw, tail)
        if h_idx > h_list[0]:
            zh[0] = _zeroed(ph, head)
        if w_idx > w_list[0]:
            zw[0] = _zeroed(pw, head)
        return tuple(zh), tuple(zw)

    def window_masks(self, t_idx, t_list, n):
        ovf = self.overlapped_num_frames
        zt = [0, 0]
        if ovf > 0:
            if t_idx < t_list[-1]:
                zt[1] = _zeroed(n, slice(-ovf // 2, None))
            if t_idx > t_list[0]:
                zt[0] = _zeroed(n, slice(None, ovf // 2))
        return tuple(zt)

    # ------------------------------------------------------------------------------------------
    def _forward(self, x):
        self.model.num_out_frames = x.shape[1]
        return self.model(x.contiguous())

    @torch.no_grad()
    def test_image(self, inputs):
        """tools/Tester.py:107-143.  inputs [B,T,C,H,W] on the device -> [B,T,C,sH,sW]."""
        B, T, C, H, W = inputs.shape
        s = self.scale
        h_list, w_list = self.tile_lists(H, W)
        E = torch.zeros((B, T, C, H * s, W * s), device=inputs.device, dtype=torch.float32)
        cnt = torch.zeros((T, H * s, W * s), device=inputs.device, dtype=torch.float32)
        for h_idx in h_list:
            for w_idx in w_list:
                in_patch = inputs[..., h_idx:h_idx + self.test_spatial[0], w_idx:w_idx + self.test_spatial[1]]
                out_patch = self._forward(in_patch)
                zh, zw = self.tile_masks(h_idx, w_idx, h_list, w_list, out_patch.shape[-2], out_patch.shape[-1])
                _ops.blend_accumulate_(E, cnt, out_patch, 0, h_idx * s, w_idx * s, (0, 0), zh, zw)
        return _ops.blend_normalize_(E, cnt)

    @torch.no_grad()
    def test_clips(self, inputs):
        """tools/Tester.py:146-177."""
        B, T, C, H, W = inputs.shape
        s = self.scale
        if T < self.test_num_frames:      # the reference's N[:, t:t+n].add_(N_clips) fails to broadcast here
            raise RuntimeError(f'test_clips needs at least dataset.num_frames={self.test_num_frames} frames, got {T}')
        E = torch.zeros((B, T, C, H * s, W * s), device=inputs.device, dtype=torch.float32)
        cnt = torch.zeros((T, H * s, W * s), device=inputs.device, dtype=torch.float32)
        self.t_idx_list = self.window_list(T)
        for t_idx in self.t_idx_list:
            clips = inputs[:, t_idx:t_idx + self.test_num_frames]
            out = self._forward(clips) if self.overlapped_spatial_length is None else self.test_image(clips)
            zt = self.window_masks(t_idx, self.t_idx_list, out.shape[1])
            _ops.blend_accumulate_(E, cnt, out, t_idx, 0, 0, zt, (0, 0), (0, 0))
        return _ops.blend_normalize_(E, cnt)

    @torch.no_grad()
    def test_clips_max(self, inputs, HR):
        """tools/Tester.py:180-215: every frame is taken from the temporal window whose output has the best PSNR
        against HR (skimage peak_signal_noise_ratio on clamped [0,1] frames: monotone in the MSE, so the window with
        the smallest squared error wins; ties -> the first window, like torch.max)."""
        B, T, C, H, W = inputs.shape
        assert B == 1
        s = self.scale
        stride = self.test_num_frames - self.overlapped_num_frames
        self.t_idx_list = _starts_t(T, self.test_num_frames, stride)
        out = torch.zeros((B, T, C, H * s, W * s), device=inputs.device, dtype=torch.float32)
        best = [None] * T
        for t_idx in self.t_idx_list:
            clips = inputs[:, t_idx:t_idx + self.test_num_frames]
            o = self._forward(clips) if self.overlapped_spatial_length is None else self.test_image(clips)
            for i in range(o.shape[1]):
                sse = float(_ops.sse_f32(o[:, i].contiguous(), HR[:, t_idx + i].contiguous(), clamp01=True).item())
                if best[t_idx + i] is None or sse < best[t_idx + i]:
                    best[t_idx + i] = sse
                    out[:, t_idx + i].copy_(o[:, i])
        return out

    # ------------------------------------------------------------------------------------------
    def _to_device_u8(self, frames):
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(np.ascontiguousarray(frames))
        assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3, 'frames must be uint8 [T,H,W,3]'
        if not frames.is_cuda:
            frames = frames.pin_memory().to(self.device, non_blocking=True)
        return frames.contiguous()

    @torch.no_grad()
    def evaluate_device(self, inputs, HR=None):
        """tools/Tester.py:217-252 without the final host copy: uint8 [T,H,W,3] (numpy / host / device) ->
        device uint8 [T,sH,sW,3]."""
        x = _ops.frames_u8_to_f32(self._to_device_u8(inputs))           # [1,T,3,H,W] = u8 / 255
        if self.data_enhance:
            x = self.augment(x, self.hflip, self.vflip, self.rot)
        hr = None
        if HR is not None:
            hr = _ops.frames_u8_to_f32(self._to_device_u8(HR))
            if self.data_enhance:
                hr = self.augment(hr, self.hflip, self.vflip, self.rot)
        if self.dataset_name == 'Vimeo90k_septuplet':
            out = self._forward(x) if self.overlapped_spatial_length is None else self.test_image(x)
        elif self.dataset_name == 'REDS':
            out = self.test_clips_max(x, hr)
        else:
            out = self.test_clips(x)
        if self.data_enhance:
            out = self.augment_inverse(out, self.hflip, self.vflip, self.rot)
        return _ops.frames_f32_to_u8(out.contiguous())

    def evaluate(self, inputs, HR=None, extra_fps=False):
        """Returns numpy uint8 [T,sH,sW,3] like the reference."""
        return self.evaluate_device(inputs, HR).cpu().numpy()

    def evaluate_fps(self, inputs, num_frames=7, HR=None):
        """tools/Tester.py:285-309 (the non-REDS branch times `evaluate`; frames / second as the reference counts them)."""
        torch.cuda.synchronize()
        t0 = time.time()
        outputs = self.evaluate(inputs, HR)
        torch.cuda.synchronize()
        dt = time.time() - t0
        if self.dataset_name == 'Vimeo90k_septuplet':
            fps = float(num_frames * 1. / dt)
        else:
            fps = float(self.test_num_frames * len(self.t_idx_list or [0]) / dt)
        return outputs, fps

    # ------------------------------------------------------------------------------------------
    def augment(self, img_clip, hflip=True, vflip=True, rot90=True):
        """tools/Tester.py:387-415 on the device (B == 1 like the reference)."""
        assert img_clip.shape[0] == 1, 'testing.'
        return _ops.augment(img_clip.contiguous().float(), hflip, vflip, rot90)

    def augment_inverse(self, img_clip, hflip=True, vflip=True, rot90=True):
        """tools/Tester.py:417-445: the reference applies the SAME flip / flip / transpose sequence again."""
        assert img_clip.shape[0] == 1, 'testing.'
        return _ops.augment(img_clip.contiguous().float(), hflip, vflip, rot90)


def _starts_t(T, n, stride):
    """tools/Tester.py:154-155: list(range(0, T-n, stride)) + [max(0, T-n)]"""
    return list(range(0, T - n, stride)) + [max(0, T - n)]


ENSEMBLE_MODES = [(h, v, r) for r in (False, True) for v in (False, True) for h in (False, True)]


@torch.no_grad()
def self_ensemble_x8(forward, x):
    """x8 self-ensemble over the 8 flip / transpose combinations of Tester.augment: the mean of
    inverse-augmented outputs.  Each mode is an independent forward (a shardable unit).  Unlike augment_inverse of the
    reference (which re-applies the forward sequence and is therefore only correct for its involutive modes), the
    inverse here undoes transpose first, then the flips."""
    acc = None
    for h, v, r in ENSEMBLE_MODES:
        y = forward(_ops.augment(x.contiguous().float(), h, v, r))
        if r:
            y = _ops.augment(y.contiguous(), False, False, True)
        y = _ops.augment(y.contiguous(), h, v, False)
        acc = y if acc is None else _ops.add(acc, y)
    return acc / len(ENSEMBLE_MODES)
