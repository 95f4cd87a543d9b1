"""Frame I/O either side of the forward: what `utils/eval_utils.py:8-36`
(read_image / read_seq_images: cv2 BGR -> RGB uint8 [N,H,W,C]) and `tools/Tester.py:217-222,249-250`
(uint8 -> fp32 / 255 -> [1,T,C,H,W]; clamp, *255, round -> uint8) do on the host with numpy, done on the
device with pinned, asynchronous copies: the host only ever touches uint8 frames (1/4 of the fp32 bytes
the reference moves over PCIe in each direction).
"""
import numpy as np
import torch

from . import ops as _ops


class FrameUploader:
    """Double-buffered pinned staging for uint8 clips: `upload` returns a device fp32 [1,T,3,H,W] tensor in [0,1]."""

    def __init__(self, device='cuda'):
        self.device = torch.device(device)
        self._pinned = {}
        self._flip = 0
        self.stream = torch.cuda.Stream(device=self.device)

    def _staging(self, shape):
        key = (tuple(shape), self._flip)
        self._flip ^= 1
        buf = self._pinned.get(key)
        if buf is None:
            buf = torch.empty(shape, dtype=torch.uint8).pin_memory()
            self._pinned[key] = buf
        return buf

    def upload(self, frames, bgr=False):
        """frames: numpy / torch uint8 [T,H,W,3] in host memory (RGB, or BGR as cv2.imread returns when bgr=True)."""
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(np.ascontiguousarray(frames))
        assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3
        stage = self._staging(frames.shape)
        stage.copy_(frames)
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.stream):
            dev = stage.to(self.device, non_blocking=True)
            x = _ops.frames_u8_to_f32(dev, reverse_channels=bgr)
        cur.wait_stream(self.stream)
        x.record_stream(cur)
        return x


def download_frames(frames_f32, bgr=False, out_host=None):
    """fp32 [1,T,3,H,W] / [T,3,H,W] on the device -> uint8 [T,H,W,3] in pinned host memory (RGB, or BGR for
    cv2.imwrite when bgr=True).  Asynchronous on the current stream: synchronise before reading the result."""
    u8 = _ops.frames_f32_to_u8(frames_f32.contiguous(), reverse_channels=bgr)
    if out_host is None:
        out_host = torch.empty(u8.shape, dtype=torch.uint8).pin_memory()
    out_host.copy_(u8, non_blocking=True)
    return out_host
