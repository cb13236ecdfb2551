"""Frame preparation on the device — what the reference's callers do on the CPU before every forward
(run_inference.py:15-34 and :79-81; dota.py:347-357): `cv2.resize(img, (224, 224), interpolation=cv2.INTER_CUBIC)` on the
uint8 BGR frame, BGR -> RGB, / 255, ImageNet mean / std, HWC -> CHW.

`resize_cubic_u8` runs OpenCV's 8-bit fixed-point bicubic algorithm (imgproc/src/resize.cpp) in a CUDA kernel
(`stad_resize_cubic_u8`); the tap tables are built here on the host exactly as resize.cpp builds them (float32
arithmetic, A = -0.75, 11-bit weights).  `prepare_frames` = resize + `stad_normalize_frames_u8` -> bf16 planes
[F, 3, H, W], the frame buffer `VisionTransformer.forward_windows` reads.
"""
import functools

import numpy as np
import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
_COEF_BITS = 11  # INTER_RESIZE_COEF_BITS


@functools.lru_cache(maxsize=32)
def cubic_taps(n_dst, n_src):
    """(ofs int32 [n_dst], weights int16 [n_dst, 4]) of cv2's INTER_CUBIC for one axis: ofs[d] is the source index of
    the second tap (taps cover ofs-1 .. ofs+2, clamped to the image by the kernel)."""
    scale = n_src / n_dst
    d = np.arange(n_dst, dtype=np.float64)
    fx = ((d + 0.5) * scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int32)
    x = (fx - sx.astype(np.float32)).astype(np.float32)
    A, one = np.float32(-0.75), np.float32(1.0)
    c = np.empty((n_dst, 4), dtype=np.float32)
    c[:, 0] = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c[:, 1] = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    c[:, 2] = ((A + np.float32(2)) * (one - x) - (A + np.float32(3))) * (one - x) * (one - x) + one
    c[:, 3] = one - c[:, 0] - c[:, 1] - c[:, 2]
    w = np.rint(c * np.float32(1 << _COEF_BITS)).astype(np.int16)   # saturate_cast<short>: round half to even
    return sx, w


def _taps_on(device, n_dst, n_src, _cache={}):
    key = (str(device), n_dst, n_src)
    if key not in _cache:
        ofs, w = cubic_taps(n_dst, n_src)
        _cache[key] = (torch.from_numpy(ofs).to(device), torch.from_numpy(w).to(device).contiguous())
    return _cache[key]


def resize_cubic_u8(frames, size, out=None):
    """frames uint8 CUDA [F, H, W, 3] -> uint8 [F, size[0], size[1], 3] (cv2.resize INTER_CUBIC, ri:79-80)."""
    _lib.init(frames.device)
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3 or not frames.is_cuda:
        raise ValueError(f"resize_cubic_u8: expected CUDA uint8 frames [F, H, W, 3], got {frames.dtype} {tuple(frames.shape)} "
                         f"on {frames.device}")
    frames = frames.contiguous()
    F_, Hs, Ws, _ = frames.shape
    Hd, Wd = int(size[0]), int(size[1])
    if out is None:
        out = torch.empty(F_, Hd, Wd, 3, dtype=torch.uint8, device=frames.device)
    xofs, xw = _taps_on(frames.device, Wd, Ws)
    yofs, yw = _taps_on(frames.device, Hd, Hs)
    _lib.check(_lib.load().stad_resize_cubic_u8(_lib.ptr(frames), _lib.ptr(out), F_, Hs, Ws, Hd, Wd, _lib.ptr(xofs),
                                                _lib.ptr(xw), _lib.ptr(yofs), _lib.ptr(yw), _lib.stream_ptr()),
               "stad_resize_cubic_u8")
    return out


def prepare_frames(frames_u8, size=(224, 224), bgr=True, mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None):
    """uint8 CUDA frames [F, H, W, 3] of any size -> normalised bf16 planes [F, 3, size[0], size[1]]: the reference's
    cv2.resize + prepare_image (ri:79-81, ri:15-34) without leaving the device."""
    if tuple(frames_u8.shape[1:3]) != tuple(size):
        frames_u8 = resize_cubic_u8(frames_u8, size)
    return _lib.normalize_frames_u8(frames_u8.contiguous(), mean, std, bgr=bgr, out=out)
