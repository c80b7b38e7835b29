"""ORACLE (test infrastructure, not product code).

numpy restatement of the reference image preprocessing, tools/infer.py:121-131 (letterbox) and :432-453
(BGR->RGB, /255, ImageNet mean/std, CHW), including a bit-exact restatement of OpenCV's 8-bit
``cv2.resize(..., interpolation=cv2.INTER_LINEAR)`` fixed-point arithmetic (OpenCV is a third-party
dependency of the reference, requirements.txt `opencv-python`; the image ships cv2 4.13 so the restatement
is pinned against cv2 itself in tests/test_oracle_pre.py).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
MEAN = np.array([0.485, 0.456, 0.406], dtype=f32)
STD = np.array([0.229, 0.224, 0.225], dtype=f32)


def _coefs(n_dst: int, n_src: int, scale: float, clamp: bool):
    d = np.arange(n_dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(f32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(f32)).astype(f32)
    if clamp:                       # x direction: fraction reset at the borders
        lo, hi = s < 0, s >= n_src - 1
        f = np.where(lo | hi, f32(0), f)
        s = np.where(lo, 0, np.where(hi, n_src - 1, s))
    a0 = np.rint((f32(1.0) - f).astype(f32) * f32(2048)).astype(np.int64)
    a1 = np.rint(f * f32(2048)).astype(np.int64)
    i0 = np.clip(s, 0, n_src - 1)
    i1 = np.clip(s + 1, 0, n_src - 1)
    return i0, i1, a0, a1


def resize_linear_u8(src: np.ndarray, nw: int, nh: int) -> np.ndarray:
    h, w = src.shape[:2]
    if (h, w) == (nh, nw):
        return src.copy()
    sx, sy = 1.0 / (nw / w), 1.0 / (nh / h)
    x0, x1, ax0, ax1 = _coefs(nw, w, sx, True)
    y0, y1, ay0, ay1 = _coefs(nh, h, sy, False)
    s = src.astype(np.int64)
    rows = s[:, x0, :] * ax0[None, :, None] + s[:, x1, :] * ax1[None, :, None]
    r0, r1 = rows[y0], rows[y1]
    out = ((((ay0[:, None, None] * (r0 >> 4)) >> 16) + ((ay1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2)
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox_ref(im: np.ndarray, new_size: int, color=(114, 114, 114)):
    """tools/infer.py:121-131 -> (padded uint8 image, scale, (left, top))."""
    h, w = im.shape[:2]
    scale = min(new_size / h, new_size / w)
    nh, nw = int(round(h * scale)), int(round(w * scale))
    res = resize_linear_u8(im, nw, nh)
    top = (new_size - nh) // 2
    left = (new_size - nw) // 2
    out = np.empty((new_size, new_size, 3), np.uint8)
    out[...] = np.array(color, np.uint8)
    out[top:top + nh, left:left + nw] = res
    return out, scale, (left, top)


def normalise_ref(lb_bgr: np.ndarray) -> np.ndarray:
    """tools/infer.py:448-452: uint8 HWC BGR -> fp32 [1,3,S,S]."""
    im = lb_bgr[..., ::-1].astype(f32) / f32(255.0)
    im = (im - MEAN) / STD
    return np.ascontiguousarray(np.transpose(im, (2, 0, 1))[None]).astype(f32)


def preprocess_ref(im_bgr: np.ndarray, img_size: int):
    lb, scale, (left, top) = letterbox_ref(im_bgr, img_size)
    return normalise_ref(lb), scale, left, top
