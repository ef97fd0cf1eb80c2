"""TEST INFRASTRUCTURE (parity oracle) — numpy restatement of the reference's frame pre/post pixel math.

Integer formulas are pinned bit-exact against the libraries the reference calls (OpenCV / Pillow, both
installed here) by tests/test_pixel_oracle.py and by the golden fixtures made through the reference's own
functions (tests/golden/make_golden.py).  The zimg Spline64 resize is a restatement of the published
kernel — zimg/VapourSynth is absent, so THAT EDGE IS "parity unpinned" (SURVEY.md 8c).
"""
from __future__ import annotations

import math

import numpy as np

IMAGENET_MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
IMAGENET_STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


# ---- Pillow convert('LA').convert('RGB')  (ColorizerFilter._transform, deoldify/filters.py:92-93) -------
def pil_luma(rgb: np.ndarray) -> np.ndarray:
    """uint8 [...,3] -> uint8 [...]: L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16."""
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    return ((19595 * r + 38470 * g + 7471 * b + 0x8000) >> 16).astype(np.uint8)


# ---- OpenCV 8-bit COLOR_RGB2YUV / COLOR_YUV2RGB (Q14)  (filters.py:100-110, imfilters.py:312-321) --------
def _sat8(v):
    return np.clip(v, 0, 255)


def cv_rgb2yuv(rgb: np.ndarray) -> np.ndarray:
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    y = (4899 * r + 9617 * g + 1868 * b + 8192) >> 14
    u = _sat8(((b - y) * 8061 + (128 << 14) + 8192) >> 14)
    v = _sat8(((r - y) * 14369 + (128 << 14) + 8192) >> 14)
    return np.stack([y, u, v], -1).astype(np.uint8)


def cv_yuv2rgb(yuv: np.ndarray) -> np.ndarray:
    y, u, v = (yuv[..., i].astype(np.int64) for i in range(3))
    r = _sat8(y + (((v - 128) * 18678 + 8192) >> 14))
    g = _sat8(y + (((u - 128) * -6472 + (v - 128) * -9519 + 8192) >> 14))
    b = _sat8(y + (((u - 128) * 33292 + 8192) >> 14))
    return np.stack([r, g, b], -1).astype(np.uint8)


def chroma_post_process(color: np.ndarray, orig: np.ndarray) -> np.ndarray:
    """Luma of `orig`, chroma of `color` (ColorizerFilter._post_process filters.py:100-110 ==
    chroma_post_process vsslib/imfilters.py:312-321).  uint8 [H,W,3] each."""
    cy, oy = cv_rgb2yuv(color), cv_rgb2yuv(orig)
    hires = oy.copy()
    hires[..., 1:3] = cy[..., 1:3]
    return cv_yuv2rgb(hires)


# ---- normalise / de-normalise  (filters.py:50-67, fastai/vision/data.py:56-79,300) ------------------------
def normalize_gray(L: np.ndarray) -> np.ndarray:
    """uint8 [H,W] -> float32 [3,H,W]: (L/255 - mean)/std per channel (the 3 channels are the same gray)."""
    x = L.astype(np.float32) / np.float32(255.0)
    return (x[None] - IMAGENET_MEAN[:, None, None]) / IMAGENET_STD[:, None, None]


def denorm_quantize(y: np.ndarray) -> np.ndarray:
    """Network output float32 [3,H,W] -> uint8 [H,W,3]: y*std+mean, clamp(0,1), *255, astype(uint8) = TRUNCATION."""
    d = y.astype(np.float32) * IMAGENET_STD[:, None, None] + IMAGENET_MEAN[:, None, None]
    d = np.clip(d, np.float32(0), np.float32(1))
    return np.transpose((d * np.float32(255.0)).astype(np.uint8), (1, 2, 0))


# ---- zimg Spline64 / Spline36 (restatement; parity unpinned) -------------------------------------------------
def _spline64(x):
    x = abs(x)
    if x < 1:
        return ((49 / 41 * x - 6387 / 2911) * x - 3 / 2911) * x + 1
    if x < 2:
        t = x - 1
        return ((-24 / 41 * t + 4032 / 2911) * t - 2328 / 2911) * t
    if x < 3:
        t = x - 2
        return ((6 / 41 * t - 1008 / 2911) * t + 582 / 2911) * t
    if x < 4:
        t = x - 3
        return ((-1 / 41 * t + 168 / 2911) * t - 97 / 2911) * t
    return 0.0


def _spline36(x):
    x = abs(x)
    if x < 1:
        return ((13 / 11 * x - 453 / 209) * x - 3 / 209) * x + 1
    if x < 2:
        t = x - 1
        return ((-6 / 11 * t + 270 / 209) * t - 156 / 209) * t
    if x < 3:
        t = x - 2
        return ((1 / 11 * t - 45 / 209) * t + 26 / 209) * t
    return 0.0


_KERNELS = {"spline64": (_spline64, 4), "spline36": (_spline36, 3)}


def resize_matrix(src: int, dst: int, kernel: str = "spline64") -> np.ndarray:
    """zimg-style 1-D filter bank as a dense [dst, src] float64 matrix: half-pixel centres, support widened by
    the shrink ratio, normalised rows, out-of-range taps reflected back into the image."""
    f, support = _KERNELS[kernel]
    scale = dst / src
    step = min(scale, 1.0)
    n = max(int(math.ceil(support / step)) * 2, 1)
    m = np.zeros((dst, src))
    for i in range(dst):
        pos = (i + 0.5) / scale
        begin = math.floor(pos - n / 2 + 0.5) + 0.5
        w = np.array([f((begin + j - pos) * step) for j in range(n)])
        w /= w.sum()
        for j in range(n):
            xp = begin + j
            xr = -xp if xp < 0 else (2 * src - xp if xp >= src else xp)
            m[i, min(max(int(math.floor(xr)), 0), src - 1)] += w[j]
    return m


def resize_plane_u8(img: np.ndarray, out_w: int, out_h: int, kernel: str = "spline64") -> np.ndarray:
    """uint8 [H,W] (or [H,W,C]) -> uint8 resized; horizontal pass then vertical pass in float32,
    round-half-even and clamp at the end (no dithering)."""
    h, w = img.shape[:2]
    mh = resize_matrix(w, out_w, kernel).astype(np.float32)
    mv = resize_matrix(h, out_h, kernel).astype(np.float32)
    x = img.astype(np.float32)
    if x.ndim == 2:
        t = x @ mh.T
        o = mv @ t
    else:
        t = np.einsum("hwc,ow->hoc", x, mh)
        o = np.einsum("ph,hoc->poc", mv, t)
    return np.clip(np.rint(o), 0, 255).astype(np.uint8)
