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


def _banded(m: np.ndarray):
    """dense [dst, src] filter matrix -> (start [dst], weights [dst, T]) of its non-zero band."""
    nz = m != 0.0
    first = nz.argmax(1)
    last = m.shape[1] - 1 - nz[:, ::-1].argmax(1)
    T = int((last - first).max()) + 1
    start = np.clip(np.minimum(first, m.shape[1] - T), 0, None)
    w = np.zeros((m.shape[0], T))
    for o in range(m.shape[0]):
        seg = m[o, start[o]:start[o] + T]
        w[o, :len(seg)] = seg
    return start, w


_BAND_CACHE = {}


def _resize_axis(x: np.ndarray, axis: int, dst: int, kernel: str) -> np.ndarray:
    """One separable pass along `axis` of a float32 array, tap by tap over the filter's band (the dense [dst, src] product of
    round 1 spent 7 s per 1080p frame in zeros); float32 accumulation like zimg's float path."""
    src = x.shape[axis]
    key = (src, dst, kernel)
    if key not in _BAND_CACHE:
        _BAND_CACHE[key] = _banded(resize_matrix(src, dst, kernel))
    start, w = _BAND_CACHE[key]
    xm = np.moveaxis(x, axis, -1)                                   # [..., src]
    out = np.zeros(xm.shape[:-1] + (dst,), np.float32)
    for t in range(w.shape[1]):
        idx = np.minimum(start + t, src - 1)                        # taps past the band carry zero weight
        out += xm[..., idx] * w[:, t].astype(np.float32)
    return np.moveaxis(out, -1, axis)


def resize_plane_u8(img: np.ndarray, out_w: int, out_h: int, kernel: str = "spline64") -> np.ndarray:
    """uint8 [H,W] (or [H,W,C]) -> uint8 resized; two separable float32 passes with a float intermediate, the
    cheaper pass first (horizontal-then-vertical when the width shrinks more work away, vertical-then-horizontal
    otherwise - zimg orders its passes by cost too), round-half-even and clamp at the end (no dithering)."""
    h, w = img.shape[:2]
    x = img.astype(np.float32)
    if x.ndim == 2:
        x = x[..., None]
    h_first = out_w * h <= out_h * w          # size of the intermediate image = cost of the second pass
    if h_first:
        o = _resize_axis(_resize_axis(x, 1, out_w, kernel), 0, out_h, kernel)
    else:
        o = _resize_axis(_resize_axis(x, 0, out_h, kernel), 1, out_w, kernel)
    o = np.clip(np.rint(o), 0, 255).astype(np.uint8)
    return o[..., 0] if img.ndim == 2 else o


# ---- Pillow Image.resize (ImagingResample: separable, 8-bit fixed-point coefficients, u8 intermediate) ------
# Used by BaseFilter._scale_to_square / _unsquare with BILINEAR (deoldify/filters.py:37-41,70-73) and by
# colorizers/util.py:21-22 with BICUBIC.  Restated from Pillow's libImaging/Resample.c; pinned bit-exact
# against the installed Pillow by tests/test_pixel_oracle.py.
_PRECISION_BITS = 32 - 8 - 2


def _bilinear_filter(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def _bicubic_filter(x, a=-0.5):
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


_PIL_FILTERS = {"bilinear": (_bilinear_filter, 1.0), "bicubic": (_bicubic_filter, 2.0)}


def pil_coeffs(in_size: int, out_size: int, filt: str = "bilinear"):
    """(bounds [out,2] = (xmin, count), integer coefficients [out, ksize]) as Pillow's precompute_coeffs +
    normalize_coeffs_8bpc build them (float64 weights, normalised, then rounded to 22-bit fixed point)."""
    f, support0 = _PIL_FILTERS[filt]
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = support0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.array([f((x + xmin - center + 0.5) * ss) for x in range(xmax)], dtype=np.float64)
        ww = w.sum()
        if ww != 0.0:
            w = w / ww
        for x in range(xmax):
            v = w[x] * (1 << _PRECISION_BITS)
            kk[xx, x] = int(v - 0.5) if v < 0 else int(v + 0.5)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pil_pass(img: np.ndarray, out_size: int, axis: int, filt: str) -> np.ndarray:
    """One separable pass along `axis` (0 = vertical, 1 = horizontal) on uint8 [H,W,C]."""
    in_size = img.shape[axis]
    bounds, kk = pil_coeffs(in_size, out_size, filt)
    src = np.moveaxis(img.astype(np.int64), axis, 0)           # [in, ...]
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, cnt = bounds[xx]
        acc = np.tensordot(kk[xx, :cnt], src[xmin:xmin + cnt], axes=(0, 0)) + (1 << (_PRECISION_BITS - 1))
        out[xx] = np.clip(acc >> _PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis)


def pil_resize(img: np.ndarray, out_w: int, out_h: int, filt: str = "bilinear") -> np.ndarray:
    """uint8 [H,W,C] -> uint8 [out_h,out_w,C]; horizontal pass first, then vertical, each only if the size
    changes (Pillow skips identity passes), with a uint8 intermediate image."""
    h, w = img.shape[:2]
    x = img
    if out_w != w:
        x = _pil_pass(x, out_w, 1, filt)
    if out_h != h:
        x = _pil_pass(x, out_h, 0, filt)
    return x


def pil_blend(a: np.ndarray, b: np.ndarray, alpha: float) -> np.ndarray:
    """PIL.Image.blend(a, b, alpha) (ModelImageRender, deoldify/visualize.py:129,135; image_weighted_merge,
    vsslib/imfilters.py:113-124): a + alpha*(b-a) in float32; inside [0,1] the result is truncated, outside it is
    clipped to [0,255] first."""
    if alpha == 0.0:
        return a.copy()
    if alpha == 1.0:
        return b.copy()
    af, bf = a.astype(np.float32), b.astype(np.float32)
    t = af + np.float32(alpha) * (bf - af)
    if 0.0 <= alpha <= 1.0:
        return t.astype(np.uint8)
    return np.clip(t, 0, 255).astype(np.uint8)
