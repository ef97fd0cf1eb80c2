"""TEST INFRASTRUCTURE (parity oracle) — numpy restatement of the reference's vsslib model-merge and
chroma-adjust filters.

Every function names the reference lines it follows.  The third-party arithmetic underneath (OpenCV 8-bit
YUV / HSV, Pillow blend / ImageEnhance / point) is restated as explicit integer / float32 formulas and pinned
bit-exact against the installed libraries by tests/test_filters_oracle.py; the composed filters are pinned
against the REAL reference functions through tests/golden/vsslib_filters.npz (tests/golden/make_golden.py).
`std.Merge` (VapourSynth) has no library here: its restatement is "parity unpinned".

All images are uint8 [H, W, 3] RGB.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .pixel_oracle import cv_rgb2yuv, cv_yuv2rgb, pil_blend, pil_luma

f32 = np.float32

# vsslib/constants.py:28-29
DEF_STANDARD_DARK = 0.22
DEF_STANDARD_BRIGHT = 0.78


# ---------------------------------------------------------------------------------------------------------
# OpenCV 8-bit HSV (H in [0,179]) — exact models (exhaustively checked against cv2 4.13 in the tests)
# ---------------------------------------------------------------------------------------------------------
def _hsv_tables():
    sdiv = np.zeros(256, np.int64)
    hdiv = np.zeros(256, np.int64)
    for i in range(1, 256):
        sdiv[i] = int(np.rint((255 << 12) / (1.0 * i)))
        hdiv[i] = int(np.rint((180 << 12) / (6.0 * i)))
    return sdiv, hdiv


_SDIV, _HDIV = _hsv_tables()


def cv_rgb2hsv(rgb: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(.., COLOR_RGB2HSV) on uint8: integer table model."""
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    v = np.maximum(np.maximum(r, g), b)
    vmin = np.minimum(np.minimum(r, g), b)
    d = v - vmin
    s = (d * _SDIV[v] + 2048) >> 12
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * d, r - g + 4 * d))
    hh = (h * _HDIV[d] + 2048) >> 12
    hh = np.where(hh < 0, hh + 180, hh)
    return np.stack([hh, s, v], -1).astype(np.uint8)


def _fma32(a, b, c):
    """float32 fused multiply-add (the vectorised OpenCV path contracts 1 - s*h): exact in float64, one rounding."""
    return (a.astype(np.float64) * b.astype(np.float64) + np.asarray(c, np.float64)).astype(f32)


# OpenCV converts rows in SIMD blocks and finishes each row with scalar code; for HSV -> RGB the two paths differ
# in the last step: the vector body TRUNCATES x*255, the scalar tail ROUNDS it (half to even).  The block is 32
# pixels on the AVX2 build used here (cv2 4.13.0); `simd_width` makes that host dependence explicit.
CV_HSV2RGB_SIMD_WIDTH = 32


def cv_hsv2rgb(hsv: np.ndarray, simd_width: int = CV_HSV2RGB_SIMD_WIDTH) -> np.ndarray:
    """cv2.cvtColor(.., COLOR_HSV2RGB) on uint8 [H, W, 3]: float32 sector model with the FMA contractions of the
    compiled code (1 - s*f and 1 - s*(1-f) are fused); x*255 truncated inside the per-row SIMD blocks and rounded
    half-to-even in the scalar tail.  Exhaustively identical to cv2 over all 2^24 inputs on both paths."""
    assert hsv.ndim == 3
    H = hsv[..., 0].astype(f32) * f32(6.0 / 180.0)
    S = hsv[..., 1].astype(f32) * f32(1.0 / 255.0)
    V = hsv[..., 2].astype(f32) * f32(1.0 / 255.0)
    sector = np.floor(H).astype(np.int32)
    fr = H - sector.astype(f32)
    sector = sector % 6
    one = f32(1.0)
    t0 = V
    t1 = V * (one - S)
    t2 = V * _fma32(-S, fr, one)
    t3 = V * _fma32(-S, one - fr, one)
    tab = np.stack([t0, t1, t2, t3], 0)
    sd = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])   # (b, g, r) table index per sector
    idx = sd[sector]
    pick = lambda k: np.take_along_axis(tab, idx[..., k][None], 0)[0]
    out = np.stack([pick(2), pick(1), pick(0)], -1) * f32(255.0)
    W = hsv.shape[1]
    body = (W // simd_width) * simd_width if simd_width > 0 else 0
    q = np.rint(out)
    q[:, :body] = np.floor(out[:, :body])
    return np.clip(q, 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------------------
# frame statistics
# ---------------------------------------------------------------------------------------------------------
def get_image_luma(img: np.ndarray, maxrange: int = 255) -> float:
    """imfilters.py:597-601: round(mean(cv Y) / maxrange, 6)."""
    y = cv_rgb2yuv(img)[..., 0]
    return round(float(np.mean(y)) / maxrange, 6)


# ---------------------------------------------------------------------------------------------------------
# Pillow ImageEnhance (exact: blend(degenerate, img, k); k outside [0,1] clips then truncates)
# ---------------------------------------------------------------------------------------------------------
def pil_blend_any(a: np.ndarray, b: np.ndarray, alpha: float) -> np.ndarray:
    """PIL.Image.blend for any alpha (libImaging/Blend.c): float32 a + alpha*(b-a); inside [0,1] plain truncation,
    outside clip to [0,255] first."""
    al = f32(alpha)
    t = a.astype(f32) + al * (b.astype(f32) - a.astype(f32))
    if 0.0 <= alpha <= 1.0:
        return t.astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t.astype(np.int32))).astype(np.uint8)


def enhance_color(img: np.ndarray, k: float) -> np.ndarray:
    """ImageEnhance.Color(img).enhance(k): degenerate = convert('L').convert('RGB')."""
    gray = np.repeat(pil_luma(img)[..., None], 3, -1)
    return pil_blend_any(gray, img, k)


def enhance_brightness(img: np.ndarray, k: float) -> np.ndarray:
    return pil_blend_any(np.zeros_like(img), img, k)


def enhance_contrast(img: np.ndarray, k: float) -> np.ndarray:
    """ImageEnhance.Contrast: degenerate = constant int(mean(L) + 0.5)."""
    L = pil_luma(img)
    mean = int(float(L.astype(np.int64).sum()) / L.size + 0.5)
    return pil_blend_any(np.full_like(img, mean), img, k)


def apply_gamma(img: np.ndarray, gamma: float) -> np.ndarray:
    """imfilters.py:507-517 builds a 256-entry numpy LUT and calls `img.point(table * 3)`: on a numpy array `* 3`
    scales the values instead of repeating the table, and Pillow rejects a 256-entry LUT for a 3-band image
    ("wrong number of lut entries").  The reference therefore RAISES for gamma != 1; so does the restatement."""
    raise ValueError("wrong number of lut entries")


# ---------------------------------------------------------------------------------------------------------
# hue-range mini language (restcolor.py:378-470)
# ---------------------------------------------------------------------------------------------------------
_HUE_NAMES = {"red": (0, 30), "orange": (30, 60), "yellow": (60, 90), "yellow-green": (90, 120), "green": (120, 150),
              "blue-green": (150, 180), "cyan": (180, 210), "blue": (210, 240), "blue-violet": (240, 270),
              "violet": (270, 300), "red-violet": (300, 330), "rose": (330, 360)}


def parse_hue_range(hue_range: str) -> List[Tuple[float, float]]:
    out = []
    for part in hue_range.split(","):
        if part in _HUE_NAMES:
            out.append(tuple(float(v) for v in _HUE_NAMES[part]))
            continue
        p = part.split(":")
        if len(p) == 2 and p[0].isnumeric() and p[1].isnumeric():
            out.append((float(p[0]), float(p[1])))
        else:
            raise ValueError("HybridAVC: unknown hue name: " + part)
    return out


def parse_hue_adjust(hue_adjust: str):
    """restcolor.py:378-410 -> (hue_range, sat, hue, weight) or None."""
    p = hue_adjust.split("|")
    sat, hue, weight = 1.0, 0, 0
    if len(p) < 1 or len(p) > 2:
        return None
    if len(p) == 1:
        return p[0], sat, hue, weight
    sw = p[1].split(",")

    def isfloat(x):
        try:
            float(x)
            return True
        except ValueError:
            return False
    if len(sw) != 2 or not isfloat(sw[0]) or not isfloat(sw[1]):
        return None
    if sw[0][0] in ("-", "+"):
        hue = int(sw[0])
    else:
        sat = float(sw[0])
    if sat > 10:
        hue = int(sat)
        sat = 1.0
    return p[0], sat, hue, float(sw[1])


def hue_mask(h: np.ndarray, hue_range: str) -> np.ndarray:
    """_build_hue_conditions (restcolor.py:412-428): cv hue strictly inside (min/2, max/2) of any range."""
    cond = np.zeros(h.shape, bool)
    for lo, hi in parse_hue_range(hue_range):
        cond |= (h > lo * 0.5) & (h < hi * 0.5)
    return cond


def np_hue_add(h: np.ndarray, hue: float) -> np.ndarray:
    """nputils.py:330-340 followed by the uint8 store of its callers (float -> uint8 truncation)."""
    if hue == 0:
        return h
    half = 0.5 * min(max(int(hue), -360), 360)
    x = h.astype(np.float64) + half
    x = np.where(x > 180, x - 180, x)
    x = np.where(x < 0, x + 180, x)
    return x.astype(np.uint8)


def _scale_u8(c: np.ndarray, k: float) -> np.ndarray:
    """`hsv[:, :, 1] = hsv[:, :, 1] * k`: float64 product stored into a uint8 array (C cast: truncation mod 256)."""
    return (c.astype(np.float64) * k).astype(np.int64).astype(np.uint8)


def mask_select(img1: np.ndarray, img2: np.ndarray, cond: np.ndarray) -> np.ndarray:
    """np_image_mask_merge with a 0/255 mask (nputils.py:203-219): exact select."""
    return np.where(cond[..., None], img2, img1)


def np_weighted_merge(a: np.ndarray, b: np.ndarray, w: float) -> np.ndarray:
    """nputils.py:265-283: float64 a*(1-w) + b*w, clip, truncate."""
    m = a.astype(np.float64) * (1 - w) + b.astype(np.float64) * w
    return m.clip(0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------------------
# luma-ramp merges (nputils.py:101-253, imfilters.py:66-100)
# ---------------------------------------------------------------------------------------------------------
def np_luma(img: np.ndarray) -> np.ndarray:
    """float64 (R*0.299 + G*0.587) + B*0.114, clipped to [0,255]."""
    return (img[..., 0] * 0.299 + img[..., 1] * 0.587 + img[..., 2] * 0.114).clip(0, 255)


def image_luma_merge(img_dark: np.ndarray, img_white: np.ndarray, luma: float) -> np.ndarray:
    """imfilters.py:66-78 with np_rgb_to_gray (nputils.py:101-123)."""
    lum = np_luma(img_white)
    if luma > 0:
        return mask_select(img_dark, img_white, lum > round(luma * 255))
    # threshold 0: the "mask" is the luma itself stored as uint8, normalised by 255
    mw = lum.astype(np.uint8).astype(np.float64) / 255
    m = img_dark * (1 - mw)[..., None] + img_white * mw[..., None]
    return m.clip(0, 255).astype(np.uint8)


def luma_ramp(img_white: np.ndarray, dark_luma: float, white_luma: float) -> np.ndarray:
    """w_np_rgb_to_gray(as_weight=True) (nputils.py:140-185): float32 ramp weight per pixel."""
    lum = np_luma(img_white)
    if dark_luma > 0:
        max_white = round(white_luma * 255)
        tresh = min(round(dark_luma * 255), max_white - 10)
        grad = round(1 / (max_white - tresh), 3)
        g = (lum - tresh) * grad
        return np.where(g > 1.0, 1.0, np.where(g < 0.0, 0.0, g)).astype(f32)
    return (lum / 255.0)


def w_image_luma_merge(img_dark: np.ndarray, img_white: np.ndarray, dark_luma: float, white_luma: float) -> np.ndarray:
    """imfilters.py:80-100 + w_np_image_mask_merge (nputils.py:228-253)."""
    if dark_luma >= white_luma:
        return img_dark
    mw = luma_ramp(img_white, dark_luma, white_luma).astype(np.float64)
    m = img_dark * (1 - mw)[..., None] + img_white * mw[..., None]
    return m.clip(0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------------------
# image_tweak (imfilters.py:463-504) — gamma LUT, Pillow enhancers, optional hue-range restriction
# ---------------------------------------------------------------------------------------------------------
def image_tweak(img: np.ndarray, sat: float = 1, cont: float = 1.0, bright: float = 0, hue: float = 0, gamma: float = 1.0,
                hue_range: str = "none") -> np.ndarray:
    out = img
    if gamma != 1.0:
        out = apply_gamma(out, gamma)
    if hue != 0.0:
        raise NotImplementedError("_apply_hue_shift (Pillow HSV) is not restated")
    if bright != 0.0:
        out = enhance_brightness(out, 1 + bright / 255)
    if cont != 1.0:
        out = enhance_contrast(out, cont)
    if sat != 1.0:
        out = enhance_color(out, sat)
    if hue_range in ("none", ""):
        return out
    return mask_select(img, out, hue_mask(cv_rgb2hsv(img)[..., 0], hue_range))   # np_adjust_chroma2, restcolor.py:344-370


def red_fix(img_stab: np.ndarray) -> np.ndarray:
    """Dark-frame red-shift adjustment shared by ConstrainedChromaMerge / ChromaBoundAdaptiveMerge (mcomb.py:351-362)."""
    luma = get_image_luma(img_stab, 255)
    if luma > 0.3:
        return img_stab
    if luma > 0.2:
        dark = image_tweak(img_stab, sat=0.9, hue_range="280:360,0:30")
        return w_image_luma_merge(dark, img_stab, 0.2, 0.3)
    if luma > 0.1:
        dark = image_tweak(img_stab, sat=0.8, hue_range="280:360,0:30")
        return w_image_luma_merge(dark, img_stab, 0.1, 0.2)
    return image_tweak(img_stab, sat=0.7)


# ---------------------------------------------------------------------------------------------------------
# merges (mcomb.py)
# ---------------------------------------------------------------------------------------------------------
def image_weighted_merge(a: np.ndarray, b: np.ndarray, w: float) -> np.ndarray:
    """imfilters.py:113-124."""
    if w == 0.0:
        return a
    if w == 1.0:
        return b
    return pil_blend(a, b, w)


def chroma_stabilizer(a: np.ndarray, b: np.ndarray, alpha: float = 0.15, weight: float = 1.0) -> np.ndarray:
    """imfilters.py:160-200: clamp U,V of b into [U_a(1-alpha), U_a(1+alpha)] (bounds truncated to uint8), Y from a."""
    yuv1, yuv2 = cv_rgb2yuv(a), cv_rgb2yuv(b)
    out = yuv2.copy()
    out[..., 0] = yuv1[..., 0]
    for c in (1, 2):
        up = (yuv1[..., c] * (1 + alpha)).clip(0, 255).astype(np.uint8)
        dn = (yuv1[..., c] * (1 - alpha)).clip(0, 255).astype(np.uint8)
        m = np.where(yuv2[..., c] > up, up, yuv2[..., c])      # array_max then array_min (nputils.py:27-79)
        out[..., c] = np.where(m < dn, dn, m)
    rgb = cv_yuv2rgb(out)
    return pil_blend(a, rgb, weight) if weight < 1.0 else rgb


def cv_laplacian(y: np.ndarray) -> np.ndarray:
    """cv2.Laplacian(float32, CV_32F), ksize 1: 4-neighbour stencil, BORDER_REFLECT_101."""
    p = np.pad(y.astype(f32), 1, mode="reflect")
    return (p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:]) - f32(4.0) * p[1:-1, 1:-1]


def chroma_stabilizer_adaptive(a: np.ndarray, b: np.ndarray, base_tol: int = 18, max_extra: int = 22,
                               weight: float = 1.0) -> np.ndarray:
    """imfilters.py:202-269: chroma tolerance base_tol + max_extra * clip(|Laplacian(Y_a)|/255, 0, 1)."""
    yuv1, yuv2 = cv_rgb2yuv(a), cv_rgb2yuv(b)
    y1 = yuv1[..., 0].astype(f32)
    tex = np.clip(np.abs(cv_laplacian(y1)) / f32(255.0), f32(0), f32(1))
    tol = base_tol + max_extra * tex                       # float32 (python ints are weak scalars)
    out = [yuv1[..., 0]]
    for c in (1, 2):
        c1 = yuv1[..., c].astype(np.int16) - 128
        c2 = yuv2[..., c].astype(np.int16) - 128
        lo = np.clip(c1 - tol, -128, 127)
        hi = np.clip(c1 + tol, -128, 127)
        m = np.clip(c2, lo, hi)
        out.append((m + 128).astype(np.uint8))
    rgb = cv_yuv2rgb(np.stack(out, -1))
    return pil_blend(a, rgb, weight) if weight < 1.0 else rgb


def constrained_chroma_merge(a, b, weight=0.5, chroma_threshold=0.2, red_fix_on=True):
    """ConstrainedChromaMerge.merge_frame (mcomb.py:333-367)."""
    stab = chroma_stabilizer(a, b, chroma_threshold, weight)
    return red_fix(stab) if red_fix_on else stab


def chroma_bound_adaptive_merge(a, b, red_fix_on=True, base_tol=14, max_extra=18, weight=0.5):
    """ChromaBoundAdaptiveMerge.merge_frame (mcomb.py:370-437)."""
    stab = chroma_stabilizer_adaptive(a, b, base_tol, max_extra, weight)
    return red_fix(stab) if red_fix_on else stab


def luma_masked_merge(a, b, c, luma_limit=0.4, white_limit=0.7, weight=0.5):
    """LumaMaskedMerge.merge_frame (mcomb.py:238-271); c = a desaturated by luma_mask_sat (a itself when 1.0)."""
    if luma_limit == white_limit:
        masked = image_luma_merge(c, b, luma_limit)
    else:
        masked = w_image_luma_merge(c, b, luma_limit, white_limit)
    return image_weighted_merge(a, masked, weight) if weight < 1.0 else masked


def adaptive_luma_weight(luma: float, luma_limit: float, alpha: float, weight: float, min_w: float) -> float:
    if luma < luma_limit:
        return max(weight * pow(luma / luma_limit, alpha), min_w)
    return weight


def adaptive_luma_merge(a, b, luma_threshold=0.6, alpha=1.0, weight=0.5, min_weight=0.15):
    """AdaptiveLumaMerge.merge_frame (mcomb.py:289-314): the blend weight shrinks with the mean luma of b."""
    w = adaptive_luma_weight(get_image_luma(b), luma_threshold, alpha, weight, min_weight)
    return pil_blend_any(a, b, w)


def gradient_mask(s: np.ndarray, tht: int, alpha: float, algo: int) -> np.ndarray:
    """w_np_gradient_mask (restcolor.py:137-202); returns the integer mask 0..255."""
    if algo == 0:
        lum = s.clip(0, 255)
        steep = 2.0
        grad = np.where(lum < tht, steep * lum / alpha - tht, steep * (lum - tht) * alpha)
        return (255.0 - tht - grad).clip(0, 255).astype(int)
    sf = s.astype(f32)
    tht = int(np.clip(tht, 0, 255))
    if tht == 0:
        return np.zeros_like(s, dtype=np.uint8)
    if algo == 1:
        max_s = min(2 * tht, 200)
        norm = (1.0 - (np.clip(sf, 0, max_s) / max_s)) ** alpha
    else:
        rel = np.clip(sf / tht, 0, 2)
        norm = np.exp(-alpha * rel * np.log(2))
        norm = np.where(sf >= 2 * tht, 0.0, norm)
    return np.clip(norm * 255, 0, 255).astype(np.uint8)


def restore_color_gradient(color: np.ndarray, gray: np.ndarray, sat=1.0, tht=50, weight=0.0, alpha=2.0, algo=0):
    """restcolor.py:98-134: gray pixels (low cv S) of `gray` take the (desaturated) colours of `color`."""
    hsv_c = cv_rgb2hsv(color)
    hsv_g = cv_rgb2hsv(gray)
    if sat != 1.0:
        hsv_c[..., 1] = _scale_u8(hsv_c[..., 1], min(max(sat, 0), 10))
    color_sat = cv_hsv2rgb(hsv_c)
    mask = gradient_mask(hsv_g[..., 1], tht, alpha, algo)
    mask8 = np.asarray(mask).astype(np.uint8)                 # stored into a uint8 image (restcolor.py:118-121)
    mw = mask8.astype(np.float64) / 255
    m = gray * (1 - mw)[..., None] + color_sat * mw[..., None]
    out = m.clip(0, 255).astype(np.uint8)
    if weight > 0:
        out = np_weighted_merge(out, color_sat, weight)
    if weight < 0:
        out = np_weighted_merge(out, gray, -weight)
    return out


def vs_merge(a: np.ndarray, b: np.ndarray, weight: float) -> np.ndarray:
    """std.Merge on 8-bit integer clips (VapourSynth merge.c restated; PARITY UNPINNED — library absent):
    15-bit fixed-point weight, dst = a + (((b - a) * w15 + 2^14) >> 15)."""
    if weight == 0:
        return a
    if weight == 1:
        return b
    w15 = min(max(int(weight * (1 << 15) + 0.5), 0), 1 << 15)
    ai, bi = a.astype(np.int64), b.astype(np.int64)
    return (ai + (((bi - ai) * w15 + (1 << 14)) >> 15)).astype(np.uint8)


def chroma_retention_merge(a, b, sat=0.8, tht=30, weight=0.9, alpha=2.0, mask_weight=0.0, algo=0, chroma_resize=False):
    """ChromaRetentionMerge (mcomb.py:450-516) -> vs_sc_recover_gradient_color (vsfilters.py:366-422) -> vs_simple_merge
    (vsfilters.py:730-739).  chroma_resize=True (mcomb.py:481-512): both clips are squeezed to frame_size x frame_size with
    Spline64 first (frame_size from 0.4 * width, 16 <= rf <= 48; only if that is a downscale), the restore runs there (its
    frame-luma gate too), the result goes back with Spline64 and takes the luma of clip_a (vs_sc_recover_clip_luma)."""
    import math
    from . import pixel_oracle as px
    alpha = max(min(alpha, 10.0), 1.0)                      # DEF_MAX/MIN_COLOR_ALPHA (constants.py)
    H, W = a.shape[:2]
    ca, cb = a, b
    if chroma_resize:
        fs = min(min(max(math.trunc(0.4 * W / 16), 16), 48) * 16, W)
        if fs < W:
            ca, cb = px.resize_plane_u8(a, fs, fs, "spline64"), px.resize_plane_u8(b, fs, fs, "spline64")
        else:
            chroma_resize = False
    w = mask_weight
    luma = get_image_luma(ca, 255)
    if not (DEF_STANDARD_DARK <= luma <= DEF_STANDARD_BRIGHT):
        w = min(w, -0.5)
        alpha = max(alpha, 4.0)
    restored = restore_color_gradient(cb, ca, sat, tht, w, alpha, algo)
    if chroma_resize:
        restored = px.chroma_post_process(px.resize_plane_u8(restored, W, H, "spline64"), a)
    return vs_merge(a, restored, weight)


def combine_models(a, b, method: int, weight: float, cmc_p=(0.15, True, 20, 24), lmm_p=(0.15, 0.65, 1.0),
                   alm_p=(0.8, 1.0, 0.15), crt_p=(0.8, 30, 2, False, 0, 0)):
    """vs_sc_combine_models (mcomb.py:125-192) with sat=(1,1), hue=(0,0) (vs_tweak is the identity)."""
    if method == 2:
        return image_weighted_merge(a, b, weight)
    if method == 3:
        ccm = constrained_chroma_merge(a, b, weight, cmc_p[0], cmc_p[1] if len(cmc_p) > 1 else True)
        m = image_weighted_merge(a, b, min(weight, 0.6))
        return image_weighted_merge(ccm, m, 0.3)
    if method == 4:
        c = a
        if lmm_p[2] < 1:                      # mcomb.py:239-242: clipc = vs_tweak(clipa, sat=luma_mask_sat) (zimg round trip, restated)
            from . import zimg_oracle
            c = zimg_oracle.vs_tweak(a, sat=lmm_p[2])
        return luma_masked_merge(a, b, c, lmm_p[0], lmm_p[1], weight)
    if method == 5:
        return adaptive_luma_merge(a, b, alm_p[0], alm_p[1], weight, alm_p[2])
    if method == 6:
        return chroma_retention_merge(a, b, crt_p[0], crt_p[1], weight, crt_p[2], crt_p[4], crt_p[5], chroma_resize=bool(crt_p[3]))
    if method == 7:
        red, base_tol, max_extra = (cmc_p[1], cmc_p[2], cmc_p[3]) if len(cmc_p) > 1 else (True, 20, 24)
        return chroma_bound_adaptive_merge(a, b, red, base_tol, max_extra, weight)
    raise ValueError("HAVC: only dd_method in (0,6) is supported")


# ---------------------------------------------------------------------------------------------------------
# chroma-adjust filters (row 21)
# ---------------------------------------------------------------------------------------------------------
def adjust_chroma(img: np.ndarray, hue_range: str, sat: float = 0.3, hue: int = 0, weight: float = 0) -> np.ndarray:
    """restcolor.py:239-286: pixels whose hue lies in `hue_range` are replaced by a desaturated / hue-shifted copy."""
    if hue_range in ("none", ""):
        return img
    hsv = cv_rgb2hsv(img)
    g = hsv.copy()
    if hue != 0:
        g[..., 0] = np_hue_add(g[..., 0], hue)
    if sat != 1:
        g[..., 1] = _scale_u8(g[..., 1], min(max(sat, 0), 10))
    gray_rgb = cv_hsv2rgb(g)
    out = mask_select(img, gray_rgb, hue_mask(hsv[..., 0], hue_range))
    if weight > 0:
        out = np_weighted_merge(out, gray_rgb if hue == 0 else img, weight)
    if weight < 0:
        out = np_weighted_merge(out, img, -weight)
    return out


def adjust_hue_range(img: np.ndarray, hue_adjust: str) -> np.ndarray:
    """restcolor.py:221-237 (vs_sc_adjust_clip_hue, vsfilters.py:435-455)."""
    if hue_adjust in ("none", ""):
        return img
    p = parse_hue_adjust(hue_adjust)
    if p is None:
        return img
    return adjust_chroma(img, p[0], p[1], p[2], p[3])


def luma_adjusted_levels(img: np.ndarray, luma_min: float = 0, gamma: float = 1.0, gamma_luma_min: float = 0,
                         gamma_alpha: float = 0, gamma_min: float = 0.2, i_min: int = 0, i_max: int = 255) -> np.ndarray:
    """imfilters.py:335-372 (sc_constrained_tweak, vsfilters.py:656-675)."""
    yuv = cv_rgb2yuv(img)
    y = yuv[..., 0]
    luma = float(np.mean(y)) / 255
    i_alpha = int(255 * (luma_min - luma)) if luma < luma_min else 0
    y_new = y
    if i_alpha > 1:
        y_new = (y.astype(np.int64) + i_alpha).clip(i_min, i_max).astype(np.uint8)
    if gamma != 1 and luma < gamma_luma_min:
        g_new = max(gamma * pow(luma / gamma_luma_min, gamma_alpha), gamma_min) if gamma_alpha != 0 else gamma
        t = np.power(y_new / 255, 1 / g_new)
        y_new = (t * 255).clip(i_min, i_max).astype(np.uint8)
    out = yuv.copy().clip(i_min, i_max)
    out[..., 0] = y_new
    return cv_yuv2rgb(out)


# ---------------------------------------------------------------------------------------------------------
# HAVC_stabilizer per-frame stages: vs_dark_tweak, vs_chroma_bright_tweak, vs_colormap
# ---------------------------------------------------------------------------------------------------------
def np_image_chroma_tweak(img: np.ndarray, sat: float = 1, bright: float = 0, hue: int = 0, hue_adjust: str = "none") -> np.ndarray:
    """restcolor.py:288-342: cv HSV hue add / S scale / V scale, then (optionally) the "chroma adjustment" stage whose
    mask is taken from the hue AFTER the first stage and whose unmasked pixels come from the ORIGINAL image."""
    hsv = cv_rgb2hsv(img)
    hsv[..., 0] = np_hue_add(hsv[..., 0], hue)
    hsv[..., 1] = _scale_u8(hsv[..., 1], min(max(sat, 0), 10))
    hsv[..., 2] = _scale_u8(hsv[..., 2], min(max(1 + bright, 0), 10))
    color = cv_hsv2rgb(hsv)
    if hue_adjust in ("none", ""):
        return color
    p = parse_hue_adjust(hue_adjust)
    if p is None:
        return color
    hue_range, sat2, hue2, weight = p
    g = cv_rgb2hsv(color)
    if hue2 != 0:
        g[..., 0] = np_hue_add(g[..., 0], hue2)
    if sat2 != 1:
        g[..., 1] = _scale_u8(g[..., 1], min(max(sat2, 0), 10))
    gray_rgb = cv_hsv2rgb(g)
    out = mask_select(img, gray_rgb, hue_mask(hsv[..., 0], hue_range))
    if weight > 0:
        out = np_weighted_merge(out, gray_rgb if hue2 == 0 else img, weight)
    if weight < 0:
        out = np_weighted_merge(out, img, -weight)
    return out


def image_chroma_tweak(img, sat=1, bright=0, hue=0, hue_adjust="none"):
    """imfilters.py:540-550."""
    if sat == 1 and bright == 0 and hue == 0 and hue_adjust == "none":
        return img
    return np_image_chroma_tweak(img, sat, bright, hue, hue_adjust)


def _luma_merge(img_dark, img_white, lo, hi):
    return image_luma_merge(img_dark, img_white, lo) if lo == hi else w_image_luma_merge(img_dark, img_white, lo, hi)


def dark_tweak(img: np.ndarray, dark_threshold: float = 0.3, dark_amount: float = 0.8, dark_hue_adjust: str = "none") -> np.ndarray:
    """vs_sc_dark_tweak.merge_frame (vsfilters.py:604-636)."""
    d_threshold = 0.1
    d_white = min(max(dark_threshold, d_threshold), 0.50)
    d_sat = min(max(1.1 - dark_amount, 0.10), 0.80)
    d_bright = -min(max(dark_amount, 0.20), 0.90)
    img2 = image_tweak(img, bright=d_bright, sat=d_sat, hue_range=dark_hue_adjust)
    return _luma_merge(img2, img, d_threshold, d_white)


def chroma_bright_tweak(img: np.ndarray, black_threshold=0.3, white_threshold=0.6, dark_sat=0.8, dark_bright=-0.10,
                        chroma_adjust: str = "none") -> np.ndarray:
    """vs_sc_chroma_bright_tweak.merge_frame (vsfilters.py:525-552)."""
    img2 = image_chroma_tweak(img, bright=dark_bright, sat=dark_sat, hue_adjust=chroma_adjust)
    return _luma_merge(img2, img, black_threshold, white_threshold)


def colormap(img: np.ndarray, colormap_adjust: str) -> np.ndarray:
    """_vs_sc_colormap.merge_frame (vsfilters.py:577-590)."""
    return image_chroma_tweak(img, hue_adjust=colormap_adjust)


def stabilizer_stages(img: np.ndarray, dark=False, dark_p=(0.2, 0.8), smooth=False, smooth_p=(0.3, 0.7, 0.9, 0.0, "none"),
                      colormap_adjust: str = "none") -> np.ndarray:
    """The per-frame stages of HAVC_stabilizer on the squeezed frame (vsdeoldify/__init__.py:2823-2861), stab=False."""
    out = img
    if dark:
        out = dark_tweak(out, dark_p[0], dark_p[1], (dark_p[2] if len(dark_p) > 2 else "none").lower())
    if smooth:
        out = chroma_bright_tweak(out, smooth_p[0], smooth_p[1], smooth_p[2], -smooth_p[3],
                                  (smooth_p[4] if len(smooth_p) > 4 else "none").lower())
    if colormap_adjust not in ("none", ""):
        out = colormap(out, colormap_adjust)
    return out
