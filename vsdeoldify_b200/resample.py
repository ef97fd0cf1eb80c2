"""Host-side resampling tables for the table-driven GPU resize kernels (havc_resample_h & co).

zimg (VapourSynth `resize.Spline64/Spline36`, used by HAVC_colorizer at vsdeoldify/__init__.py:2504 and
:3547 and by vsresize.py) is not available in this environment, so the filter bank follows the published
zimg / Avisynth definitions: separable, half-pixel centres, support widened by the
shrink ratio, per-output normalised weights, taps that fall outside the image mirrored back in.
"""
from __future__ import annotations

import math
from typing import Callable, Tuple

import numpy as np


def spline64(x: float) -> float:
    x = abs(x)
    if x < 1.0:
        return ((49.0 / 41.0 * x - 6387.0 / 2911.0) * x - 3.0 / 2911.0) * x + 1.0
    if x < 2.0:
        t = x - 1.0
        return ((-24.0 / 41.0 * t + 4032.0 / 2911.0) * t - 2328.0 / 2911.0) * t
    if x < 3.0:
        t = x - 2.0
        return ((6.0 / 41.0 * t - 1008.0 / 2911.0) * t + 582.0 / 2911.0) * t
    if x < 4.0:
        t = x - 3.0
        return ((-1.0 / 41.0 * t + 168.0 / 2911.0) * t - 97.0 / 2911.0) * t
    return 0.0


def spline36(x: float) -> float:
    x = abs(x)
    if x < 1.0:
        return ((13.0 / 11.0 * x - 453.0 / 209.0) * x - 3.0 / 209.0) * x + 1.0
    if x < 2.0:
        t = x - 1.0
        return ((-6.0 / 11.0 * t + 270.0 / 209.0) * t - 156.0 / 209.0) * t
    if x < 3.0:
        t = x - 2.0
        return ((1.0 / 11.0 * t - 45.0 / 209.0) * t + 26.0 / 209.0) * t
    return 0.0


def bicubic(x: float, b: float = 1.0 / 3.0, c: float = 1.0 / 3.0) -> float:
    """zimg's Bicubic (Mitchell-Netravali family; VapourSynth resize.Bicubic defaults b = c = 1/3), support 2."""
    x = abs(x)
    if x < 1.0:
        return (6.0 - 2.0 * b) / 6.0 + x * x * ((-18.0 + 12.0 * b + 6.0 * c) / 6.0 + x * (12.0 - 9.0 * b - 6.0 * c) / 6.0)
    if x < 2.0:
        return (8.0 * b + 24.0 * c) / 6.0 + x * ((-12.0 * b - 48.0 * c) / 6.0 + x * ((6.0 * b + 30.0 * c) / 6.0 + x * (-b - 6.0 * c) / 6.0))
    return 0.0


KERNELS = {"spline64": (spline64, 4), "spline36": (spline36, 3), "bicubic": (bicubic, 2)}


def filter_matrix(src: int, dst: int, kernel: str = "spline64", shift: float = 0.0) -> np.ndarray:
    """Dense [dst, src] float64 resampling matrix (zimg compute_filter semantics); `shift` moves the sampling positions by that
    many SOURCE samples (chroma siting of the 4:2:0 conversions)."""
    f, support = KERNELS[kernel]
    scale = dst / src
    step = min(scale, 1.0)
    fsupport = support / step
    fsize = max(int(math.ceil(fsupport)) * 2, 1)
    m = np.zeros((dst, src), dtype=np.float64)
    for i in range(dst):
        pos = (i + 0.5) / scale + shift
        begin = math.floor(pos - fsize / 2.0 + 0.5) + 0.5
        ws = [f((begin + j - pos) * step) for j in range(fsize)]
        total = sum(ws)
        for j in range(fsize):
            xpos = begin + j
            if xpos < 0.0:
                real = -xpos
            elif xpos >= src:
                real = 2.0 * src - xpos
            else:
                real = xpos
            idx = min(max(int(math.floor(real)), 0), src - 1)
            m[i, idx] += ws[j] / total
    return m


def build_tables(src: int, dst: int, kernel: str = "spline64", shift: float = 0.0) -> Tuple[np.ndarray, np.ndarray]:
    """(start int32 [dst], weights float32 [dst, T]) with out[o] = sum_t weights[o,t] * in[start[o]+t]."""
    m = filter_matrix(src, dst, kernel, shift)
    nz = m != 0.0
    first = np.where(nz.any(1), nz.argmax(1), 0)
    last = np.where(nz.any(1), src - 1 - nz[:, ::-1].argmax(1), 0)
    T = int((last - first).max()) + 1
    start = np.minimum(first, src - T).astype(np.int32)
    start = np.maximum(start, 0)
    w = np.zeros((dst, T), dtype=np.float32)
    for o in range(dst):
        seg = m[o, start[o]:start[o] + T]
        w[o, :len(seg)] = seg
    return start, w


# ---- phase-periodic plans for the horizontal passes (csrc/pixel.cu: resample_h_periodic_kernel, post_horizontal_periodic_kernel) ----
PERIODIC_DOWN_TAPS = {4: 35, 5: 43, 6: 51}          # padded taps the squeeze kernel is instantiated for, per ratio
PERIODIC_UP_TAPS = (8, 9, 10)


def _same_bits(a: np.ndarray, b: np.ndarray) -> bool:
    return np.array_equal(a.view(np.int32), b.view(np.int32))


def _interior(ok: np.ndarray, mid: int):
    """The maximal run of True around `mid`: [lo, hi)."""
    if not ok[mid]:
        return mid, mid
    lo = mid
    while lo > 0 and ok[lo - 1]:
        lo -= 1
    hi = mid + 1
    while hi < len(ok) and ok[hi]:
        hi += 1
    return lo, hi


def periodic_plan_down(start: np.ndarray, w: np.ndarray, src: int, dst: int):
    """Plan of the phase-periodic squeeze (src = ratio * dst): dict(ratio, taps, offset, lo, hi, w) or None.  Interior output
    columns must satisfy start[o] == ratio * o + c and carry bit-identical weights; the kernel handles blocks of 8 columns."""
    if dst <= 0 or src % dst or dst % 8 or src % 4:
        return None
    R, T = src // dst, int(w.shape[1])
    if R not in PERIODIC_DOWN_TAPS:
        return None
    w = np.ascontiguousarray(w, np.float32)
    mid = dst // 2
    c, wref = int(start[mid]) - R * mid, w[mid]
    ok = np.array([int(start[o]) == R * o + c and _same_bits(w[o], wref) for o in range(dst)])
    lo, hi = _interior(ok, mid)
    sh = c % 4
    offset, taps = c - sh, sh + T
    lo8, hi8 = -(-lo // 8), hi // 8
    while lo8 < hi8 and offset + R * 8 * lo8 < 0:
        lo8 += 1
    span = 4 * ((R * 7 + PERIODIC_DOWN_TAPS[R] + 3) // 4)
    while hi8 > lo8 and offset + R * 8 * (hi8 - 1) + span > src + 64:
        hi8 -= 1
    if taps > PERIODIC_DOWN_TAPS[R] or hi8 - lo8 < max(1, dst // 16):
        return None
    pw = np.zeros(72, np.float32)
    pw[sh:sh + T] = wref
    return dict(ratio=R, taps=taps, offset=offset, lo=lo8, hi=hi8, w=pw)


def periodic_plan_up(start: np.ndarray, w: np.ndarray, src: int, dst: int):
    """Plan of the phase-periodic way back (dst = ratio * src): output ratio * i + p has the window start i + c_p and the weights
    of phase p for every interior input position i; the kernel handles units of 4 positions."""
    if src <= 0 or dst % src or src % 4:
        return None
    R, T = dst // src, int(w.shape[1])
    if R not in (4, 5, 6):
        return None
    w = np.ascontiguousarray(w, np.float32)
    mid = src // 2
    cp = [int(start[R * mid + p]) - mid for p in range(R)]
    ok = np.array([all(int(start[R * i + p]) == i + cp[p] and _same_bits(w[R * i + p], w[R * mid + p]) for p in range(R)) for i in range(src)])
    lo, hi = _interior(ok, mid)
    # common window of the phases = the range of their NON-ZERO taps (a phase that samples exactly on a source pixel has the
    # single weight 1 at its centre; zero taps are exact no-ops of the fmaf chain, so re-basing a window does not change a bit)
    lo_nz, hi_nz = [], []
    for p in range(R):
        nz = np.nonzero(w[R * mid + p])[0]
        if len(nz) == 0:
            return None
        lo_nz.append(cp[p] + int(nz[0])), hi_nz.append(cp[p] + int(nz[-1]))
    cmin = min(lo_nz)
    taps = max(max(hi_nz) - cmin + 1, PERIODIC_UP_TAPS[0])
    ulo, uhi = -(-lo // 4), hi // 4
    while ulo < uhi and 4 * ulo + cmin < 0:
        ulo += 1
    while uhi > ulo and 4 * (uhi - 1) + 3 + cmin + taps > src:
        uhi -= 1
    if taps not in PERIODIC_UP_TAPS or not 0 <= -cmin <= 8 or uhi - ulo < max(1, src // 8) or R * taps > 72:
        return None
    pw = np.zeros(72, np.float32)
    for p in range(R):
        for t in range(T):
            k = cp[p] + t - cmin
            if w[R * mid + p, t] != 0.0:
                pw[p * taps + k] = w[R * mid + p, t]
    return dict(ratio=R, taps=taps, offset=-cmin, lo=ulo, hi=uhi, w=pw)


# ---- Pillow Image.resize tables (ImagingResample, libImaging/Resample.c: precompute_coeffs + normalize_coeffs_8bpc) ----
# BILINEAR: BaseFilter._scale_to_square / _unsquare (vsdeoldify/deoldify/filters.py:37-41,70-73); BICUBIC (a = -0.5):
# colorizers/util.py:21-22.  The support is widened by the shrink ratio, weights are normalised in float64 and then
# rounded to 22-bit fixed point; the GPU pass (havc_pil_resample_u8) accumulates them in integers like Pillow does.
PIL_PRECISION_BITS = 32 - 8 - 2


def _pil_bilinear(x: float) -> float:
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def _pil_bicubic(x: float, a: float = -0.5) -> float:
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


PIL_FILTERS = {"bilinear": (_pil_bilinear, 1.0), "bicubic": (_pil_bicubic, 2.0)}


def pil_tables(in_size: int, out_size: int, filt: str = "bilinear") -> Tuple[np.ndarray, np.ndarray]:
    """(bounds int32 [out, 2] = (first tap, tap count), coeffs int32 [out, ksize]) for one axis."""
    f, support = PIL_FILTERS[filt]
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    bounds = np.zeros((out_size, 2), np.int32)
    coeffs = np.zeros((out_size, ksize), np.int32)
    for o in range(out_size):
        center = (o + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        n = hi - lo
        w = np.array([f((j + lo - center + 0.5) * inv) for j in range(n)], np.float64)
        tot = w.sum()
        if tot != 0.0:
            w = w / tot
        fixed = w * (1 << PIL_PRECISION_BITS)
        coeffs[o, :n] = np.where(fixed < 0, np.trunc(fixed - 0.5), np.trunc(fixed + 0.5)).astype(np.int32)
        bounds[o] = (lo, n)
    return bounds, coeffs


def chroma420_tables(width: int, height: int):
    """Bicubic tables of zimg's 4:4:4 <-> 4:2:0 chroma resampling with chroma location 'left' (MPEG-2: co-sited horizontally with
    the even luma column, centred vertically): (down_h, down_v, up_h, up_v), each (start, weights).  Down: the chroma sample i sits
    at luma column 2i, i.e. half a source sample left of the centred position; up: the inverse siting, +0.25 chroma samples."""
    return (build_tables(width, width // 2, "bicubic", -0.5), build_tables(height, height // 2, "bicubic", 0.0),
            build_tables(width // 2, width, "bicubic", 0.25), build_tables(height // 2, height, "bicubic", 0.0))
