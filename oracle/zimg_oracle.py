"""TEST INFRASTRUCTURE (parity oracle) — restatement of the zimg / VapourSynth steps the reference's `vs_tweak` chains.

PARITY UNPINNED: VapourSynth / zimg are not installable here, the reference has no tests or golden vectors for
this path, so nothing can confirm the chroma siting, the filter normalisation or the dither order against the real library.
What is restated, from the published zimg algorithm (zimg 3.x: graph/graphbuilder.cpp, resize/filter.cpp, colorspace/matrix3.cpp,
depth/dither.cpp) and the VapourSynth std.Expr / std.Lut definitions:

  vs_tweak (vsdeoldify/vsslib/vsfilters.py:753-850):
    RGB24 --resize.Bicubic(format=YUV420P8, matrix_s="709", range_s="full")-->  YUV420P8
        u8 -> float32 (x / 255); BT.709 matrix (Kr 0.2126, Kb 0.0722); chroma planes 2:1 with Bicubic (b = c = 1/3), chroma location
        "left" (MPEG-2: co-sited horizontally, centred vertically), vertical pass first; float -> u8 round-half-even, chroma +128
    --std.Expr on (U, V): hue rotation / saturation gain around 128 in float32, clamp [0, 255], round-half-even-->
    --std.Lut on Y: int((i - min) * cont + bright + min + 0.5) clamped-->
    --resize.Bicubic(format=RGB24, matrix_in_s="709", range_s="full", dither_type="error_diffusion")-->  RGB24
        chroma 1:2 with Bicubic (horizontal pass first), inverse matrix, x * 255, Floyd-Steinberg error diffusion per plane
        (7/16 left, 3/16 upper-right, 5/16 upper, 1/16 upper-left; clamp, lrint), in raster order.

All arithmetic is float32 in the stated order so that the GPU kernels (csrc/zimg.cu) can reproduce it bit for bit.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32
KR, KB = 0.2126, 0.0722
KG = 1.0 - KR - KB
MATRICES = {"709": (0.2126, 0.0722), "601": (0.299, 0.114)}      # Kr, Kb (BT.709; BT.601 = VapourSynth's 470bg / 170m)


def quant(limited: bool):
    """float -> 8-bit integer scaling of zimg's depth conversion: (y_scale, y_off, c_scale, c_off)."""
    return (F(219.0), F(16.0), F(224.0), F(128.0)) if limited else (F(255.0), F(0.0), F(255.0), F(128.0))


def _bicubic(x: float, b: float = 1.0 / 3.0, c: float = 1.0 / 3.0) -> float:
    """Mitchell-Netravali family (zimg BicubicFilter), support 2."""
    x = abs(x)
    p0 = (6.0 - 2.0 * b) / 6.0
    p2 = (-18.0 + 12.0 * b + 6.0 * c) / 6.0
    p3 = (12.0 - 9.0 * b - 6.0 * c) / 6.0
    q0 = (8.0 * b + 24.0 * c) / 6.0
    q1 = (-12.0 * b - 48.0 * c) / 6.0
    q2 = (6.0 * b + 30.0 * c) / 6.0
    q3 = (-b - 6.0 * c) / 6.0
    if x < 1.0:
        return p0 + x * x * (p2 + x * p3)
    if x < 2.0:
        return q0 + x * (q1 + x * (q2 + x * q3))
    return 0.0


def filter_bank(src: int, dst: int, shift: float = 0.0):
    """zimg compute_filter for Bicubic: (start int32 [dst], weights float32 [dst, T]) with out[i] = sum_t w[i,t] * in[start[i]+t];
    support widened by the shrink ratio, weights normalised per output sample, out-of-range taps mirrored back."""
    support = 2
    scale = dst / src
    step = min(scale, 1.0)
    fsize = max(int(math.ceil(support / step)) * 2, 1)
    m = np.zeros((dst, src), np.float64)
    for i in range(dst):
        pos = (i + 0.5) / scale + shift
        begin = math.floor(pos - fsize / 2.0 + 0.5) + 0.5
        ws = [_bicubic((begin + j - pos) * step) for j in range(fsize)]
        total = sum(ws)
        for j in range(fsize):
            xp = begin + j
            real = -xp if xp < 0.0 else (2.0 * src - xp if xp >= src else xp)
            m[i, min(max(int(math.floor(real)), 0), src - 1)] += ws[j] / total
    nz = m != 0.0
    first = nz.argmax(1)
    last = src - 1 - nz[:, ::-1].argmax(1)
    T = int((last - first).max()) + 1
    start = np.clip(np.minimum(first, src - T), 0, None).astype(np.int32)
    w = np.zeros((dst, T), np.float32)
    for o in range(dst):
        seg = m[o, start[o]:start[o] + T]
        w[o, :len(seg)] = seg.astype(np.float32)
    return start, w


def chroma_down_tables(W: int, H: int):
    """4:4:4 -> 4:2:0, chroma location 'left': output sample i sits on luma sample 2i horizontally (shift -0.5 source pixels from
    the centred position), centred vertically."""
    return filter_bank(W, W // 2, -0.5), filter_bank(H, H // 2, 0.0)


def chroma_up_tables(W: int, H: int):
    """4:2:0 -> 4:4:4: the inverse siting (+0.25 chroma samples horizontally)."""
    return filter_bank(W // 2, W, 0.25), filter_bank(H // 2, H, 0.0)


def _resample(x: np.ndarray, axis: int, tab) -> np.ndarray:
    """float32 separable pass, taps accumulated in ascending order with separate multiply and add."""
    start, w = tab
    xm = np.moveaxis(x, axis, -1)
    out = np.zeros(xm.shape[:-1] + (len(start),), F)
    n = xm.shape[-1]
    for t in range(w.shape[1]):
        idx = np.minimum(start + t, n - 1)
        out = (out + xm[..., idx] * w[:, t]).astype(F)
    return np.moveaxis(out, -1, axis)


def _rint_u8(x: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def forward_matrix(matrix: str = "709"):
    """RGB -> YCbCr in float32, rows Y, Cb, Cr (zimg ncl_rgb_to_yuv_matrix), and its inverse."""
    kr, kb = MATRICES[matrix]
    kg = 1.0 - kr - kb
    m = np.array([[kr, kg, kb],
                  [-kr / (2 * (1 - kb)), -kg / (2 * (1 - kb)), 0.5],
                  [0.5, -kg / (2 * (1 - kr)), -kb / (2 * (1 - kr))]], np.float64)
    # inverse in closed form (what the numerical 3x3 inverse zimg takes in double rounds to in float32)
    inv = np.array([[1.0, 0.0, 2 * (1 - kr)],
                    [1.0, -2 * kb * (1 - kb) / kg, -2 * kr * (1 - kr) / kg],
                    [1.0, 2 * (1 - kb), 0.0]], np.float64)
    return m.astype(F), inv.astype(F)


def rgb24_to_yuv420p8(rgb: np.ndarray, matrix: str = "709", limited: bool = False, dither: bool = False):
    """uint8 [H,W,3] -> (Y u8 [H,W], U u8 [H/2,W/2], V u8 [H/2,W/2]).  vs_tweak: BT.709, full range, no dither; restore_format
    (havc_utils.py:199-222): the clip's matrix / range with error-diffusion dither on all three planes."""
    H, W = rgb.shape[:2]
    assert H % 2 == 0 and W % 2 == 0
    fwd, _ = forward_matrix(matrix)
    ys, yo, cs, co = quant(limited)
    c = rgb.astype(F) * F(1.0 / 255.0)
    r, g, b = c[..., 0], c[..., 1], c[..., 2]
    planes = [((fwd[k, 0] * r).astype(F) + (fwd[k, 1] * g).astype(F)).astype(F) + (fwd[k, 2] * b).astype(F) for k in range(3)]
    q = (lambda x: error_diffusion_u8(x)) if dither else _rint_u8
    out = [q(((planes[0] * ys).astype(F) + yo).astype(F))]
    th, tv = chroma_down_tables(W, H)
    for k in (1, 2):
        ch = _resample(_resample(planes[k].astype(F), 0, tv), 1, th)           # vertical pass first (zimg orders by cost)
        out.append(q(((ch * cs).astype(F) + co).astype(F)))
    return tuple(out)


def error_diffusion_u8(x: np.ndarray) -> np.ndarray:
    """zimg dither_ed on one float32 plane already scaled to [0, 255]: raster order, Floyd-Steinberg weights."""
    H, W = x.shape
    out = np.zeros((H, W), np.uint8)
    top = np.zeros(W + 2, F)
    for yy in range(H):
        cur = np.zeros(W + 2, F)
        row = x[yy]
        for j in range(W):
            je = j + 1
            err = F(0.0)
            err = F(err + F(cur[je - 1] * F(7.0 / 16.0)))
            err = F(err + F(top[je + 1] * F(3.0 / 16.0)))
            err = F(err + F(top[je] * F(5.0 / 16.0)))
            err = F(err + F(top[je - 1] * F(1.0 / 16.0)))
            v = F(row[j] + err)
            v = min(max(v, F(0.0)), F(255.0))
            q = int(np.rint(v))
            out[yy, j] = q
            cur[je] = F(v - F(q))
        top = cur
    return out


def yuv420p8_to_rgb24(y: np.ndarray, u: np.ndarray, v: np.ndarray, dither: bool = True, matrix: str = "709",
                      limited: bool = False) -> np.ndarray:
    """(Y [H,W], U, V [H/2,W/2]) u8 -> uint8 [H,W,3]: Bicubic chroma up-sampling, inverse matrix, error diffusion."""
    H, W = y.shape
    _, inv = forward_matrix(matrix)
    ys, yo, cs, co = quant(limited)
    th, tv = chroma_up_tables(W, H)
    yf = ((y.astype(F) - yo).astype(F) * F(F(1.0) / ys)).astype(F)
    ch = []
    for p in (u, v):
        c = ((p.astype(F) - co).astype(F) * F(F(1.0) / cs)).astype(F)
        ch.append(_resample(_resample(c.astype(F), 1, th), 0, tv))               # horizontal pass first when up-sampling
    planes = [((inv[k, 0] * yf).astype(F) + (inv[k, 1] * ch[0]).astype(F)).astype(F) + (inv[k, 2] * ch[1]).astype(F) for k in range(3)]
    outs = []
    for k in range(3):
        s = (planes[k].astype(F) * F(255.0)).astype(F)
        outs.append(error_diffusion_u8(s) if dither else _rint_u8(s))
    return np.stack(outs, -1)


def gray8_to_rgb24(y: np.ndarray, limited: bool = True) -> np.ndarray:
    """resize.Bicubic(format=RGB24, range_in_s="limited", range_s="full") of a GRAY8 clip (havc_utils.py:145-151): no dither."""
    ys, yo, _, _ = quant(limited)
    v = (((y.astype(F) - yo).astype(F) * F(F(1.0) / ys)).astype(F) * F(255.0)).astype(F)
    return np.repeat(_rint_u8(v)[..., None], 3, -1)


def vs_tweak(rgb: np.ndarray, hue: float = 0.0, sat: float = 1.0, bright: float = 0.0, cont: float = 1.0) -> np.ndarray:
    """vs_tweak (vsfilters.py:753-850) with gamma = 1, coring = False on one uint8 [H,W,3] frame."""
    if hue == 0 and sat == 1 and bright == 0 and cont == 1:
        return rgb
    y, u, v = rgb24_to_yuv420p8(rgb)
    if -1.0 < bright < 1.0:
        bright = bright * 255.0
    if hue != 0 or sat != 1:
        h = hue * math.pi / 180.0
        c1, c2 = F(math.cos(h) * sat), F(math.sin(h) * sat)
        uf, vf = u.astype(F) - F(128.0), v.astype(F) - F(128.0)
        nu = ((uf * c1).astype(F) + (vf * c2).astype(F)).astype(F) + F(128.0)     # x 128 - c1 * y 128 - c2 * + 128 +
        nv = ((vf * c1).astype(F) - (uf * c2).astype(F)).astype(F) + F(128.0)     # y 128 - c1 * x 128 - c2 * - 128 +
        u = _rint_u8(np.minimum(np.maximum(nu, F(0.0)), F(255.0)))
        v = _rint_u8(np.minimum(np.maximum(nv, F(0.0)), F(255.0)))
    if bright != 0 or cont != 1:
        lut = np.array([min(max(int(i * cont + bright + 0.5), 0), 255) for i in range(256)], np.uint8)
        y = lut[y]
    return yuv420p8_to_rgb24(y, u, v)
