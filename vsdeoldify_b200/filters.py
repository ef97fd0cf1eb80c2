"""Host side of the vsslib model merges and chroma-adjust filters.

Mirrors the reference's dispatcher `vs_sc_combine_models` (vsdeoldify/vsslib/mcomb.py:125-192) and the helpers
around it: it parses the parameter lists and the hue-range mini language (restcolor.py:378-470), prepares the few
host-side tables (gradient masks, restcolor.py:137-202) and sequences libhavc_b200 launches on planar RGB24 device
batches `[B, 3, H, W]`.  Every frame-global quantity (mean luma) stays on the device, so a merge is graph-capturable.
No pixel arithmetic happens here.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import math

import numpy as np
import torch

from . import _lib
from .constants import DEF_ALM_p, DEF_CMC_p, DEF_CRT_p, DEF_LMM_p

# per-row vector block of OpenCV's 8-bit HSV->RGB on the AVX2 build the reference was pinned against (its SIMD
# body truncates, the row tail rounds — see include/havc_b200.h)
CV_SIMD_WIDTH = 32
DEF_MAX_COLOR_ALPHA, DEF_MIN_COLOR_ALPHA = 10.0, 1.0          # vsslib/constants.py

_HUE_NAMES = {"red": (0, 30), "orange": (30, 60), "yellow": (60, 90), "yellow-green": (90, 120), "green": (120, 150),
              "blue-green": (150, 180), "cyan": (180, 210), "blue": (210, 240), "blue-violet": (240, 270),
              "violet": (270, 300), "red-violet": (300, 330), "rose": (330, 360)}


class FilterError(ValueError):
    """Raised for parameter combinations the reference rejects (or that are not built); havc.py turns it into vs.Error."""


def parse_hue_range(hue_range: str) -> List[Tuple[float, float]]:
    """_parse_hue_range per comma-separated item (restcolor.py:430-470): a colour name or 'min:max' in degrees."""
    out = []
    for part in hue_range.split(","):
        if part in _HUE_NAMES:
            lo, hi = _HUE_NAMES[part]
        else:
            p = part.split(":")
            if len(p) == 2 and p[0].isnumeric() and p[1].isnumeric():
                lo, hi = float(p[0]), float(p[1])
            else:
                raise FilterError("HybridAVC: unknown hue name: " + part)
        out.append((float(lo), float(hi)))
    return out


def parse_hue_adjust(hue_adjust: str):
    """_parse_hue_adjust (restcolor.py:378-410): 'ranges|sat_or_+-hue,weight' -> (ranges, sat, hue, weight) or None."""
    p = hue_adjust.split("|")
    sat, hue, weight = 1.0, 0, 0.0
    if len(p) < 1 or len(p) > 2:
        return None
    if len(p) == 1:
        return p[0], sat, hue, weight

    def isfloat(x):
        try:
            float(x)
            return True
        except ValueError:
            return False
    sw = p[1].split(",")
    if len(sw) != 2 or not isfloat(sw[0]) or not isfloat(sw[1]):
        return None
    if sw[0][0] in ("-", "+"):
        hue = int(sw[0])
    else:
        sat = float(sw[0])
    if sat > 10:
        hue, sat = int(sat), 1.0
    return p[0], sat, hue, float(sw[1])


def hue_ranges_struct(hue_range: Optional[str]) -> Optional[_lib.HueRanges]:
    if hue_range in (None, "", "none"):
        return None
    rs = parse_hue_range(hue_range)
    if len(rs) > 8:
        raise FilterError("at most 8 hue ranges are supported")
    h = _lib.HueRanges()
    h.n = len(rs)
    for i, (lo, hi) in enumerate(rs):
        h.lo_deg[i], h.hi_deg[i] = lo, hi
    return h


def gradient_mask_lut(tht: int, alpha: float, algo: int) -> np.ndarray:
    """w_np_gradient_mask (restcolor.py:137-202) tabulated over the 256 possible saturation values; the table is what
    the reference stores into its uint8 mask image."""
    s = np.arange(256, dtype=np.uint8)
    if algo == 0:
        lum = s.clip(0, 255)
        steep = 2.0
        grad = np.where(lum < tht, steep * lum / alpha - tht, steep * (lum - tht) * alpha)
        return (255.0 - tht - grad).clip(0, 255).astype(int).astype(np.uint8)
    sf = s.astype(np.float32)
    tht = int(np.clip(tht, 0, 255))
    if tht == 0:
        return np.zeros(256, np.uint8)
    if algo == 1:
        max_s = min(2 * tht, 200)
        norm = (1.0 - (np.clip(sf, 0, max_s) / max_s)) ** alpha
    else:
        rel = np.clip(sf / tht, 0, 2)
        norm = np.exp(-alpha * rel * np.log(2))
        norm = np.where(sf >= 2 * tht, 0.0, norm)
    return np.clip(norm * 255, 0, 255).astype(np.uint8)


class SquareSqueeze:
    """clip.resize.Spline64(width=S, height=S) of planar u8 batches [B,3,H,W] and the way back with the luma transplant
    (_clip_chroma_resize, vsdeoldify/__init__.py:3545-3554 = Spline64 to W x H + vs_recover_clip_luma): the table-driven
    separable passes of the frame pre / post pipeline, reused by ChromaRetentionMerge(chroma_resize=True) (mcomb.py:481-512)
    and HAVC_merge(clip_luma=...) (vsdeoldify/__init__.py:2660-2673)."""

    def __init__(self, B: int, H: int, W: int, S: int, device, kernel: str = "spline64", out_hw: Optional[Tuple[int, int]] = None):
        from .engine import _Tables
        self.lib, self.B, self.H, self.W, self.S = _lib.lib(), B, H, W, S
        self.dev = torch.device(device)
        self.OH, self.OW = out_hw if out_hw is not None else (H, W)          # size of the way back (the luma clip's)
        self.t_down_h, self.t_down_v = _Tables(W, S, kernel, self.dev), _Tables(H, S, kernel, self.dev)
        self.t_up_h, self.t_up_v = _Tables(S, self.OW, kernel, self.dev), _Tables(S, self.OH, kernel, self.dev)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.tmp_down = torch.empty(B, 3, H, S, **f32)
        self.tmp_up = torch.empty(B, 3, self.OH, S, **f32)
        self.x_scratch = torch.empty(B, S, S, 8, dtype=torch.float16, device=self.dev)     # pre_vertical's network-input by-product

    def down(self, src, dst, stream: int = 0):
        """src u8 [B,3,H,W] -> dst u8 [B,3,S,S]."""
        lib, B, H, W, S, chk = self.lib, self.B, self.H, self.W, self.S, _lib.check
        td, tv = self.t_down_h, self.t_down_v
        td.resample_h(lib, src.data_ptr(), self.tmp_down.data_ptr(), B * 3 * H, stream, "squeeze.h")
        chk(lib.havc_pre_vertical(self.tmp_down.data_ptr(), dst.data_ptr(), self.x_scratch.data_ptr(), B, H, S, tv.start.data_ptr(),
                                  tv.w.data_ptr(), tv.taps, 0, stream), "squeeze.v")

    def up(self, src, luma, dst, stream: int = 0, transplant: bool = True):
        """src u8 [B,3,S,S] -> dst u8 [B,3,OH,OW], keeping the luma of `luma` (same size as dst) when transplant."""
        lib, B, S, chk = self.lib, self.B, self.S, _lib.check
        uh, uv = self.t_up_h, self.t_up_v
        chk(lib.havc_resample_v(src.data_ptr(), self.tmp_up.data_ptr(), B * 3, S, self.OH, S, uv.start.data_ptr(), uv.w.data_ptr(),
                                uv.taps, stream), "unsqueeze.v")
        uh.post_horizontal(lib, self.tmp_up.data_ptr(), luma.data_ptr() if luma is not None else None, dst.data_ptr(), B, self.OH,
                           int(transplant), stream, "unsqueeze.h")


class RectResize:
    """clip.resize.Spline36 / Spline64(width, height) of planar u8 batches [B,3,H,W] -> [B,3,OH,OW] (zimg restated: two separable
    float32 passes, the cheaper one first, one rounding at the end - oracle/pixel_oracle.py:resize_plane_u8): the resize of
    resize_min_HW / resize_to_chroma (vsslib/vsresize.py:30-127)."""

    def __init__(self, B: int, H: int, W: int, OH: int, OW: int, device, kernel: str = "spline36"):
        from .engine import _Tables
        self.lib, self.B, self.H, self.W, self.OH, self.OW = _lib.lib(), B, H, W, OH, OW
        self.dev = torch.device(device)
        self.h_first = OW * H <= OH * W                     # size of the intermediate image = cost of the second pass
        self.t_h, self.t_v = _Tables(W, OW, kernel, self.dev), _Tables(H, OH, kernel, self.dev)
        shape = (B, 3, H, OW) if self.h_first else (B, 3, OH, W)
        self.tmp = torch.empty(*shape, dtype=torch.float32, device=self.dev)

    def run(self, src, dst, stream: int = 0):
        lib, B, H, W, OH, OW, chk = self.lib, self.B, self.H, self.W, self.OH, self.OW, _lib.check
        th, tv = self.t_h, self.t_v
        if self.h_first:
            th.resample_h(lib, src.data_ptr(), self.tmp.data_ptr(), B * 3 * H, stream, "resize.h")
            chk(lib.havc_resample_v_f32_u8(self.tmp.data_ptr(), dst.data_ptr(), B * 3, H, OH, OW, tv.start.data_ptr(), tv.w.data_ptr(),
                                           tv.taps, stream), "resize.v")
        else:
            chk(lib.havc_resample_v(src.data_ptr(), self.tmp.data_ptr(), B * 3, H, OH, W, tv.start.data_ptr(), tv.w.data_ptr(), tv.taps,
                                    stream), "resize.v")
            th.post_horizontal(lib, self.tmp.data_ptr(), None, dst.data_ptr(), B, OH, 0, stream, "resize.h")


class FilterBank:
    """Scratch buffers + launch sequencing for one (B, H, W) on one device."""

    def __init__(self, B: int, H: int, W: int, device, simd_width: int = CV_SIMD_WIDTH):
        self.B, self.H, self.W, self.dev = B, H, W, torch.device(device)
        self.simd = simd_width
        self.lib = _lib.lib()
        u8 = dict(dtype=torch.uint8, device=self.dev)
        self.tmp = [torch.empty(B, 3, H, W, **u8) for _ in range(3)]
        self.stats = [torch.zeros(B, 2, dtype=torch.int64, device=self.dev) for _ in range(2)]
        self._luts = {}
        self.keep = []

    # ---- helpers ---------------------------------------------------------------------------------------
    def _chk(self, t: torch.Tensor):
        assert t.dtype == torch.uint8 and t.is_contiguous() and tuple(t.shape) == (self.B, 3, self.H, self.W) and t.device == self.dev

    def _lut(self, tht, alpha, algo) -> torch.Tensor:
        key = (int(tht), float(alpha), int(algo))
        if key not in self._luts:
            self._luts[key] = torch.from_numpy(gradient_mask_lut(*key)).to(self.dev)
        return self._luts[key]

    def _dims(self):
        return self.B, self.H, self.W

    def frame_stats(self, img, stats, stream=0, bright: float = 1.0):
        _lib.check(self.lib.havc_frame_stats(img.data_ptr(), *self._dims(), bright, stats.data_ptr(), stream), "frame_stats")

    def select_frames(self, dst, src, skip, stream=0):
        """Scene-change gate: frames with skip[b] != 0 become the corresponding frame of `src`."""
        _lib.check(self.lib.havc_select_frames(dst.data_ptr(), src.data_ptr(), skip.data_ptr(), self.B, 3 * self.H * self.W, stream),
                   "select_frames")

    def blend(self, a, b, out, w: float, stream=0):
        """image_weighted_merge (imfilters.py:113-124): Image.blend with the 0 / 1 shortcuts."""
        if w == 0.0 or w == 1.0:
            src = a if w == 0.0 else b
            if out.data_ptr() != src.data_ptr():
                _lib.check(self.lib.havc_blend_u8(src.data_ptr(), src.data_ptr(), out.data_ptr(), out.numel(), 0.0, stream), "copy")
            return
        _lib.check(self.lib.havc_blend_u8(a.data_ptr(), b.data_ptr(), out.data_ptr(), out.numel(), w, stream), "blend")

    # ---- merges (mcomb.py) -------------------------------------------------------------------------------
    def combine(self, a, b, out, method: int, weight: float, cmc_p: Sequence = DEF_CMC_p, lmm_p: Sequence = DEF_LMM_p,
                alm_p: Sequence = DEF_ALM_p, crt_p: Sequence = DEF_CRT_p, stream: int = 0):
        """vs_sc_combine_models (mcomb.py:125-192) for two S x S batches (sat / hue tweaks are the caller's)."""
        for t in (a, b, out):
            self._chk(t)
        lib, (B, H, W) = self.lib, self._dims()
        chroma_threshold = cmc_p[0]
        red_fix, base_tol, max_extra = (cmc_p[1], cmc_p[2], cmc_p[3]) if len(cmc_p) > 1 else (True, 20, 24)
        t0, t1, t2 = self.tmp
        if method == 2:                                            # SimpleMerge
            return self.blend(a, b, out, weight, stream)
        if method in (3, 7):                                       # ConstrainedChromaMerge / ChromaBoundAdaptiveMerge
            stab = t0 if (red_fix or method == 3) else out
            _lib.check(lib.havc_chroma_stabilizer(a.data_ptr(), b.data_ptr(), stab.data_ptr(), B, H, W, int(method == 7),
                                                  float(chroma_threshold), int(base_tol), int(max_extra), float(weight),
                                                  self.stats[0].data_ptr() if red_fix else None, stream), "chroma_stabilizer")
            res = stab
            if red_fix:
                res = t1 if method == 3 else out
                _lib.check(lib.havc_red_fix(stab.data_ptr(), res.data_ptr(), B, H, W, self.stats[0].data_ptr(), stream), "red_fix")
            if method == 3:                                        # mcomb.py:173-177
                self.blend(a, b, t2, min(weight, 0.6), stream)
                self.blend(res, t2, out, 0.3, stream)
            return
        if method == 4:                                            # LumaMaskedMerge
            c = a
            if lmm_p[2] < 1:                                       # mcomb.py:242-245: clipc = vs_tweak(clipa, sat=luma_mask_sat)
                self.vs_tweak(a, t0, sat=float(lmm_p[2]), stream=stream)
                c = t0
            _lib.check(lib.havc_luma_masked_merge(a.data_ptr(), b.data_ptr(), c.data_ptr(), out.data_ptr(), B, H, W, float(lmm_p[0]),
                                                  float(lmm_p[1]), float(weight), stream), "luma_masked_merge")
            return
        if method == 5:                                            # AdaptiveLumaMerge
            self.frame_stats(b, self.stats[0], stream)
            _lib.check(lib.havc_adaptive_luma_merge(a.data_ptr(), b.data_ptr(), out.data_ptr(), B, H, W, self.stats[0].data_ptr(),
                                                    float(alm_p[0]), float(alm_p[1]), float(weight), float(alm_p[2]), stream),
                       "adaptive_luma_merge")
            return
        if method == 6:                                            # ChromaRetentionMerge
            sat, tht, alpha, resize, mask_weight, algo = crt_p[0], crt_p[1], crt_p[2], crt_p[3], crt_p[4], crt_p[5]
            alpha = max(min(alpha, DEF_MAX_COLOR_ALPHA), DEF_MIN_COLOR_ALPHA)
            if resize:                                             # mcomb.py:481-493: work on a Spline64-squeezed copy
                import math
                fs = min(min(max(math.trunc(0.4 * W / 16), 16), 48) * 16, W)
                if fs < W:                                         # "sanity check, avoid upscale"
                    return self._crt_resized(a, b, out, fs, sat, tht, alpha, mask_weight, algo, weight, stream)
            self.frame_stats(a, self.stats[0], stream)
            lut, lut_g = self._lut(tht, alpha, algo), self._lut(tht, max(alpha, 4.0), algo)
            if weight == 0:                                        # vs_simple_merge (vsfilters.py:730-739) returns clipa
                return self.blend(a, a, out, 0.0, stream)
            mw = -1.0 if weight == 1 else float(weight)            # weight 1: the restored clip itself, no std.Merge
            _lib.check(lib.havc_restore_color_gradient(b.data_ptr(), a.data_ptr(), out.data_ptr(), B, H, W, float(sat), lut.data_ptr(),
                                                       lut_g.data_ptr(), float(mask_weight), float(min(mask_weight, -0.5)),
                                                       self.stats[0].data_ptr(), float(mw), self.simd, stream),
                       "restore_color_gradient")
            return
        raise FilterError("HAVC: only dd_method in (0,6) is supported")          # mcomb.py:192

    def _crt_resized(self, a, b, out, fs, sat, tht, alpha, mask_weight, algo, weight, stream):
        """ChromaRetentionMerge(chroma_resize=True) (mcomb.py:481-512): both clips squeezed to fs x fs with Spline64, gradient
        colour restore there (frame-luma gate included, on the squeezed frame), Spline64 back, luma of clip_a recovered
        (vs_sc_recover_clip_luma), then std.Merge(clip_a, restored, weight)."""
        if weight == 0:
            return self.blend(a, a, out, 0.0, stream)
        key = ("crt", fs)
        if key not in self._luts:
            sq = SquareSqueeze(self.B, self.H, self.W, fs, self.dev)
            u8 = dict(dtype=torch.uint8, device=self.dev)
            small = [torch.empty(self.B, 3, fs, fs, **u8) for _ in range(3)]
            self._luts[key] = (sq, small, torch.zeros(self.B, 2, dtype=torch.int64, device=self.dev))
        sq, (sa, sb, sr), stats = self._luts[key]
        lib, B = self.lib, self.B
        sq.down(a, sa, stream)
        sq.down(b, sb, stream)
        _lib.check(lib.havc_frame_stats(sa.data_ptr(), B, fs, fs, 1.0, stats.data_ptr(), stream), "frame_stats")
        lut, lut_g = self._lut(tht, alpha, algo), self._lut(tht, max(alpha, 4.0), algo)
        _lib.check(lib.havc_restore_color_gradient(sb.data_ptr(), sa.data_ptr(), sr.data_ptr(), B, fs, fs, float(sat), lut.data_ptr(),
                                                   lut_g.data_ptr(), float(mask_weight), float(min(mask_weight, -0.5)),
                                                   stats.data_ptr(), -1.0, self.simd, stream), "restore_color_gradient")
        restored = out if weight == 1 else self.tmp[0]
        sq.up(sr, a, restored, stream, transplant=True)
        if weight != 1:
            _lib.check(lib.havc_vs_merge_u8(a.data_ptr(), restored.data_ptr(), out.data_ptr(), out.numel(), float(weight), stream),
                       "vs_merge")

    # ---- vs_tweak: the zimg YUV420P8 round trip (vsfilters.py:753-850) ---------------------------------------------
    def _zimg_state(self):
        if "zimg" not in self._luts:
            import math
            from . import resample
            B, H, W, dev = self.B, self.H, self.W, self.dev
            if H % 2 or W % 2:
                raise FilterError("vs_tweak: the YUV420P8 round trip needs an even frame size")
            up = lambda t: (torch.from_numpy(t[0]).to(dev), torch.from_numpy(np.ascontiguousarray(t[1])).to(dev), int(t[1].shape[1]))
            dh, dv, uh, uv = (up(t) for t in resample.chroma420_tables(W, H))
            f32 = dict(dtype=torch.float32, device=dev)
            u8 = dict(dtype=torch.uint8, device=dev)
            self._luts["zimg"] = dict(dh=dh, dv=dv, uh=uh, uv=uv, y=torch.empty(B, H, W, **u8), c=torch.empty(B, 2, H // 2, W // 2, **u8),
                                      s444=torch.empty(B, 2, H, W, **f32), s_half=torch.empty(B, 2, H // 2, W, **f32),
                                      s_rgb=torch.empty(B, 3, H, W, **f32), luts={})
        return self._luts["zimg"]

    def vs_tweak(self, img, out, hue: float = 0.0, sat: float = 1.0, bright: float = 0.0, cont: float = 1.0, gamma: float = 1.0,
                 stream: int = 0) -> bool:
        """vs_tweak (vsfilters.py:753-850) without gamma (std.Levels) / coring: RGB24 -> YUV420P8 (zimg Bicubic, BT.709 full
        range) -> hue / saturation rotation of (U, V) -> brightness / contrast table on Y -> RGB24 with error-diffusion dither.
        Returns False when it is the identity (`out` untouched).  The zimg steps are restated (library absent: parity unpinned)."""
        import math
        if hue == 0 and sat == 1 and bright == 0 and cont == 1 and gamma == 1:
            return False
        if gamma != 1:
            raise FilterError("vs_tweak: gamma != 1 (std.Levels) is not built")
        self._chk(img), self._chk(out)
        z, lib, (B, H, W) = self._zimg_state(), self.lib, self._dims()
        dh, dv, uh, uv = z["dh"], z["dv"], z["uh"], z["uv"]
        _lib.check(lib.havc_zimg_rgb_to_yuv420p8(img.data_ptr(), z["y"].data_ptr(), z["c"].data_ptr(), z["s444"].data_ptr(),
                                                 z["s_half"].data_ptr(), None, B, H, W, dv[0].data_ptr(), dv[1].data_ptr(), dv[2],
                                                 dh[0].data_ptr(), dh[1].data_ptr(), dh[2], 0, 0, 0, stream), "zimg.rgb_to_yuv420p8")
        if -1.0 < bright < 1.0:
            bright = bright * 255.0                                # vsfilters.py:792-793
        lut = None
        if bright != 0 or cont != 1:                               # vsfilters.py:828-841 (integer formats, no coring)
            key = (float(bright), float(cont))
            if key not in z["luts"]:
                tab = np.array([min(max(int(i * cont + bright + 0.5), 0), 255) for i in range(256)], np.uint8)
                z["luts"][key] = torch.from_numpy(tab).to(self.dev)
            lut = z["luts"][key]
        do_uv = hue != 0 or sat != 1
        h = hue * math.pi / 180.0
        c1, c2 = float(np.float32(math.cos(h) * sat)), float(np.float32(math.sin(h) * sat))
        _lib.check(lib.havc_zimg_tweak_yuv(z["y"].data_ptr(), z["c"].data_ptr(), B, H, W, c1, c2, int(do_uv),
                                           lut.data_ptr() if lut is not None else None, stream), "zimg.tweak_yuv")
        _lib.check(lib.havc_zimg_yuv420p8_to_rgb(z["y"].data_ptr(), z["c"].data_ptr(), out.data_ptr(), z["s_half"].data_ptr(),
                                                 z["s_rgb"].data_ptr(), B, H, W, uh[0].data_ptr(), uh[1].data_ptr(), uh[2],
                                                 uv[0].data_ptr(), uv[1].data_ptr(), uv[2], 0, 0, 1, stream), "zimg.yuv420p8_to_rgb")
        return True

    # ---- chroma-adjust filters ------------------------------------------------------------------------------
    def adjust_hue_range(self, img, out, hue_adjust: str, stream: int = 0) -> bool:
        """adjust_hue_range (restcolor.py:221-237).  Returns False when the filter is the identity."""
        if hue_adjust in ("none", ""):
            return False
        p = parse_hue_adjust(hue_adjust)
        if p is None:
            return False
        rng = hue_ranges_struct(p[0])
        if rng is None:
            return False
        _lib.check(self.lib.havc_adjust_chroma(img.data_ptr(), out.data_ptr(), *self._dims(), C.byref(rng), float(p[1]), int(p[2]),
                                               float(p[3]), self.simd, stream), "adjust_chroma")
        return True

    def image_tweak(self, img, out, sat=1.0, cont=1.0, bright=0.0, hue=0.0, gamma=1.0, hue_range="none", stream: int = 0) -> bool:
        """image_tweak (imfilters.py:463-504).  gamma != 1 raises like the reference does (its LUT has the wrong size);
        hue != 0 (Pillow HSV round trip) is not built."""
        if gamma != 1.0:
            raise FilterError("wrong number of lut entries")
        if hue != 0.0:
            raise FilterError("image_tweak: hue shift (Pillow HSV) is not built")
        rng = hue_ranges_struct(hue_range)
        if sat == 1.0 and cont == 1.0 and bright == 0.0:
            return False
        _lib.check(self.lib.havc_image_tweak(img.data_ptr(), out.data_ptr(), *self._dims(), float(sat), float(cont), float(bright),
                                             C.byref(rng) if rng is not None else None, self.stats[1].data_ptr(), stream), "image_tweak")
        return True

    def luma_adjusted_levels(self, img, out, luma_min=0.0, gamma=1.0, gamma_luma_min=0.0, gamma_alpha=0.0, gamma_min=0.2,
                             stream: int = 0):
        """luma_adjusted_levels (imfilters.py:335-372)."""
        self.frame_stats(img, self.stats[1], stream)
        _lib.check(self.lib.havc_luma_adjusted_levels(img.data_ptr(), out.data_ptr(), *self._dims(), self.stats[1].data_ptr(),
                                                      float(luma_min), float(gamma), float(gamma_luma_min), float(gamma_alpha),
                                                      float(gamma_min), stream), "luma_adjusted_levels")


    # ---- HAVC_stabilizer per-frame stages (vsdeoldify/__init__.py:2823-2861; vsfilters.py:525-641) ----------------
    def chroma_tweak(self, img, out, sat=1.0, bright=0.0, hue=0, hue_adjust="none", luma_merge: Optional[Tuple[float, float]] = None,
                     stream: int = 0) -> bool:
        """image_chroma_tweak (imfilters.py:540-550 -> restcolor.np_image_chroma_tweak :288-342), optionally fused with the luma
        merge of vs_sc_chroma_bright_tweak (img_dark = tweaked, img_white = img).  Returns False when it is the identity."""
        if sat == 1 and bright == 0 and hue == 0 and hue_adjust == "none":
            if luma_merge is None:
                return False
            # the tweak returns img itself, but the float luma merge of img with img still truncates (a*(1-m) + a*m can land
            # just below a), so it runs
            _lib.check(self.lib.havc_luma_masked_merge(img.data_ptr(), img.data_ptr(), img.data_ptr(), out.data_ptr(), *self._dims(),
                                                       float(luma_merge[0]), float(luma_merge[1]), 1.0, stream), "luma_merge")
            return True
        rng, sat2, hue2, weight = None, 1.0, 0, 0.0
        if hue_adjust not in ("none", ""):
            p = parse_hue_adjust(hue_adjust)
            if p is not None:
                rng = hue_ranges_struct(p[0])
                if rng is None:
                    raise FilterError("HybridAVC: unknown hue name: " + str(p[0]))
                sat2, hue2, weight = p[1], p[2], p[3]
        lm = luma_merge is not None
        _lib.check(self.lib.havc_chroma_tweak(img.data_ptr(), out.data_ptr(), *self._dims(), float(sat), float(bright), int(hue),
                                              C.byref(rng) if rng is not None else None, float(sat2), int(hue2), float(weight),
                                              int(lm), float(luma_merge[0]) if lm else 0.0, float(luma_merge[1]) if lm else 0.0,
                                              self.simd, stream), "chroma_tweak")
        return True

    def dark_tweak(self, img, out, dark_threshold=0.3, dark_amount=0.8, dark_hue_adjust="none", stream: int = 0):
        """vs_sc_dark_tweak.merge_frame (vsfilters.py:604-636): image_tweak(bright, sat, hue_range) merged back by luma."""
        d_threshold = 0.1
        d_white = min(max(dark_threshold, d_threshold), 0.50)
        d_sat = min(max(1.1 - dark_amount, 0.10), 0.80)
        d_bright = -min(max(dark_amount, 0.20), 0.90)
        t = self.tmp[2]
        self.image_tweak(img, t, sat=d_sat, bright=d_bright, hue_range=dark_hue_adjust, stream=stream)   # never the identity
        _lib.check(self.lib.havc_luma_masked_merge(img.data_ptr(), img.data_ptr(), t.data_ptr(), out.data_ptr(), *self._dims(),
                                                   float(d_threshold), float(d_white), 1.0, stream), "dark_tweak.luma_merge")

    def stabilizer_stages(self, img, out, dark=False, dark_p=(0.2, 0.8), smooth=False, smooth_p=(0.3, 0.7, 0.9, 0.0, "none"),
                          colormap_adjust: str = "none", stream: int = 0) -> bool:
        """The per-frame stages of HAVC_stabilizer on S x S batches, in the reference's order: vs_dark_tweak,
        vs_chroma_bright_tweak, vs_colormap.  The result is in `out`; returns False if no stage ran (out untouched)."""
        self._chk(img), self._chk(out)
        cur, ping = img, [out, self.tmp[0]]

        def nxt():
            return ping[0] if cur is not ping[0] else ping[1]
        if dark:
            dst = nxt()
            self.dark_tweak(cur, dst, dark_p[0], dark_p[1], (dark_p[2] if len(dark_p) > 2 else "none").lower(), stream)
            cur = dst
        if smooth:
            dst = nxt()
            adj = (smooth_p[4] if len(smooth_p) > 4 else "none").lower()
            if self.chroma_tweak(cur, dst, sat=smooth_p[2], bright=-smooth_p[3], hue_adjust=adj,
                                 luma_merge=(smooth_p[0], smooth_p[1]), stream=stream):
                cur = dst
        if colormap_adjust not in ("none", ""):
            dst = nxt()
            if self.chroma_tweak(cur, dst, hue_adjust=colormap_adjust, stream=stream):
                cur = dst
        if cur is img:
            return False
        if cur is not out:
            self.blend(cur, cur, out, 0.0, stream)
        return True


def stab_weight_list(nframes: int, mode: str) -> list:
    """The temporal weights of vs_chroma_stabilizer_ex / vs_clip_color_stabilizer (vsfilters.py:40-51, 118-160), scale 100."""
    if nframes % 2 == 0:
        nframes += 1
    n = max(3, min(nframes, 15))
    nh = round((n - 1) / 2)
    if mode in ("A", "arithmetic", "center"):
        wi = math.trunc(100.0 / n)
        return [wi] * nh + [100 - (n - 1) * wi] + [wi] * nh
    if mode in ("W", "weighted", "left", "right"):
        base = n * (n + 1) * 0.5
        side = [math.trunc(100 * (i + 1) / base) for i in range(nh)]
        return side + [100 - 2 * sum(side)] + side
    raise FilterError("HybridAVC: unknown average method: " + str(mode))


def scene_folded_weights(weights, prev_flags, next_flags) -> list:
    """std.AverageFrames(scenechange=True) (vsfilters.py:58; VapourSynth's filter restated, unpinned): the weights of the frames
    beyond a _SceneChangePrev (left) / _SceneChangeNext (right) boundary are added to the boundary frame's."""
    w = list(weights)
    n, c = len(w), len(w) // 2
    lo, hi = 0, n - 1
    for i in range(c, 0, -1):
        if prev_flags[i]:
            lo = i
            break
    for i in range(c, n - 1):
        if next_flags[i]:
            hi = i
            break
    for i in range(lo):
        w[lo] += w[i]
        w[i] = 0
    for i in range(n - 1, hi, -1):
        w[hi] += w[i]
        w[i] = 0
    return w


class TemporalStabilizer:
    """The temporal chroma stabiliser, scope row N3: vs_chroma_stabilizer_ex (vsslib/vsfilters.py:84-287; algo = 0, the value
    HAVC_stabilizer passes) on a batch of B frames that carries its `nh` halo frames on either side: `seq` u8
    [B + 2 nh, 3, H, W] (frame b of the batch is seq[nh + b]; the caller clamps the halo to the clip's ends like
    std.AverageFrames does), result u8 [B, 3, H, W].

      tht > 0 (_average_clips_ex, :214-249): for every offset d != 0 the frame with the chroma of frame n + d
        (vs_get_clip_frame, :259-287: YUV420P8 round trip, error diffusion on the way back) is repaired against frame n
        (vs_recover_clip_color -> restore_color: gray pixels take frame n's colours; frames n < 15 pass), converted to YUV420P8
        with error diffusion, and the chroma planes of the 2 nh + 1 clips are averaged with the weight table; luma = first clip's.
      tht = 0 (vs_clip_color_stabilizer, :38-63): std.AverageFrames of the neighbours' chroma, scene-change aware
        (per-frame weights, folded on the host from the frame props).
    Every zimg / std.AverageFrames step is a restatement (csrc/zimg.cu, oracle/zimg_oracle.py: parity unpinned).  The serial part
    (Floyd-Steinberg) runs as one wavefront warp per plane and frame: B * 3 planes in flight per launch."""

    def __init__(self, B: int, H: int, W: int, device, nframes: int = 5, mode: str = "A", sat: float = 1.0, tht: int = 0,
                 weight: float = 0.5, tht_scen: float = 0.8, hue_adjust: str = "none", simd_width: int = CV_SIMD_WIDTH):
        from . import resample
        if H % 2 or W % 2:
            raise FilterError("HAVC_stabilizer: the temporal stabiliser's YUV420P8 round trips need an even frame size")
        self.lib, self.dev = _lib.lib(), torch.device(device)
        self.B, self.H, self.W = B, H, W
        self.wl = stab_weight_list(nframes, mode)
        self.nh = len(self.wl) // 2
        self.sat, self.tht, self.weight, self.tht_scen = float(sat), int(tht), float(weight), float(tht_scen)
        self.hue_adjust = (hue_adjust or "none").lower()
        self.simd = simd_width
        dev, n, nc, T, K = self.dev, H * W, (H // 2) * (W // 2), B + 2 * self.nh, 2 * self.nh + 1
        up = lambda t: (torch.from_numpy(t[0]).to(dev), torch.from_numpy(np.ascontiguousarray(t[1])).to(dev), int(t[1].shape[1]))
        self.dh, self.dv, self.uh, self.uv = (up(t) for t in resample.chroma420_tables(W, H))
        u8, f32 = dict(dtype=torch.uint8, device=dev), dict(dtype=torch.float32, device=dev)
        self.y_nd, self.c_nd = torch.empty(T, H, W, **u8), torch.empty(T, 2, H // 2, W // 2, **u8)      # no-dither YUV of the sequence
        self.ys, self.cs = torch.empty(K, B, H, W, **u8), torch.empty(K, B, 2, H // 2, W // 2, **u8)    # the clips std.AverageFrames sees
        self.c_avg = torch.empty(B, 2, H // 2, W // 2, **u8)
        self.rgb_a, self.rgb_b = torch.empty(B, 3, H, W, **u8), torch.empty(B, 3, H, W, **u8)
        self.s444, self.s_half = torch.empty(T, 2, H, W, **f32), torch.empty(T, 2, H // 2, W, **f32)
        self.s_rgb, self.s_q = torch.empty(B, 3, H, W, **f32), torch.empty(B * n * 3 // 2, **f32)
        self.stats = torch.zeros(B, 2, dtype=torch.int64, device=dev)
        self.lut = torch.from_numpy(np.where(np.arange(256) < self.tht, 255, 0).astype(np.uint8)).to(dev)
        self.w_shared = torch.tensor(self.wl, dtype=torch.int32, device=dev)
        self.bank = FilterBank(B, H, W, dev, simd_width) if self.hue_adjust not in ("none", "") else None

    def _to_yuv(self, rgb, y, c, count, dither, st):
        dh, dv = self.dh, self.dv
        _lib.check(self.lib.havc_zimg_rgb_to_yuv420p8(rgb.data_ptr(), y.data_ptr(), c.data_ptr(), self.s444.data_ptr(), self.s_half.data_ptr(),
                                                      self.s_q.data_ptr() if dither else None, count, self.H, self.W, dv[0].data_ptr(),
                                                      dv[1].data_ptr(), dv[2], dh[0].data_ptr(), dh[1].data_ptr(), dh[2], 0, 0, int(dither), st),
                   "temporal.rgb_to_yuv420p8")

    def _to_rgb(self, y, c, rgb, st):
        uh, uv = self.uh, self.uv
        _lib.check(self.lib.havc_zimg_yuv420p8_to_rgb(y.data_ptr(), c.data_ptr(), rgb.data_ptr(), self.s_half.data_ptr(), self.s_rgb.data_ptr(),
                                                      self.B, self.H, self.W, uh[0].data_ptr(), uh[1].data_ptr(), uh[2], uv[0].data_ptr(),
                                                      uv[1].data_ptr(), uv[2], 0, 0, 1, st), "temporal.yuv420p8_to_rgb")

    def run(self, seq, out, active=None, weights=None, stream: int = 0):
        """seq [B + 2 nh, 3, H, W] -> out [B, 3, H, W].  active: device u8 [B], 0 where the frame number is < 15 (tht > 0 only);
        weights: device int32 [B, 2 nh + 1] (tht == 0 with scene changes; None = the plain table for every frame)."""
        lib, B, H, W, nh, st = self.lib, self.B, self.H, self.W, self.nh, stream
        T, K, n, nc = B + 2 * nh, 2 * nh + 1, H * W, 2 * (H // 2) * (W // 2)
        assert tuple(seq.shape) == (T, 3, H, W) and tuple(out.shape) == (B, 3, H, W) and seq.is_contiguous() and out.is_contiguous()
        self._to_yuv(seq, self.y_nd, self.c_nd, T, False, st)                              # vsfilters.py:56 / :275
        cur = seq[nh:nh + B]
        if self.tht == 0:
            w = self.w_shared if weights is None else weights
            _lib.check(lib.havc_average_frames_u8(self.c_nd.data_ptr(), nc, K, w.data_ptr(), int(weights is not None), 100,
                                                  self.c_avg.data_ptr(), B, nc, st), "temporal.average_frames")
            self._to_rgb(self.y_nd[nh:nh + B], self.c_avg, out, st)                        # vsfilters.py:60
            return
        assert active is not None
        for k, d in enumerate(range(-nh, nh + 1)):
            if d == 0:
                self._to_yuv(cur, self.ys[k], self.cs[k], B, True, st)                     # vsfilters.py:222
                continue
            self._to_rgb(self.y_nd[nh:nh + B], self.c_nd[nh + d:nh + d + B], self.rgb_a, st)          # vs_get_clip_frame
            _lib.check(lib.havc_gray_mask_stats(self.rgb_a.data_ptr(), B, H, W, self.tht, self.stats.data_ptr(), st), "temporal.stats")
            _lib.check(lib.havc_restore_color(cur.data_ptr(), self.rgb_a.data_ptr(), self.rgb_b.data_ptr(), B, H, W, self.sat, self.tht,
                                              self.weight, self.tht_scen, self.stats.data_ptr(), active.data_ptr(), self.lut.data_ptr(),
                                              self.simd, st), "temporal.restore_color")
            self._to_yuv(self.rgb_b, self.ys[k], self.cs[k], B, True, st)                  # vsfilters.py:230,239
        _lib.check(lib.havc_average_frames_u8(self.cs.data_ptr(), B * nc, K, self.w_shared.data_ptr(), 0, 100, self.c_avg.data_ptr(), B, nc,
                                              st), "temporal.average_frames")
        dst = out if self.bank is None else self.rgb_a
        self._to_rgb(self.ys[0], self.c_avg, dst, st)                                      # vsfilters.py:242-247: luma of the first clip
        if self.bank is not None and not self.bank.adjust_hue_range(dst, out, self.hue_adjust, st):       # vsfilters.py:113
            self.bank.blend(dst, dst, out, 0.0, st)


class TemporalEngine:
    """vs_chroma_stabilizer_ex on batches of planar RGB24 host frames at the clip's own size (the function is also used on its own,
    outside HAVC_stabilizer): H2D of the batch with its halo -> TemporalStabilizer -> D2H, one CUDA graph."""

    def __init__(self, width: int, height: int, batch: int = 8, device: str = "cuda:0", **stab):
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        self.temporal = TemporalStabilizer(batch, height, width, self.dev, **stab)
        self.nh, self.out_B, self.B, self.H, self.W = self.temporal.nh, batch, batch + 2 * self.temporal.nh, height, width
        u8 = dict(dtype=torch.uint8, device=self.dev)
        self.d_in, self.d_out = torch.empty(self.B, 3, height, width, **u8), torch.empty(batch, 3, height, width, **u8)
        self.active = torch.ones(batch, **u8)
        self.tw = torch.zeros(batch, 2 * self.nh + 1, dtype=torch.int32, device=self.dev)
        self.h_in = torch.empty(self.B, 3, height, width, dtype=torch.uint8).pin_memory()
        self.h_out = torch.empty(batch, 3, height, width, dtype=torch.uint8).pin_memory()
        self.h_active = torch.ones(batch, dtype=torch.uint8).pin_memory()
        self.h_tw = torch.zeros(batch, 2 * self.nh + 1, dtype=torch.int32).pin_memory()
        self.stream = torch.cuda.Stream(device=self.dev)
        with torch.cuda.stream(self.stream):
            self.d_in.zero_()
            self._launch(self.stream.cuda_stream)
        self.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self._launch(torch.cuda.current_stream().cuda_stream)
        self.stream.synchronize()

    def _launch(self, st: int):
        self.temporal.run(self.d_in, self.d_out, active=self.active, weights=self.tw if self.temporal.tht == 0 else None, stream=st)

    def process_sequence(self, frames, n0, weights=None):
        return StabilizerEngine.process_sequence(self, frames, n0, weights)


class ResizeEngine:
    """resize_min_HW and resize_to_chroma (vsslib/vsresize.py:30-127) on batches of planar RGB24 host frames - the
    `chroma_resize=True` detour HAVC_main takes for every preset but the two slowest (vsdeoldify/__init__.py:492-494, restore_format
    havc_utils.py:183-184):
      down(high) : Spline36 to (low_w, low_h)
      chroma(high, low): Spline36 of `low` back to the size of `high`, both to YUV420P8 (BT.709, full range, no dither), the Y plane
                  of `high` with the chroma of `low`, RGB24 with error-diffusion dither.
    zimg is restated (parity unpinned)."""

    def __init__(self, width: int, height: int, low_w: int, low_h: int, batch: int = 8, device: str = "cuda:0"):
        from . import resample
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        if height % 2 or width % 2:
            raise FilterError("resize_to_chroma: the YUV420P8 round trip needs an even frame size")
        self.B, self.H, self.W, self.h, self.w = batch, height, width, low_h, low_w
        B, H, W, dev = batch, height, width, self.dev
        self.same = (low_w, low_h) == (width, height)
        self.r_down = None if self.same else RectResize(B, H, W, low_h, low_w, dev, "spline36")
        self.r_up = None if self.same else RectResize(B, low_h, low_w, H, W, dev, "spline36")
        u8, f32 = dict(dtype=torch.uint8, device=dev), dict(dtype=torch.float32, device=dev)
        self.d_high, self.d_up, self.d_out = (torch.empty(B, 3, H, W, **u8) for _ in range(3))
        self.d_low = torch.empty(B, 3, low_h, low_w, **u8)
        self.h_high, self.h_out = (torch.empty(B, 3, H, W, dtype=torch.uint8).pin_memory() for _ in range(2))
        self.h_low = torch.empty(B, 3, low_h, low_w, dtype=torch.uint8).pin_memory()
        up = lambda t: (torch.from_numpy(t[0]).to(dev), torch.from_numpy(np.ascontiguousarray(t[1])).to(dev), int(t[1].shape[1]))
        self.dh, self.dv, self.uh, self.uv = (up(t) for t in resample.chroma420_tables(W, H))
        self.y_hi, self.c_hi = torch.empty(B, H, W, **u8), torch.empty(B, 2, H // 2, W // 2, **u8)
        self.y_lo, self.c_lo = torch.empty(B, H, W, **u8), torch.empty(B, 2, H // 2, W // 2, **u8)
        self.s444, self.s_half, self.s_rgb = torch.empty(B, 2, H, W, **f32), torch.empty(B, 2, H // 2, W, **f32), torch.empty(B, 3, H, W, **f32)
        self.stream = torch.cuda.Stream(device=dev)

    def down(self, frames: np.ndarray) -> np.ndarray:
        n = frames.shape[0]
        if self.same:
            return frames
        self.h_high[:n].copy_(torch.from_numpy(np.ascontiguousarray(frames)))
        with torch.cuda.stream(self.stream):
            self.d_high.copy_(self.h_high, non_blocking=True)
            self.r_down.run(self.d_high, self.d_low, self.stream.cuda_stream)
            self.h_low.copy_(self.d_low, non_blocking=True)
        self.stream.synchronize()
        return self.h_low[:n].numpy().copy()

    def chroma(self, high: np.ndarray, low: np.ndarray) -> np.ndarray:
        n = high.shape[0]
        self.h_high[:n].copy_(torch.from_numpy(np.ascontiguousarray(high)))
        self.h_low[:n].copy_(torch.from_numpy(np.ascontiguousarray(low)))
        lib, B, H, W, dh, dv, uh, uv, st = _lib.lib(), self.B, self.H, self.W, self.dh, self.dv, self.uh, self.uv, self.stream.cuda_stream
        with torch.cuda.stream(self.stream):
            self.d_high.copy_(self.h_high, non_blocking=True)
            self.d_low.copy_(self.h_low, non_blocking=True)
            if self.same:
                colour = self.d_low
            else:
                self.r_up.run(self.d_low, self.d_up, st)
                colour = self.d_up
            for rgb, y, c in ((self.d_high, self.y_hi, self.c_hi), (colour, self.y_lo, self.c_lo)):
                _lib.check(lib.havc_zimg_rgb_to_yuv420p8(rgb.data_ptr(), y.data_ptr(), c.data_ptr(), self.s444.data_ptr(), self.s_half.data_ptr(),
                                                         None, B, H, W, dv[0].data_ptr(), dv[1].data_ptr(), dv[2], dh[0].data_ptr(),
                                                         dh[1].data_ptr(), dh[2], 0, 0, 0, st), "resize_to_chroma.to_yuv")
            _lib.check(lib.havc_zimg_yuv420p8_to_rgb(self.y_hi.data_ptr(), self.c_lo.data_ptr(), self.d_out.data_ptr(), self.s_half.data_ptr(),
                                                     self.s_rgb.data_ptr(), B, H, W, uh[0].data_ptr(), uh[1].data_ptr(), uh[2], uv[0].data_ptr(),
                                                     uv[1].data_ptr(), uv[2], 0, 0, 1, st), "resize_to_chroma.to_rgb")
            self.h_out.copy_(self.d_out, non_blocking=True)
        self.stream.synchronize()
        return self.h_out[:n].numpy().copy()


class MergeEngine:
    """HAVC_merge on batches of planar RGB24 host frames: H2D -> FilterBank.combine / std.Merge -> D2H."""

    def __init__(self, width: int, height: int, batch: int = 8, device: str = "cuda:0"):
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        self.B, self.H, self.W = batch, height, width
        self.bank = FilterBank(batch, height, width, self.dev)
        u8 = dict(dtype=torch.uint8, device=self.dev)
        self.d_a, self.d_b, self.d_out = (torch.empty(batch, 3, height, width, **u8) for _ in range(3))
        self.h_a, self.h_b, self.h_out = (torch.empty(batch, 3, height, width, dtype=torch.uint8).pin_memory() for _ in range(3))
        self.stream = torch.cuda.Stream(device=self.dev)

    def merge_batch(self, a: np.ndarray, b: np.ndarray, method: int, weight: float, cmc_p=DEF_CMC_p, lmm_p=DEF_LMM_p,
                    alm_p=DEF_ALM_p, crt_p=DEF_CRT_p) -> np.ndarray:
        """a, b: uint8 [n<=B, 3, H, W].  method 2 = std.Merge (HAVC_merge, vsdeoldify/__init__.py:2648), 3..7 =
        vs_combine_models (mcomb.py:125-192)."""
        n = a.shape[0]
        assert a.shape == b.shape and n <= self.B and a.shape[1:] == (3, self.H, self.W)
        self.h_a[:n].copy_(torch.from_numpy(np.ascontiguousarray(a)))
        self.h_b[:n].copy_(torch.from_numpy(np.ascontiguousarray(b)))
        with torch.cuda.stream(self.stream):
            st = self.stream.cuda_stream
            self.d_a.copy_(self.h_a, non_blocking=True)
            self.d_b.copy_(self.h_b, non_blocking=True)
            if method == 2:
                _lib.check(self.bank.lib.havc_vs_merge_u8(self.d_a.data_ptr(), self.d_b.data_ptr(), self.d_out.data_ptr(),
                                                          self.d_out.numel(), float(weight), st), "vs_merge")
            else:
                self.bank.combine(self.d_a, self.d_b, self.d_out, method, weight, cmc_p, lmm_p, alm_p, crt_p, stream=st)
            self.h_out.copy_(self.d_out, non_blocking=True)
        self.stream.synchronize()
        return self.h_out[:n].numpy().copy()


class LumaMergeEngine:
    """HAVC_merge(clipa, clipb, clip_luma=...) (vsdeoldify/__init__.py:2633-2675) on host batches:
      methods 3..7: clipa / clipb are squeezed to frame_size x frame_size with Spline64 (frame_size from 0.4 * clip_luma.width,
                    :2661-2664), merged there by vs_combine_models, and the result goes through _clip_chroma_resize(clip_luma, .)
                    (:2670-2673): Spline64 to clip_luma's size + the luma of clip_luma;
      methods 0 / 1 (or weight 0 / 1): _clip_chroma_resize(clip_luma, clipa | clipb) alone."""

    def __init__(self, ab_size: Tuple[int, int], luma_size: Tuple[int, int], squeeze: bool, batch: int = 8, device: str = "cuda:0"):
        import math
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        (self.Wab, self.Hab), (self.W, self.H), self.B = ab_size, luma_size, batch
        self.lib = _lib.lib()
        u8 = dict(dtype=torch.uint8, device=self.dev)
        B = batch
        self.d_a, self.d_b = (torch.empty(B, 3, self.Hab, self.Wab, **u8) for _ in range(2))
        self.d_luma, self.d_out = (torch.empty(B, 3, self.H, self.W, **u8) for _ in range(2))
        self.squeeze = squeeze
        if squeeze:
            rf = min(max(math.trunc(0.4 * self.W / 16), 16), 32)
            self.S = min(rf * 16, self.W)
            self.sq = SquareSqueeze(B, self.Hab, self.Wab, self.S, self.dev, out_hw=(self.H, self.W))
            self.bank = FilterBank(B, self.S, self.S, self.dev)
            self.small = [torch.empty(B, 3, self.S, self.S, **u8) for _ in range(3)]
        else:       # plain Spline64 resize of one clip to clip_luma's size: vertical pass, then horizontal pass + transplant
            from .engine import _Tables
            self.t_v, self.t_h = _Tables(self.Hab, self.H, "spline64", self.dev), _Tables(self.Wab, self.W, "spline64", self.dev)
            self.tmp = torch.empty(B, 3, self.H, self.Wab, dtype=torch.float32, device=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)

    def merge_batch(self, a, b, luma, method: int, weight: float, cmc_p=DEF_CMC_p, lmm_p=DEF_LMM_p, alm_p=DEF_ALM_p, crt_p=DEF_CRT_p):
        """a / b: uint8 [n, 3, Hab, Wab] (one of them may be None when not squeezing), luma: uint8 [n, 3, H, W]."""
        n = luma.shape[0]
        lib, B, chk = self.lib, self.B, _lib.check
        up = lambda t, x: t[:n].copy_(torch.from_numpy(np.ascontiguousarray(x)), non_blocking=True)
        with torch.cuda.stream(self.stream):
            st = self.stream.cuda_stream
            up(self.d_luma, luma)
            if self.squeeze:
                up(self.d_a, a), up(self.d_b, b)
                self.sq.down(self.d_a, self.small[0], st)
                self.sq.down(self.d_b, self.small[1], st)
                self.bank.combine(self.small[0], self.small[1], self.small[2], method, weight, cmc_p, lmm_p, alm_p, crt_p, stream=st)
                self.sq.up(self.small[2], self.d_luma, self.d_out, st, transplant=True)
            else:
                src = self.d_a
                up(src, a if a is not None else b)
                chk(lib.havc_resample_v(src.data_ptr(), self.tmp.data_ptr(), B * 3, self.Hab, self.H, self.Wab, self.t_v.start.data_ptr(),
                                        self.t_v.w.data_ptr(), self.t_v.taps, st), "resize.v")
                chk(lib.havc_post_horizontal(self.tmp.data_ptr(), self.d_luma.data_ptr(), self.d_out.data_ptr(), B, self.Wab, self.H, self.W,
                                             self.t_h.start.data_ptr(), self.t_h.wt.data_ptr(), self.t_h.taps, 1, st), "resize.h")
            out = self.d_out[:n].cpu()
        self.stream.synchronize()
        return out.numpy()


class StabilizerEngine:
    """HAVC_stabilizer's per-frame path on batches of planar RGB24 host frames (vsdeoldify/__init__.py:2748-2873, stab=False):
    Spline64 squeeze to S x S (:2804) -> vs_dark_tweak / vs_chroma_bright_tweak / vs_colormap -> _clip_chroma_resize (:3545-3554:
    Spline64 back to W x H + full-resolution luma transplant), one CUDA graph per batch."""

    def __init__(self, width: int, height: int, frame_size: int, stages: dict, batch: int = 8, device: str = "cuda:0",
                 resize_kernel: str = "spline64", stab: Optional[dict] = None):
        from .engine import _Tables
        self.lib = _lib.lib()
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        # `stab` (the temporal stage, N3): the device batch carries nh halo frames on either side, every per-frame stage runs on
        # all of them (the temporal filter reads its neighbours AFTER dark / smooth / colormap, __init__.py:2843-2861), the way
        # back only on the `out_B` frames in the middle
        self.temporal = TemporalStabilizer(batch, frame_size, frame_size, self.dev, **stab) if stab else None
        self.nh = self.temporal.nh if self.temporal else 0
        self.out_B = batch
        batch = batch + 2 * self.nh
        self.B, self.H, self.W, self.S = batch, height, width, frame_size
        B, H, W, S = batch, height, width, frame_size
        self.stages = stages
        self.bank = FilterBank(B, S, S, self.dev)
        self.t_down_h, self.t_down_v = _Tables(W, S, resize_kernel, self.dev), _Tables(H, S, resize_kernel, self.dev)
        self.t_up_h, self.t_up_v = _Tables(S, W, resize_kernel, self.dev), _Tables(S, H, resize_kernel, self.dev)
        u8, f32 = dict(dtype=torch.uint8, device=self.dev), dict(dtype=torch.float32, device=self.dev)
        self.d_in, self.d_out = torch.empty(B, 3, H, W, **u8), torch.empty(B, 3, H, W, **u8)
        self.h_in = torch.empty(B, 3, H, W, dtype=torch.uint8).pin_memory()
        self.h_out = torch.empty(B, 3, H, W, dtype=torch.uint8).pin_memory()
        self.tmp_f = torch.empty(B, 3, H, S, **f32)
        self.small, self.small_out = torch.empty(B, 3, S, S, **u8), torch.empty(B, 3, S, S, **u8)
        self.small_t = torch.empty(self.out_B, 3, S, S, **u8) if self.temporal else None
        self.active = torch.ones(self.out_B, **u8)                                        # frame number >= 15 (vsfilters.py:337)
        self.tw = torch.zeros(self.out_B, 2 * self.nh + 1, dtype=torch.int32, device=self.dev)   # per-frame weights (tht == 0)
        self.h_active = torch.ones(self.out_B, dtype=torch.uint8).pin_memory()
        self.h_tw = torch.zeros(self.out_B, 2 * self.nh + 1, dtype=torch.int32).pin_memory()
        self.x_scratch = torch.empty(B, S, S, 8, dtype=torch.float16, device=self.dev)    # pre_vertical's network-input by-product
        self.stream = torch.cuda.Stream(device=self.dev)
        with torch.cuda.stream(self.stream):
            self.d_in.zero_()
            self._launch(self.stream.cuda_stream)          # warm-up also surfaces FilterError for unsupported parameters
        self.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self._launch(torch.cuda.current_stream().cuda_stream)
        self.stream.synchronize()

    def _launch(self, st: int):
        lib, B, S, W, H, chk = self.lib, self.B, self.S, self.W, self.H, _lib.check
        td, tv, uh, uv = self.t_down_h, self.t_down_v, self.t_up_h, self.t_up_v
        td.resample_h(lib, self.d_in.data_ptr(), self.tmp_f.data_ptr(), B * 3 * H, st, "squeeze.h")
        chk(lib.havc_pre_vertical(self.tmp_f.data_ptr(), self.small.data_ptr(), self.x_scratch.data_ptr(), B, H, S, tv.start.data_ptr(),
                                  tv.w.data_ptr(), tv.taps, 0, st), "squeeze.v")
        res = self.small_out if self.bank.stabilizer_stages(self.small, self.small_out, stream=st, **self.stages) else self.small
        src, nb = self.d_in, B
        if self.temporal is not None:
            self.temporal.run(res, self.small_t, active=self.active, weights=self.tw if self.temporal.tht == 0 else None, stream=st)
            res, nb, src = self.small_t, self.out_B, self.d_in[self.nh:self.nh + self.out_B]
        chk(lib.havc_resample_v(res.data_ptr(), self.tmp_f.data_ptr(), nb * 3, S, H, S, uv.start.data_ptr(), uv.w.data_ptr(), uv.taps, st),
            "unsqueeze.v")
        uh.post_horizontal(lib, self.tmp_f.data_ptr(), src.data_ptr(), self.d_out.data_ptr(), nb, H, 1, st, "unsqueeze.h")

    def process_batch(self, frames: np.ndarray, skip=None) -> np.ndarray:
        """frames: uint8 [n<=B, 3, H, W] -> uint8 [n, 3, H, W] (the stabilizer's selectors run with scenechange=False: no gate)."""
        n = frames.shape[0]
        assert n <= self.B and frames.shape[1:] == (3, self.H, self.W) and frames.dtype == np.uint8
        self.h_in[:n].copy_(torch.from_numpy(np.ascontiguousarray(frames)))
        with torch.cuda.stream(self.stream):
            self.d_in.copy_(self.h_in, non_blocking=True)
            self.graph.replay()
            self.h_out.copy_(self.d_out, non_blocking=True)
        self.stream.synchronize()
        return self.h_out[:n].numpy().copy()

    def process_sequence(self, frames: np.ndarray, n0: int, weights: Optional[np.ndarray] = None) -> np.ndarray:
        """The temporal form: frames uint8 [out_B + 2 nh, 3, H, W] = the batch that starts at clip frame n0 with its halo (clamped
        to the clip's ends by the caller) -> uint8 [out_B, 3, H, W].  weights: int32 [out_B, 2 nh + 1] (scene-change folded
        std.AverageFrames weights, tht == 0 only)."""
        assert self.temporal is not None and frames.shape == (self.B, 3, self.H, self.W) and frames.dtype == np.uint8
        self.h_in.copy_(torch.from_numpy(np.ascontiguousarray(frames)))
        self.h_active.copy_(torch.from_numpy((np.arange(n0, n0 + self.out_B) >= 15).astype(np.uint8)))
        self.h_tw.copy_(torch.from_numpy(np.ascontiguousarray(weights, dtype=np.int32)) if weights is not None
                        else torch.tensor(self.temporal.wl, dtype=torch.int32).expand(self.out_B, -1))
        with torch.cuda.stream(self.stream):
            self.d_in.copy_(self.h_in, non_blocking=True)
            self.active.copy_(self.h_active, non_blocking=True)
            self.tw.copy_(self.h_tw, non_blocking=True)
            self.graph.replay()
            self.h_out.copy_(self.d_out, non_blocking=True)
        self.stream.synchronize()
        return self.h_out[:self.out_B].numpy().copy()
