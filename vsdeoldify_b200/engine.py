"""Per-GPU frame engine: planar RGB24 frames in (host) -> colourised planar RGB24 frames out (host).

Implements the device side of HAVC_colorizer(method=0) (vsdeoldify/__init__.py:2290-2523):
  Spline64 squeeze to S x S -> DeOldify generator -> S x S luma transplant -> Spline64 back to W x H ->
  full-resolution luma transplant,
as one CUDA graph of libhavc_b200 launches per batch of B frames, fed through pinned double buffers on
separate copy streams.  torch supplies device/pinned memory, streams and graph capture only.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import os

import numpy as np
import torch

from . import _lib, ops, resample
from .unet import UnetProgram


_NO_PERIODIC = bool(os.environ.get("HAVC_B200_NO_PERIODIC"))      # A/B switch for profiling: table-driven horizontal passes


class _Tables:
    def __init__(self, src: int, dst: int, kernel: str, dev):
        start, w = resample.build_tables(src, dst, kernel)
        self.taps = int(w.shape[1])
        self.start = torch.from_numpy(start).to(dev)
        self.w = torch.from_numpy(w).contiguous().to(dev)                       # [out][taps]  (vertical passes)
        self.wt = torch.from_numpy(np.ascontiguousarray(w.T)).to(dev)           # [taps][out]  (horizontal passes)
        # integer ratios: the phase-periodic horizontal kernels (csrc/pixel.cu); None = the table kernels
        self.src, self.dst = src, dst
        plan = None if _NO_PERIODIC else (resample.periodic_plan_down(start, w, src, dst) if src > dst else
                                          resample.periodic_plan_up(start, w, src, dst))
        self.plan = None
        if plan is not None:
            self.plan = _lib.PeriodicPlan(plan["ratio"], plan["taps"], plan["offset"], plan["lo"], plan["hi"])
            for i, v in enumerate(plan["w"]):
                self.plan.w[i] = float(v)

    def resample_h(self, lib, src_ptr: int, dst_ptr: int, rows: int, stream: int, what: str = "resample.h"):
        """Horizontal squeeze pass u8 [rows][src] -> float [rows][dst]."""
        import ctypes as C
        if self.plan is not None and self.src > self.dst and rows >= 8 and src_ptr % 4 == 0 and dst_ptr % 16 == 0:
            _lib.check(lib.havc_resample_h_periodic(src_ptr, dst_ptr, rows, self.src, self.dst, self.start.data_ptr(), self.wt.data_ptr(),
                                                    self.taps, C.byref(self.plan), stream), what)
        else:
            _lib.check(lib.havc_resample_h(src_ptr, dst_ptr, rows, self.src, self.dst, self.start.data_ptr(), self.wt.data_ptr(), self.taps,
                                           stream), what)

    def post_horizontal(self, lib, in_ptr: int, orig_ptr, out_ptr: int, B: int, H: int, transplant: int, stream: int,
                        what: str = "post.h"):
        """Final horizontal pass float [B][3][H][src] -> u8 [B][3][H][dst] (+ luma transplant from `orig`)."""
        import ctypes as C
        if self.plan is not None and self.dst > self.src and out_ptr % 4 == 0 and (orig_ptr or 0) % 4 == 0:
            _lib.check(lib.havc_post_horizontal_periodic(in_ptr, orig_ptr, out_ptr, B, self.src, H, self.dst, self.start.data_ptr(),
                                                         self.wt.data_ptr(), self.taps, transplant, C.byref(self.plan), stream), what)
        else:
            _lib.check(lib.havc_post_horizontal(in_ptr, orig_ptr, out_ptr, B, self.src, H, self.dst, self.start.data_ptr(), self.wt.data_ptr(),
                                                self.taps, transplant, stream), what)


class _FormatIO:
    """convert_format_RGB24 / restore_format (vsdeoldify/havc_utils.py:57-237) for 8-bit YUV420 and GRAY clips, on the device and
    inside the engine's CUDA graph, so that such clips cross PCIe at 1.5 (1) bytes per pixel instead of 3:
      in : YUV420P8 -> RGB24  resize.Bicubic(format=RGB24, matrix_in=<_Matrix or 709>, range_in_s="limited", range_s="full",
                              dither_type="error_diffusion") (:133-143);  GRAY8 -> RGB24 (limited -> full, no dither, :145-151)
      out: RGB24 -> the clip's YUV format with its matrix / range and error-diffusion dither (:199-207); a GRAY clip comes back as
           YUV420P8 BT.709 (:208-222).
    zimg is restated (csrc/zimg.cu, oracle/zimg_oracle.py: parity unpinned).  Host / device layout of one batch: all Y planes
    [B][H][W], then all chroma planes [B][2][H/2][W/2], one flat byte buffer."""

    def __init__(self, fmt: str, B: int, H: int, W: int, dev, matrix: str = "709", out_limited: bool = True):
        if fmt not in ("yuv420p8", "gray8"):
            raise ValueError(f"unsupported clip format {fmt!r} (rgb24, yuv420p8, gray8)")
        if H % 2 or W % 2:
            raise ValueError("4:2:0 clips need an even width and height")
        if matrix not in ("709", "601"):
            raise ValueError(f"unsupported matrix {matrix!r} (BT.709, BT.601 / 470bg / 170m)")
        self.fmt, self.B, self.H, self.W, self.dev = fmt, B, H, W, dev
        self.lib = _lib.lib()
        self.matrix_in = {"709": 0, "601": 1}[matrix]
        self.matrix_out = self.matrix_in if fmt == "yuv420p8" else 0            # GRAY input comes back as BT.709 YUV (:216-219)
        self.out_limited = int(bool(out_limited))
        n = H * W
        self.in_bytes = B * n * 3 // 2 if fmt == "yuv420p8" else B * n
        self.out_bytes = B * n * 3 // 2
        up = lambda t: (torch.from_numpy(t[0]).to(dev), torch.from_numpy(np.ascontiguousarray(t[1])).to(dev), int(t[1].shape[1]))
        self.dh, self.dv, self.uh, self.uv = (up(t) for t in resample.chroma420_tables(W, H))
        f32 = dict(dtype=torch.float32, device=dev)
        self.s444 = torch.empty(B, 2, H, W, **f32)
        self.s_half = torch.empty(B, 2, H // 2, W, **f32)
        self.s_rgb = torch.empty(B, 3, H, W, **f32)
        self.s_q = torch.empty(B * n * 3 // 2, **f32)

    def planes(self, buf: np.ndarray, j: int, out: bool = False):
        """The plane arrays of frame j inside a flat batch buffer (views)."""
        B, H, W = self.B, self.H, self.W
        n = H * W
        y = buf[j * n:(j + 1) * n].reshape(H, W)
        if self.fmt == "gray8" and not out:
            return [y]
        c0 = B * n + 2 * j * (n // 4)
        return [y, buf[c0:c0 + n // 4].reshape(H // 2, W // 2), buf[c0 + n // 4:c0 + n // 2].reshape(H // 2, W // 2)]

    def to_rgb(self, raw: torch.Tensor, rgb: torch.Tensor, stream: int):
        B, H, W, chk, lib = self.B, self.H, self.W, _lib.check, self.lib
        if self.fmt == "gray8":
            chk(lib.havc_zimg_gray8_to_rgb(raw.data_ptr(), rgb.data_ptr(), B, H, W, 1, stream), "gray8_to_rgb")
            return
        uh, uv = self.uh, self.uv
        chk(lib.havc_zimg_yuv420p8_to_rgb(raw.data_ptr(), raw.data_ptr() + B * H * W, rgb.data_ptr(), self.s_half.data_ptr(),
                                          self.s_rgb.data_ptr(), B, H, W, uh[0].data_ptr(), uh[1].data_ptr(), uh[2], uv[0].data_ptr(),
                                          uv[1].data_ptr(), uv[2], self.matrix_in, 1, 1, stream), "yuv420p8_to_rgb")

    def from_rgb(self, rgb: torch.Tensor, raw: torch.Tensor, stream: int):
        B, H, W, chk, lib = self.B, self.H, self.W, _lib.check, self.lib
        dh, dv = self.dh, self.dv
        chk(lib.havc_zimg_rgb_to_yuv420p8(rgb.data_ptr(), raw.data_ptr(), raw.data_ptr() + B * H * W, self.s444.data_ptr(),
                                          self.s_half.data_ptr(), self.s_q.data_ptr(), B, H, W, dv[0].data_ptr(), dv[1].data_ptr(), dv[2],
                                          dh[0].data_ptr(), dh[1].data_ptr(), dh[2], self.matrix_out, self.out_limited, 1, stream),
            "rgb_to_yuv420p8")


class FormatEngine:
    """convert_format_RGB24 / restore_format (vsdeoldify/havc_utils.py:57-237) for YUV420P8 / GRAY8 clips on batches of host
    frames, outside a colorizer engine: the format glue of HAVC_stabilizer / HAVC_merge / vs_chroma_stabilizer_ex under the
    VapourSynth stand-in (under real VapourSynth those entry points call VapourSynth's own resize, like the reference)."""

    def __init__(self, fmt: str, width: int, height: int, batch: int = 8, device: str = "cuda:0", matrix: str = "709",
                 out_limited: bool = True):
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        self.B, self.H, self.W = batch, height, width
        self.io = _FormatIO(fmt, batch, height, width, self.dev, matrix, out_limited)
        u8 = dict(dtype=torch.uint8, device=self.dev)
        self.d_raw_in, self.d_raw_out = torch.zeros(self.io.in_bytes, **u8), torch.zeros(self.io.out_bytes, **u8)
        self.d_rgb = torch.zeros(batch, 3, height, width, **u8)
        self.h_raw_in, self.h_raw_out = torch.zeros(self.io.in_bytes, dtype=torch.uint8).pin_memory(), torch.zeros(self.io.out_bytes, dtype=torch.uint8).pin_memory()
        self.h_rgb = torch.zeros(batch, 3, height, width, dtype=torch.uint8).pin_memory()
        self.stream = torch.cuda.Stream(device=self.dev)

    def to_rgb(self, frames) -> np.ndarray:
        """frames: up to B source frames (indexable by plane) -> uint8 [n, 3, H, W]."""
        n = len(frames)
        buf = self.h_raw_in.numpy()
        for j, f in enumerate(frames):
            for d, p in zip(self.io.planes(buf, j), range(3)):
                np.copyto(d, np.asarray(f[p]))
        with torch.cuda.stream(self.stream):
            self.d_raw_in.copy_(self.h_raw_in, non_blocking=True)
            self.io.to_rgb(self.d_raw_in, self.d_rgb, self.stream.cuda_stream)
            self.h_rgb.copy_(self.d_rgb, non_blocking=True)
        self.stream.synchronize()
        return self.h_rgb[:n].numpy().copy()

    def from_rgb(self, rgb: np.ndarray):
        """uint8 [n <= B, 3, H, W] -> per frame [Y, U, V] plane arrays (copies)."""
        n = rgb.shape[0]
        self.h_rgb[:n].copy_(torch.from_numpy(np.ascontiguousarray(rgb)))
        with torch.cuda.stream(self.stream):
            self.d_rgb.copy_(self.h_rgb, non_blocking=True)
            self.io.from_rgb(self.d_rgb, self.d_raw_out, self.stream.cuda_stream)
            self.h_raw_out.copy_(self.d_raw_out, non_blocking=True)
        self.stream.synchronize()
        buf = self.h_raw_out.numpy()
        return [[pl.copy() for pl in self.io.planes(buf, j, out=True)] for j in range(n)]


class DeoldifyEngine:
    """DeOldify at render size S = render_factor*16 on frames of width x height, batch B.

    `sd` is the 'video' generator, which always runs (visualize.py:120).  With `sd_other` (the 'stable' or
    'artistic' generator) both run on the same input and their S x S results are mixed with
    Image.blend(other, video, video_weight) before the resize back (visualize.py:118-137)."""

    def __init__(self, sd: Dict[str, torch.Tensor], width: int, height: int, render_factor: int = 24, batch: int = 8,
                 dtype: torch.dtype = torch.float16, device: str = "cuda:0", resize_kernel: str = "spline64",
                 use_graph: bool = True, keep_taps: bool = False, debug_net_out: bool = False,
                 frame_size: Optional[int] = None, sd_other: Optional[Dict[str, torch.Tensor]] = None,
                 video_weight: float = 0.5, zhang: Optional[tuple] = None, merge: Optional[dict] = None,
                 hue_adjust: str = "none", run_deoldify: bool = True, ddtweak: Optional[dict] = None,
                 precision: Optional[str] = None, sat=(1.0, 1.0), hue=(0.0, 0.0), fmt: str = "rgb24", matrix: str = "709",
                 out_limited: bool = True):
        """sat / hue = (first clip, second clip): the vs_tweak of vs_sc_combine_models (mcomb.py:161-169; deoldify_p[2:4] and
        ddcolor_p[2:4] of HAVC_colorizer), applied AFTER the optional clip swap, to every frame.
        fmt: the clip format the host hands over / gets back - 'rgb24' (planar [B,3,H,W]), 'yuv420p8' or 'gray8' (flat batch
        buffers, see _FormatIO; converted on the device like convert_format_RGB24 / restore_format); matrix ('709' | '601') and
        out_limited (the clip's colour range) as the reference reads them from the frame props.
        zhang = (name, state_dict): the second colour model of HAVC_colorizer (vs_sc_ddcolor models 2 / 3,
        vsslib/vsmodels.py:339-344), colourising the same S x S frame; hue_adjust: its vs_sc_adjust_clip_hue string
        (vsmodels.py:361-362); merge = dict(method, weight, cmc_p, lmm_p, alm_p, crt_p, invert) for
        vs_sc_combine_models (mcomb.py:125-192).  run_deoldify=False is method 1 (second model only).  ddtweak =
        dict(bright, cont, luma_min, gamma, gamma_luma_min, gamma_alpha, gamma_min): the luma-constrained pre-tweak of the
        second model's input (vs_sc_tweak + sc_constrained_tweak, vsmodels.py:326-332) followed by vs_recover_clip_luma
        on its output (vsmodels.py:367-368)."""
        self.lib = _lib.lib()
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        self.W, self.H, self.B = width, height, batch
        # frame_size (vsdeoldify/__init__.py:2502) = min(max(ddcolor_rf, deoldify_rf)*16, width): the Spline64 squeeze size F.  The
        # DeOldify filter itself always renders at N = render_factor*16 (BaseFilter._scale_to_square, deoldify/filters.py:37-41,82-84):
        # when F != N it stretches the F x F frame to N x N with Pillow BILINEAR, runs the generator, resizes the result back to
        # F x F (_unsquare, filters.py:70-73) and only then transplants the luma (filters.py:100-110).  Same passes here.
        self.S = min(render_factor * 16, width) if frame_size is None else frame_size
        self.N = render_factor * 16 if run_deoldify else self.S                # network size
        if self.S > width or self.S < 16:
            raise ValueError(f"frame_size {self.S} must lie in [16, clip width {width}]")
        S, B, W, H = self.S, batch, width, height
        N = self.N
        self.dtype, self.hd = dtype, ops.havc_dtype(dtype)
        self.run_deoldify = run_deoldify
        self.prog = UnetProgram(sd, B, N, dtype, device=self.dev, keep_taps=keep_taps, precision=precision) if run_deoldify else None
        self.prog2 = UnetProgram(sd_other, B, N, dtype, device=self.dev, x=self.prog.x, precision=precision) \
            if (sd_other is not None and run_deoldify) else None
        self.zhang = None
        self.merge, self.hue_adjust, self.ddtweak = merge, hue_adjust, ddtweak
        self.tweak = [(float(sat[i]), float(hue[i])) for i in (0, 1)]
        self.has_tweak = any(t != (1.0, 0.0) for t in self.tweak)
        if zhang is not None or self.has_tweak:
            from .filters import FilterBank
            self.bank = FilterBank(B, S, S, self.dev)
            self.tweaked = [torch.empty(B, 3, S, S, dtype=torch.uint8, device=self.dev) for _ in range(2)]
        if zhang is not None:
            from .zhang import ZhangColorizer
            self.zhang = ZhangColorizer(zhang[1], zhang[0], B, S, dtype, device=self.dev, precision=precision)
            self.colored_b = torch.empty(B, 3, S, S, dtype=torch.uint8, device=self.dev)
            self.colored_b2 = torch.empty(B, 3, S, S, dtype=torch.uint8, device=self.dev)
            self.merged = torch.empty(B, 3, S, S, dtype=torch.uint8, device=self.dev)
        if not run_deoldify and zhang is None:
            raise ValueError("nothing to run: method 1 needs the second colour model")
        self.video_weight = float(video_weight)
        self.t_down_h = _Tables(W, S, resize_kernel, self.dev)
        self.t_down_v = _Tables(H, S, resize_kernel, self.dev)
        self.t_up_h = _Tables(S, W, resize_kernel, self.dev)
        self.t_up_v = _Tables(S, H, resize_kernel, self.dev)
        u8 = dict(dtype=torch.uint8, device=self.dev)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.n_slots = 2
        self.d_in = [torch.empty(B, 3, H, W, **u8) for _ in range(self.n_slots)]
        self.d_out = [torch.empty(B, 3, H, W, **u8) for _ in range(self.n_slots)]
        self.h_in = [torch.empty(B, 3, H, W, dtype=torch.uint8).pin_memory() for _ in range(self.n_slots)]
        self.h_out = [torch.empty(B, 3, H, W, dtype=torch.uint8).pin_memory() for _ in range(self.n_slots)]
        self.fmt = fmt
        self.io = None
        if fmt != "rgb24":      # what crosses PCIe is the clip's own format; the RGB24 buffers above stay device-side only
            self.io = _FormatIO(fmt, B, H, W, self.dev, matrix, out_limited)
            self.d_raw_in = [torch.empty(self.io.in_bytes, **u8) for _ in range(self.n_slots)]
            self.d_raw_out = [torch.empty(self.io.out_bytes, **u8) for _ in range(self.n_slots)]
            self.h_raw_in = [torch.empty(self.io.in_bytes, dtype=torch.uint8).pin_memory() for _ in range(self.n_slots)]
        self.out_shape = (B, 3, H, W) if self.io is None else (self.io.out_bytes,)
        self.tmp_down = torch.empty(B, 3, H, S, **f32)
        self.rgb_small = torch.empty(B, 3, S, S, **u8)
        self.colored = torch.empty(B, 3, S, S, **u8)
        self.colored2 = torch.empty(B, 3, S, S, **u8) if sd_other is not None else None
        self.tmp_up = torch.empty(B, 3, H, S, **f32)
        self.net_out = torch.empty(B, 3, N, N, **f32) if debug_net_out else None
        self.rescale = self.prog is not None and N != S
        if self.rescale:        # Pillow BILINEAR tables F -> N and N -> F, u8 staging images
            mk = lambda a, b: tuple(torch.from_numpy(t).to(self.dev) for t in resample.pil_tables(a, b, "bilinear"))
            self.t_sq, self.t_unsq = mk(S, N), mk(N, S)
            self.sq_h = torch.empty(B, 3, S, N, **u8)        # after the horizontal pass of the stretch
            self.sq = torch.empty(B, 3, N, N, **u8)
            self.model_img = torch.empty(B, 3, N, N, **u8)
            self.unsq_h = torch.empty(B, 3, N, S, **u8)      # after the horizontal pass back
            self.raw = torch.empty(B, 3, S, S, **u8)
        self.x_in = self.prog.x if (self.prog is not None and not self.rescale) else \
            torch.zeros(B, S, S, 8, dtype=dtype, device=self.dev)
        # per-frame scene-change gate (1 = leave uncoloured): one device / pinned pair per input slot, so that two batches
        # can be in flight (the launch list of slot s reads skip_slots[s])
        self.skip_slots = [torch.zeros(B, **u8) for _ in range(self.n_slots)]
        self.h_skip_slots = [torch.zeros(B, dtype=torch.uint8).pin_memory() for _ in range(self.n_slots)]
        self.skip, self.h_skip = self.skip_slots[0], self.h_skip_slots[0]
        self.compute = torch.cuda.Stream(device=self.dev)
        self.copy_in = torch.cuda.Stream(device=self.dev)
        self.copy_out = torch.cuda.Stream(device=self.dev)
        self.graphs: List[Optional[torch.cuda.CUDAGraph]] = [None] * self.n_slots
        self.use_graph = use_graph
        self.launches_per_batch = 0
        self._warm()

    # ---- launch list ------------------------------------------------------------------------------
    def _launch_pre(self, slot: int, stream: int):
        """clip.resize.Spline64(S, S) (vsdeoldify/__init__.py:2504) fused with the gray transform + normalisation of the filter."""
        lib, B, S, W, H, chk = self.lib, self.B, self.S, self.W, self.H, _lib.check
        td, tv = self.t_down_h, self.t_down_v
        td.resample_h(lib, self.d_in[slot].data_ptr(), self.tmp_down.data_ptr(), B * 3 * H, stream, "pre.h")
        chk(lib.havc_pre_vertical(self.tmp_down.data_ptr(), self.rgb_small.data_ptr(), self.x_in.data_ptr(), B, H, S,
                                  tv.start.data_ptr(), tv.w.data_ptr(), tv.taps, self.hd, stream), "pre.v")

    def _launch_post(self, slot: int, stream: int, result=None):
        """_clip_chroma_resize (vsdeoldify/__init__.py:3545-3554): Spline64 back to W x H + full-resolution luma transplant;
        the vertical pass runs on the S-wide image first, then the wide horizontal pass from shared memory."""
        lib, B, S, W, H, chk = self.lib, self.B, self.S, self.W, self.H, _lib.check
        uh, uv = self.t_up_h, self.t_up_v
        result = result if result is not None else self.colored
        chk(lib.havc_resample_v(result.data_ptr(), self.tmp_up.data_ptr(), B * 3, S, H, S,
                                uv.start.data_ptr(), uv.w.data_ptr(), uv.taps, stream), "post.v")
        uh.post_horizontal(lib, self.tmp_up.data_ptr(), self.d_in[slot].data_ptr(), self.d_out[slot].data_ptr(), B, H, 1, stream, "post.h")

    def _launch(self, slot: int, stream: int):
        lib, B, S, W, H = self.lib, self.B, self.S, self.W, self.H
        skip = self.skip_slots[slot]
        chk = _lib.check
        if self.io is not None:
            self.io.to_rgb(self.d_raw_in[slot], self.d_in[slot], stream)
        self._launch_pre(slot, stream)
        result = self.colored
        if self.run_deoldify and not self.rescale:
            self.prog.run(stream)
            chk(lib.havc_head(self.prog.logits.data_ptr(), 0, None,
                              self.prog.b11.data_ptr(), self.rgb_small.data_ptr(), self.colored.data_ptr(),
                              self.net_out.data_ptr() if self.net_out is not None else None, skip.data_ptr(), B, S,
                              self.hd, 1, stream),
                "head")
        elif self.run_deoldify:
            self._stretch_in(stream)
            self._filter_rescaled(self.prog, self.colored, self.net_out, skip, stream)
        if self.prog2 is not None:
            if not self.rescale:
                self.prog2.run(stream)
                chk(lib.havc_head(self.prog2.logits.data_ptr(), 0, None, self.prog2.b11.data_ptr(), self.rgb_small.data_ptr(),
                                  self.colored2.data_ptr(), None, skip.data_ptr(), B, S, self.hd, 1, stream), "head2")
            else:
                self._filter_rescaled(self.prog2, self.colored2, None, skip, stream)
            chk(lib.havc_blend_u8(self.colored2.data_ptr(), self.colored.data_ptr(), self.colored.data_ptr(),
                                  B * 3 * S * S, self.video_weight, stream), "blend")
        if self.zhang is not None:
            # second colour model on the same S x S frame, its hue adjustment, then the model merge
            src_b = self.rgb_small
            if self.ddtweak is not None:                       # pre-tweak of the second model's input
                t, tmp = self.ddtweak, self.bank.tmp
                if self.bank.image_tweak(src_b, tmp[0], cont=t["cont"], bright=t["bright"], stream=stream):
                    self.bank.select_frames(tmp[0], src_b, skip, stream)
                    src_b = tmp[0]
                self.bank.luma_adjusted_levels(src_b, tmp[1], t["luma_min"], t["gamma"], t["gamma_luma_min"], t["gamma_alpha"],
                                               t["gamma_min"], stream=stream)
                self.bank.select_frames(tmp[1], src_b, skip, stream)
                src_b = tmp[1]
            self.zhang.run(src_b, self.colored_b, stream)
            clipb = self.colored_b
            if self.bank.adjust_hue_range(clipb, self.colored_b2, self.hue_adjust, stream):
                clipb = self.colored_b2
            if self.ddtweak is not None:                       # vs_recover_clip_luma(clip, clipb_rgb)
                other = self.colored_b if clipb is self.colored_b2 else self.colored_b2
                chk(lib.havc_chroma_post_process(clipb.data_ptr(), self.rgb_small.data_ptr(), other.data_ptr(), B, S, S, stream),
                    "recover_luma")
                clipb = other
            self.bank.select_frames(clipb, self.rgb_small, skip, stream)        # scene-change gate of the 2nd model
            if not self.run_deoldify:                                           # mcomb.py:166-169: only clipb, its own tweak
                result = clipb
                if self.bank.vs_tweak(clipb, self.tweaked[1], hue=self.tweak[1][1], sat=self.tweak[1][0], stream=stream):
                    result = self.tweaked[1]
            else:
                m = self.merge
                a, b = (clipb, self.colored) if m.get("invert") else (self.colored, clipb)
                # vs_tweak of both clips after the swap (mcomb.py:154-169): sat[0] / hue[0] go to whatever is clip a now
                if self.bank.vs_tweak(a, self.tweaked[0], hue=self.tweak[0][1], sat=self.tweak[0][0], stream=stream):
                    a = self.tweaked[0]
                if self.bank.vs_tweak(b, self.tweaked[1], hue=self.tweak[1][1], sat=self.tweak[1][0], stream=stream):
                    b = self.tweaked[1]
                self.bank.combine(a, b, self.merged, m["method"], m["weight"], m["cmc_p"], m["lmm_p"], m["alm_p"], m["crt_p"],
                                  stream=stream)
                self.bank.select_frames(self.merged, a, skip, stream)           # merge selectors return f[0].copy()
                result = self.merged
        elif self.has_tweak:                                                    # method 0: mcomb.py:161-164, clipa alone
            if self.bank.vs_tweak(self.colored, self.tweaked[0], hue=self.tweak[0][1], sat=self.tweak[0][0], stream=stream):
                result = self.tweaked[0]
        self._launch_post(slot, stream, result)
        if self.io is not None:
            self.io.from_rgb(self.d_out[slot], self.d_raw_out[slot], stream)

    # ---- frame_size != render_factor*16: the filter's own Pillow BILINEAR stretch around the generator ---------------
    def _pil(self, src, tmp, dst, Hin, Win, out_size, tabs, stream):
        """Pillow Image.resize of square u8 planes: horizontal pass, u8 intermediate, vertical pass (ImagingResample)."""
        lib, B, chk = self.lib, self.B, _lib.check
        chk(lib.havc_pil_resample_u8(src.data_ptr(), tmp.data_ptr(), B * 3, Hin, Win, out_size, 1, tabs[0].data_ptr(),
                                     tabs[1].data_ptr(), tabs[1].shape[1], stream), "pil.h")
        chk(lib.havc_pil_resample_u8(tmp.data_ptr(), dst.data_ptr(), B * 3, Hin, out_size, out_size, 0, tabs[0].data_ptr(),
                                     tabs[1].data_ptr(), tabs[1].shape[1], stream), "pil.v")

    def _stretch_in(self, stream):
        """_scale_to_square (filters.py:37-41) + _transform + normalise: rgb_small [F x F] -> x [N x N]."""
        S, N = self.S, self.N
        self._pil(self.rgb_small, self.sq_h, self.sq, S, S, N, self.t_sq, stream)
        _lib.check(self.lib.havc_gray_normalize(self.sq.data_ptr(), self.prog.x.data_ptr(), self.B, N * N, self.hd, stream),
                   "gray_normalize")

    def _filter_rescaled(self, prog, colored, net_out, skip, stream):
        """generator at N x N -> u8 image -> _unsquare to F x F (filters.py:70-73) -> _post_process against rgb_small
        (filters.py:100-110); scene-change-skipped frames come back as rgb_small (vsslib/vsmodels.py:221-224)."""
        lib, B, S, N, chk = self.lib, self.B, self.S, self.N, _lib.check
        prog.run(stream)
        chk(lib.havc_head(prog.logits.data_ptr(), 0, None, prog.b11.data_ptr(), None, self.model_img.data_ptr(),
                          net_out.data_ptr() if net_out is not None else None, None, B, N, self.hd, 0, stream), "head")
        self._pil(self.model_img, self.unsq_h, self.raw, N, N, S, self.t_unsq, stream)
        chk(lib.havc_chroma_post_process(self.raw.data_ptr(), self.rgb_small.data_ptr(), colored.data_ptr(), B, S, S, stream),
            "post_process")
        chk(lib.havc_select_frames(colored.data_ptr(), self.rgb_small.data_ptr(), skip.data_ptr(), B, 3 * S * S, stream), "skip")

    def _warm(self):
        with torch.cuda.stream(self.compute):
            for s in range(self.n_slots):
                self.d_in[s].zero_()
                if self.io is not None:
                    self.d_raw_in[s].fill_(128)
            n0 = self.lib.havc_launch_count()
            self._launch(0, self.compute.cuda_stream)
            self.launches_per_batch = int(self.lib.havc_launch_count() - n0)
        self.compute.synchronize()
        if self.use_graph:
            for s in range(self.n_slots):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.compute):
                    self._launch(s, torch.cuda.current_stream().cuda_stream)
                self.graphs[s] = g
            self.compute.synchronize()

    def run_slot(self, slot: int):
        """Enqueue the device work for the frames resident in d_in[slot] on the compute stream."""
        with torch.cuda.stream(self.compute):
            if self.use_graph:
                self.graphs[slot].replay()
            else:
                self._launch(slot, self.compute.cuda_stream)

    # ---- synchronous convenience API (tests) ---------------------------------------------------------
    def colorize_batch(self, frames: np.ndarray, skip: Optional[np.ndarray] = None) -> np.ndarray:
        """frames: uint8 [n<=B, 3, H, W] planar RGB (host).  Returns uint8 [n, 3, H, W].
        skip[i] = True leaves frame i uncoloured (scene-change gating, vsslib/vsmodels.py:221-224): it still goes
        through the squeeze / un-squeeze / luma transplant exactly like a frame the reference's selector returned
        unchanged."""
        if self.io is not None:
            return self._colorize_batch_raw(frames, skip)
        n = frames.shape[0]
        assert n <= self.B and frames.shape[1:] == (3, self.H, self.W) and frames.dtype == np.uint8
        self.h_in[0][:n].copy_(torch.from_numpy(np.ascontiguousarray(frames)))
        self.h_skip.zero_()
        if skip is not None:
            self.h_skip[:n].copy_(torch.from_numpy(np.asarray(skip, dtype=np.uint8)))
        with torch.cuda.stream(self.compute):
            self.skip.copy_(self.h_skip, non_blocking=True)
            self.d_in[0].copy_(self.h_in[0], non_blocking=True)
        self.run_slot(0)
        with torch.cuda.stream(self.compute):
            self.h_out[0].copy_(self.d_out[0], non_blocking=True)
        self.compute.synchronize()
        return self.h_out[0][:n].numpy().copy()

    def _colorize_batch_raw(self, planes_per_frame, skip=None):
        """Synchronous path for YUV420P8 / GRAY8 clips: planes_per_frame = list of per-frame plane lists ([Y, U, V] or [Y]);
        returns a list of [Y, U, V] uint8 arrays per frame (the restored YUV420P8 frames)."""
        n = len(planes_per_frame)
        assert n <= self.B
        buf = self.h_raw_in[0].numpy()
        for j, pl in enumerate(planes_per_frame):
            for dst, src in zip(self.io.planes(buf, j), pl):
                np.copyto(dst, src)
        self.h_skip.zero_()
        if skip is not None:
            self.h_skip[:n].copy_(torch.from_numpy(np.asarray(skip, dtype=np.uint8)))
        with torch.cuda.stream(self.compute):
            self.skip.copy_(self.h_skip, non_blocking=True)
            self.d_raw_in[0].copy_(self.h_raw_in[0], non_blocking=True)
        self.run_slot(0)
        with torch.cuda.stream(self.compute):
            out = self.d_raw_out[0].cpu()
        self.compute.synchronize()
        arr = out.numpy()
        return [[p.copy() for p in self.io.planes(arr, j, out=True)] for j in range(n)]

    # ---- plane views of host batch buffers (clip adapters) -------------------------------------------------
    def in_planes(self, buf: np.ndarray, j: int):
        return [buf[j, p] for p in range(3)] if self.io is None else self.io.planes(buf, j)

    def out_planes(self, arr: np.ndarray, j: int):
        return [arr[j, p] for p in range(3)] if self.io is None else self.io.planes(arr, j, out=True)

    # ---- asynchronous batch API (clip rendering with read-ahead) ----------------------------------------
    def _async_state(self):
        if not hasattr(self, "_ev"):
            self._ev = [dict(inp=torch.cuda.Event(), done=torch.cuda.Event(), out=torch.cuda.Event(), busy=False, used=False)
                        for _ in range(self.n_slots)]
            self._next = 0
        return self._ev[self._next]

    def next_input(self) -> np.ndarray:
        """The pinned host buffer [B, 3, H, W] of the slot the next submit() will use: a caller that assembles its batch
        plane by plane can write straight into it and call submit(None, n=...) - one host copy per frame instead of three."""
        if self._async_state()["busy"]:
            raise RuntimeError("DeoldifyEngine.next_input: every input slot is in flight; collect() a ticket first")
        return (self.h_in if self.io is None else self.h_raw_in)[self._next].numpy()

    def submit(self, frames: Optional[np.ndarray], skip: Optional[np.ndarray] = None, n: Optional[int] = None):
        """Enqueue one batch (uint8 [n<=B, 3, H, W] host frames, or None when the caller filled next_input()[:n]) on the next
        input slot: H2D, the CUDA graph and D2H run on the copy / compute streams; returns a ticket for collect().  At most
        n_slots batches may be outstanding."""
        ev = self._async_state()
        s = self._next
        if ev["busy"]:
            raise RuntimeError("DeoldifyEngine.submit: every input slot is in flight; collect() a ticket first")
        self._next = (s + 1) % self.n_slots
        if frames is not None:
            n = frames.shape[0]
            assert n <= self.B and frames.shape[1:] == (3, self.H, self.W) and frames.dtype == np.uint8
            self.h_in[s][:n].copy_(torch.from_numpy(np.ascontiguousarray(frames)))
        assert n is not None and 0 < n <= self.B
        self.h_skip_slots[s].zero_()
        if skip is not None:
            self.h_skip_slots[s][:n].copy_(torch.from_numpy(np.asarray(skip, dtype=np.uint8)))
        with torch.cuda.stream(self.copy_in):
            if ev["used"]:
                self.copy_in.wait_event(ev["done"])          # the previous graph on this slot has consumed d_in / skip
            if self.io is None:
                self.d_in[s].copy_(self.h_in[s], non_blocking=True)
            else:
                self.d_raw_in[s].copy_(self.h_raw_in[s], non_blocking=True)
            self.skip_slots[s].copy_(self.h_skip_slots[s], non_blocking=True)
            ev["inp"].record(self.copy_in)
        self.compute.wait_event(ev["inp"])
        if ev["used"]:
            self.compute.wait_event(ev["out"])               # d_out of this slot has been downloaded
        self.run_slot(s)
        ev["done"].record(self.compute)
        ob = self._acquire_out()                             # a pinned result buffer no delivered frame references any more
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(ev["done"])
            ob[0].copy_(self.d_out[s] if self.io is None else self.d_raw_out[s], non_blocking=True)
            ev["out"].record(self.copy_out)
        ev["busy"], ev["used"], ev["ob"] = True, True, ob
        return (s, n, ob)

    def prepare_async(self, buffers: int = 8):
        """Pin the result-buffer pool of the submit() / collect_view() API now (clip adapters call this when they are built): eight
        buffers cover two batches in flight plus the batches an adapter keeps cached around its cursor.  Pinning 200 MB takes
        ~0.14 s and stalls the device queues of every GPU of the process (measured), so it must not happen inside the steady
        state - nor lazily inside the first frames of a multi-GPU clip."""
        if not hasattr(self, "_out_pool"):
            self._out_pool = []                              # separate from h_out (the synchronous / streaming APIs own those)
        while len(self._out_pool) < buffers:
            t = torch.empty(*self.out_shape, dtype=torch.uint8).pin_memory()
            self._out_pool.append((t, t.numpy()))

    def _acquire_out(self):
        """(pinned tensor, its numpy view) from the result-buffer pool: the first one whose numpy view is referenced by nobody
        else (frames handed out by collect_view() are views of it and keep it busy through `.base`), else a new one."""
        import sys
        if not hasattr(self, "_out_pool"):
            self.prepare_async()
        for ob in self._out_pool:
            if sys.getrefcount(ob[1]) <= 2 and not any(ob is e.get("ob") for e in getattr(self, "_ev", [])):
                return ob
        t = torch.empty(*self.out_shape, dtype=torch.uint8).pin_memory()
        ob = (t, t.numpy())
        self._out_pool.append(ob)
        return ob

    def collect_view(self, ticket) -> np.ndarray:
        """Wait for a submitted batch and return its uint8 [n, 3, H, W] result as a VIEW of the pinned buffer it was downloaded
        into: no host copy.  The buffer returns to the pool when the last view of it is dropped."""
        s, n, ob = ticket
        ev = self._ev[s]
        ev["out"].synchronize()
        ev["busy"] = False
        ev["ob"] = None
        return ob[1][:n] if self.io is None else ob[1]

    def collect(self, ticket, out: Optional[np.ndarray] = None, pool=None) -> np.ndarray:
        """Wait for a submitted batch and return its uint8 [n, 3, H, W] result: a fresh copy, or `out[:n]` when the caller
        supplies a (recycled) array - a fresh 200 MB allocation costs ten times the copy itself in page faults.  `pool`
        (a concurrent.futures executor) spreads the per-frame copies over host threads (numpy releases the GIL)."""
        s, n, ob = ticket
        ev = self._ev[s]
        ev["out"].synchronize()
        ev["busy"] = False
        ev["ob"] = None
        src = ob[1][:n]
        if out is None:
            return src.copy()
        dst = out[:n]
        if pool is None:
            np.copyto(dst, src)
        else:
            list(pool.map(lambda i: np.copyto(dst[i], src[i]), range(n)))
        return dst

    # ---- pipelined API (bench e2e / clip rendering) ----------------------------------------------------
    def colorize_stream(self, batches, on_result):
        """batches: iterable of uint8 [B,3,H,W] host arrays or (pinned) torch tensors (a full batch each); pinned tensors
        are copied to the device directly and must stay untouched until their result is delivered; on_result(i, out) is called
        in order with a view of the pinned output buffer that is valid ONLY during the callback (the download of a later batch
        reuses the slot as soon as the callback returns: copy what you keep).
        H2D of batch i+1 and D2H of batch i-1 overlap the compute of batch i."""
        ev_in = [torch.cuda.Event() for _ in range(self.n_slots)]
        ev_done = [torch.cuda.Event() for _ in range(self.n_slots)]
        ev_out = [torch.cuda.Event() for _ in range(self.n_slots)]
        ev_free = [None] * self.n_slots
        pending = []
        i = -1
        for i, fr in enumerate(batches):
            s = i % self.n_slots
            if len(pending) >= self.n_slots:           # slot about to be reused: deliver its previous result
                j, sj = pending.pop(0)
                ev_out[sj].synchronize()
                on_result(j, self.h_out[sj].numpy())
            if isinstance(fr, torch.Tensor) and fr.is_pinned():
                src = fr                                   # caller-owned pinned memory: DMA straight from it
            else:                                          # pageable input: stage through our pinned buffer
                self.h_in[s].copy_(torch.from_numpy(fr) if isinstance(fr, np.ndarray) else fr)
                src = self.h_in[s]
            with torch.cuda.stream(self.copy_in):
                if ev_free[s] is not None:
                    self.copy_in.wait_event(ev_free[s])   # previous compute on this slot has consumed d_in
                self.d_in[s].copy_(src, non_blocking=True)
                ev_in[s].record(self.copy_in)
            self.compute.wait_event(ev_in[s])
            self.compute.wait_event(ev_out[s]) if i >= self.n_slots else None
            self.run_slot(s)
            ev_done[s].record(self.compute)
            ev_free[s] = ev_done[s]
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(ev_done[s])
                self.h_out[s].copy_(self.d_out[s], non_blocking=True)
                ev_out[s].record(self.copy_out)
            pending.append((i, s))
        for j, sj in pending:
            ev_out[sj].synchronize()
            on_result(j, self.h_out[sj].numpy())
        return i + 1


class ImageRenderEngine:
    """vsdeoldify.deoldify.visualize.ModelImageRender on the GPU (deoldify/visualize.py:41-137): the direct, per-image
    path (BASELINE cfg1) in which the FILTER squeezes the W x H image to S x S with Pillow BILINEAR
    (BaseFilter._scale_to_square, filters.py:37-41), runs the generator, resizes the result back with Pillow BILINEAR
    (_unsquare, filters.py:70-73) and only then transplants the original luma at full resolution
    (ColorizerFilter._post_process, filters.py:100-110).  The video generator always runs; 'stable' / 'artistic' run a
    second generator through the same steps and mix the two full-size results with Image.blend (visualize.py:118-137).
    Both resizes are the bit-exact integer Pillow passes (havc_pil_resample_u8)."""

    def __init__(self, sd: Dict[str, torch.Tensor], width: int, height: int, render_factor: int = 24, batch: int = 1,
                 dtype: torch.dtype = torch.float16, device: str = "cuda:0", sd_other: Optional[Dict[str, torch.Tensor]] = None,
                 video_weight: float = 0.5, precision: Optional[str] = None):
        self.lib = _lib.lib()
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        self.W, self.H, self.B, self.S = width, height, batch, render_factor * 16
        S, B, W, H = self.S, batch, width, height
        self.hd = ops.havc_dtype(dtype)
        self.prog = UnetProgram(sd, B, S, dtype, device=self.dev, precision=precision)
        self.prog2 = UnetProgram(sd_other, B, S, dtype, device=self.dev, x=self.prog.x, precision=precision) if sd_other is not None else None
        self.video_weight = float(video_weight)
        mk = lambda a, b: tuple(torch.from_numpy(t).to(self.dev) for t in resample.pil_tables(a, b, "bilinear"))
        self.t_dw, self.t_dh = (mk(W, S) if W != S else None), (mk(H, S) if H != S else None)      # squeeze
        self.t_uw, self.t_uh = (mk(S, W) if W != S else None), (mk(S, H) if H != S else None)      # back
        u8 = dict(dtype=torch.uint8, device=self.dev)
        self.d_in = torch.empty(B, 3, H, W, **u8)
        self.d_out = torch.empty(B, 3, H, W, **u8)
        self.d_out2 = torch.empty(B, 3, H, W, **u8) if sd_other is not None else None
        self.sq_h = torch.empty(B, 3, H, S, **u8)          # after the horizontal squeeze pass
        self.sq = torch.empty(B, 3, S, S, **u8)
        self.model_img = torch.empty(B, 3, S, S, **u8)
        self.up_h = torch.empty(B, 3, S, W, **u8)          # after the horizontal pass back
        self.raw = torch.empty(B, 3, H, W, **u8)
        self.stream = torch.cuda.Stream(device=self.dev)

    def _pil_resize(self, src, tmp, dst, Hin, Win, Hout, Wout, tw, th, st):
        lib, B, chk = self.lib, self.B, _lib.check
        cur, h, w = src, Hin, Win
        if tw is not None:                                 # Pillow: horizontal pass first, each pass only if the size changes
            target = tmp if th is not None else dst
            chk(lib.havc_pil_resample_u8(cur.data_ptr(), target.data_ptr(), B * 3, h, w, Wout, 1, tw[0].data_ptr(), tw[1].data_ptr(),
                                         tw[1].shape[1], st), "pil.h")
            cur, w = target, Wout
        if th is not None:
            chk(lib.havc_pil_resample_u8(cur.data_ptr(), dst.data_ptr(), B * 3, h, w, Hout, 0, th[0].data_ptr(), th[1].data_ptr(),
                                         th[1].shape[1], st), "pil.v")
            cur = dst
        return cur

    def _filter(self, prog, out, st):
        lib, B, S, W, H, chk = self.lib, self.B, self.S, self.W, self.H, _lib.check
        prog.run(st)
        chk(lib.havc_head(prog.logits.data_ptr(), 0, None, prog.b11.data_ptr(), None, self.model_img.data_ptr(), None, None, B, S,
                          self.hd, 0, st), "head")
        raw = self._pil_resize(self.model_img, self.up_h, self.raw, S, S, H, W, self.t_uw, self.t_uh, st)
        chk(lib.havc_chroma_post_process(raw.data_ptr(), self.d_in.data_ptr(), out.data_ptr(), B, H, W, st), "post_process")

    def render_batch(self, frames: np.ndarray) -> np.ndarray:
        """frames: uint8 [n<=B, 3, H, W] planar RGB -> uint8 [n, 3, H, W] (get_transformed_image per image)."""
        n = frames.shape[0]
        assert n <= self.B and frames.shape[1:] == (3, self.H, self.W) and frames.dtype == np.uint8
        lib, B, S, W, H, chk = self.lib, self.B, self.S, self.W, self.H, _lib.check
        with torch.cuda.stream(self.stream):
            st = self.stream.cuda_stream
            self.d_in[:n].copy_(torch.from_numpy(np.ascontiguousarray(frames)), non_blocking=False)
            sq = self._pil_resize(self.d_in, self.sq_h, self.sq, H, W, S, S, self.t_dw, self.t_dh, st)
            chk(lib.havc_gray_normalize(sq.data_ptr(), self.prog.x.data_ptr(), B, S * S, self.hd, st), "gray_normalize")
            self._filter(self.prog, self.d_out, st)
            if self.prog2 is not None:
                self._filter(self.prog2, self.d_out2, st)
                chk(lib.havc_blend_u8(self.d_out2.data_ptr(), self.d_out.data_ptr(), self.d_out.data_ptr(), self.d_out.numel(),
                                      self.video_weight, st), "blend")
            out = self.d_out.cpu()
        self.stream.synchronize()
        return out[:n].numpy().copy()
