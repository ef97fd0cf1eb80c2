"""Zhang et al. colorizers (eccv16 / siggraph17) as programs of libhavc_b200 launches.

Reference: vsdeoldify/colorization/colorizers/eccv16.py:9-98 (ECCVGenerator), siggraph17.py:7-161
(SIGGRAPHGenerator with input_B = mask_B = 0), base_color.py:5-23, and the per-frame wrapper
ModelColorization.colorize_frame (colorization/__init__.py:76-95) with preprocess_img / postprocess_tens
(colorizers/util.py:21-55).

Per batch of S x S RGB frames (planar u8, on the device):
  Pillow BICUBIC resize to 256 x 256 (two integer passes, u8 intermediate)  ->  L of the resized image, (L-50)/100  ->
  the network (implicit-GEMM convs with bias/ReLU/BatchNorm epilogues; stride-2 convs through phase-split inputs;
  dilated convs as shifted TMA boxes; ConvTranspose2d(4, 2, 1) as four 2x2-tap sub-pixel launches; the siggraph17
  shortcut sums as epilogue residuals)  ->  ab at 256 x 256  ->  bilinear resize to S x S, concatenation with the L of
  the S x S frame, LAB -> RGB in float64, uint8 truncation.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _lib, ops, resample
from .ops import Pair, chan_storage, pad_to
from .unet import LaunchProgram, bn_affine

SD = Dict[str, torch.Tensor]
NET_SIZE = 256      # ModelColorization.colorize_frame always resizes to 256 x 256 (colorization/__init__.py:80-82)

# (cout, stride, dilation) per conv of a block; a BatchNorm closes each block (eccv16.py:13-72, siggraph17.py:11-60)
_ECCV = {"model1": [(64, 1, 1), (64, 2, 1)], "model2": [(128, 1, 1), (128, 2, 1)],
         "model3": [(256, 1, 1), (256, 1, 1), (256, 2, 1)], "model4": [(512, 1, 1)] * 3, "model5": [(512, 1, 2)] * 3,
         "model6": [(512, 1, 2)] * 3, "model7": [(512, 1, 1)] * 3}
_SIG = {"model1": [(64, 1, 1), (64, 1, 1)], "model2": [(128, 1, 1), (128, 1, 1)], "model3": [(256, 1, 1)] * 3,
        "model4": [(512, 1, 1)] * 3, "model5": [(512, 1, 2)] * 3, "model6": [(512, 1, 2)] * 3, "model7": [(512, 1, 1)] * 3}


class ZhangProgram(LaunchProgram):
    """Launch list of one Zhang generator for a fixed batch: x [B,256,256,8] (channel 0 = normalised L) -> ab
    float32 [B,256,256,2] (already multiplied by ab_norm = 110)."""

    # blocks that run split-precision (hi + lo operands) unless the policy is 'fast': the first four blocks inject > 90 % of the
    # 16-bit storage error of these BatchNorm-only stacks (nothing damps an early perturbation) for < 25 % of the FLOPs
    X3_BLOCKS = ("model1", "model2", "model3", "model4")

    def __init__(self, sd: SD, name: str, batch: int, dtype: torch.dtype = torch.float16, device="cuda", keep_taps: bool = False,
                 size: int = NET_SIZE, precision: Optional[str] = None):
        assert name in ("eccv16", "siggraph17")
        if size % 8 != 0:
            raise ValueError("the Zhang networks need an input size that is a multiple of 8")
        super().__init__(sd, batch, size, dtype, device, keep_taps, precision)
        self.x3_early = self.precision != "fast"
        self.name = name
        self.x = self.buf(batch, size, size, 8, zero=True)
        self.ab = self.buf(batch, size, size, 2, dtype=torch.float32)
        (self._build_eccv16 if name == "eccv16" else self._build_siggraph17)()

    # ---- building blocks ------------------------------------------------------------------------------------
    def _first_conv(self, name: str, w: torch.Tensor, b: torch.Tensor, relu=True):
        """3x3 conv on the single L channel (the ab / mask channels of siggraph17 are identically zero) as an im2col
        (K = 3 rows x 8) + GEMM."""
        B, S, lib, hd = self.B, self.S, self.lib, self.hd
        Kp = 24
        x3 = self.x3_early
        col = Pair(self.buf(2, B, S, S, Kp, zero=True)) if x3 else self.buf(B, S, S, Kp, zero=True)
        xp = self.x.data_ptr()
        # split precision: channel 4 of x holds the lo part of the normalised L (havc_zhang_pre); a second im2col gathers it
        jobs = [(0, col.hi.data_ptr()), (4, col.lo.data_ptr())] if x3 else [(0, col.data_ptr())]

        def im2col(stream):
            for c0, cp in jobs:
                _lib.check(lib.havc_im2col_small(xp, cp, B, S, S, 8, c0, 1, 3, 1, 1, Kp, hd, stream), name + ".im2col")
        self.aux(name + ".im2col", im2col, nbytes=2.0 * len(jobs) * B * S * S * (8 + Kp))
        cout = w.shape[0]
        wk = torch.zeros(cout, 3, 8)
        wk[:, :, :3] = w[:, 0]                                     # k = kh*8 + kw
        y = self.conv(name, col, wk.reshape(cout, Kp, 1, 1), bias=b, relu1=relu, x3=x3)
        self.ops[-1].flops = 2.0 * B * S * S * cout * 9 * w.shape[1]
        return y

    def _block(self, name: str, spec, x, first_done=False, subsample_in=False):
        """conv -> ReLU (x n) -> BatchNorm; the BatchNorm is the scale/shift of the last conv's epilogue."""
        sd = self.sd
        x3 = self.x3_early and name in self.X3_BLOCKS
        if subsample_in:                                           # siggraph17: conv on x[:, :, ::2, ::2]
            ph = self.phase_split(name + ".subsample", x, 1)
            x = ph.sub(0) if isinstance(ph, Pair) else ph[0]
        for i, (cout, stride, dil) in enumerate(spec):
            p = f"{name}.{2 * i}"
            last = i == len(spec) - 1
            sc = sh = None
            if last:
                sc, sh = bn_affine(sd, f"{name}.{2 * len(spec)}")
            w, b = sd[p + ".weight"].float(), sd[p + ".bias"].float()
            if i == 0 and first_done:
                x = self._first_conv(p, w, b)
                continue
            if stride == 2:
                ph = self.phase_split(p + ".split", x, 4)
                x = self.conv(p, ph, w, ks=3, stride=2, bias=b, relu1=True, scale=sc, shift=sh, x3=x3)
            else:
                x = self.conv(p, x, w, ks=3, dilation=dil, bias=b, relu1=True, scale=sc, shift=sh, x3=x3)
        self.tap(name, x)
        return x

    def _conv_transpose(self, name: str, x, wt: torch.Tensor, b: torch.Tensor, residual: Optional[torch.Tensor] = None):
        """nn.ConvTranspose2d(cin, cout, 4, stride 2, padding 1) (+ residual) + ReLU as four sub-pixel launches:
        out[2m+a, 2n+c] = sum over 2x2 taps; a = 0 takes kernel rows (1, 3) at input rows (m, m-1), a = 1 takes
        (0, 2) at (m+1, m); columns alike.  wt: [cin, cout, 4, 4]."""
        B, H, W, _ = x.shape
        cin, cout = wt.shape[0], wt.shape[1]
        w16 = wt.permute(1, 0, 2, 3).contiguous()                  # [cout, cin, ky, kx]: 16 weight taps ky*4+kx
        out = self.buf(B, 2 * H, 2 * W, chan_storage(cout), zero=True)
        rows = {0: [(0, 1), (-1, 3)], 1: [(1, 0), (0, 2)]}         # (input offset, kernel index) per output parity
        for a in (0, 1):
            for c in (0, 1):
                taps = [(dh, dw, 0, ky * 4 + kx) for dh, ky in rows[a] for dw, kx in rows[c]]
                res = residual[:, a::2, c::2, :] if residual is not None else None
                self.conv(f"{name}.p{a}{c}", x, w16, ks=4, taps=taps, phase=(2, a, c), out=out, bias=b, residual=res,
                          relu2=True, out_c=cout, flops=2.0 * B * H * W * cout * cin * 4)
        return out

    # ---- eccv16 ------------------------------------------------------------------------------------------------
    def _build_eccv16(self):
        sd, B, lib = self.sd, self.B, self.lib
        x = self.x
        for bi, (name, spec) in enumerate(_ECCV.items()):
            x = self._block(name, spec, x, first_done=(bi == 0))
        x = self._conv_transpose("model8.0", x, sd["model8.0.weight"].float(), sd["model8.0.bias"].float())
        x = self.conv("model8.2", x, sd["model8.2.weight"].float(), ks=3, bias=sd["model8.2.bias"].float(), relu1=True)
        x = self.conv("model8.4", x, sd["model8.4.weight"].float(), ks=3, bias=sd["model8.4.bias"].float(), relu1=True)
        logits = self.conv("model8.6", x, sd["model8.6.weight"].float(), bias=sd["model8.6.bias"].float(), out_dtype=torch.float32)
        self.tap("logits", logits)
        n_cls = sd["model8.6.weight"].shape[0]
        w_out = self.dev_f32(sd["model_out.weight"].float().reshape(2, n_cls))
        h = logits.shape[1]
        reg = self.buf(B, h, h, 2, dtype=torch.float32)
        lp, ld, wp, rp, npx = logits.data_ptr(), logits.shape[-1], w_out.data_ptr(), reg.data_ptr(), B * h * h

        def head(stream):
            _lib.check(lib.havc_eccv_head(lp, ld, n_cls, wp, rp, npx, stream), "eccv.softmax_out")
        self.aux("eccv.softmax_out", head, nbytes=4.0 * npx * (ld + 2))
        self.head_flops = 2.0 * npx * n_cls * 2
        ap, S = self.ab.data_ptr(), self.S

        def up4(stream):                                          # nn.Upsample(scale_factor=4, bilinear) * ab_norm
            _lib.check(lib.havc_bilinear_ab(rp, ap, B, h, h, S, S, 110.0, stream), "eccv.upsample4")
        self.aux("eccv.upsample4", up4, nbytes=8.0 * B * (h * h + S * S))
        self.tap("out_reg", reg)

    # ---- siggraph17 ------------------------------------------------------------------------------------------------
    def _build_siggraph17(self):
        sd, B, lib, S = self.sd, self.B, self.lib, self.S
        g = lambda p: (sd[p + ".weight"].float(), sd[p + ".bias"].float())
        c1 = self._block("model1", _SIG["model1"], self.x, first_done=True)
        c2 = self._block("model2", _SIG["model2"], c1, subsample_in=True)
        c3 = self._block("model3", _SIG["model3"], c2, subsample_in=True)
        x = self._block("model4", _SIG["model4"], c3, subsample_in=True)
        for name in ("model5", "model6", "model7"):
            x = self._block(name, _SIG[name], x)
        # conv8_up = model8up(conv7_3) + model3short8(conv3_3); model8 = ReLU, conv, ReLU, conv, ReLU, BN
        w, b = g("model3short8.0")
        short = self.conv("model3short8", c3, w, ks=3, bias=b)
        w, b = g("model8up.0")
        y = self._conv_transpose("model8up", x, w, b, residual=short)
        w, b = g("model8.1")
        y = self.conv("model8.1", y, w, ks=3, bias=b, relu1=True)
        w, b = g("model8.3")
        sc, sh = bn_affine(sd, "model8.5")
        c8 = self.conv("model8.3", y, w, ks=3, bias=b, relu1=True, scale=sc, shift=sh)
        self.tap("model8", c8)
        w, b = g("model2short9.0")
        short = self.conv("model2short9", c2, w, ks=3, bias=b)
        w, b = g("model9up.0")
        y = self._conv_transpose("model9up", c8, w, b, residual=short)
        w, b = g("model9.1")
        sc, sh = bn_affine(sd, "model9.3")
        c9 = self.conv("model9.1", y, w, ks=3, bias=b, relu1=True, scale=sc, shift=sh)
        self.tap("model9", c9)
        w, b = g("model1short10.0")
        short = self.conv("model1short10", c1, w, ks=3, bias=b)
        w, b = g("model10up.0")
        y = self._conv_transpose("model10up", c9, w, b, residual=short)
        # model10: conv + LeakyReLU(0.2); model_out: 1x1 conv 128 -> 2 fused into the same epilogue, then tanh * 110
        w, b = g("model10.1")
        wo, bo = g("model_out.0")
        n = w.shape[0]
        hw = torch.zeros(3, pad_to(n, 16))
        hw[:2, :n] = wo.reshape(2, n)
        self.head_w = self.dev_f32(hw)
        self.head_b = self.dev_f32(bo)
        self.head_out = self.buf(B, S, S, 4, dtype=torch.float32, zero=True)
        self.conv("model10.1+out", y, w, ks=3, bias=b, relu1=True, leaky1=0.2, head_w=self.head_w, head_out=self.head_out,
                  flops=2.0 * B * S * S * n * (w.shape[1] * 9 + 2))
        hp, bp, ap, npx = self.head_out.data_ptr(), self.head_b.data_ptr(), self.ab.data_ptr(), B * S * S

        def tanh(stream):
            _lib.check(lib.havc_zhang_tanh(hp, bp, ap, npx, 110.0, stream), "siggraph.tanh")
        self.aux("siggraph.tanh", tanh, nbytes=24.0 * npx)


class ZhangColorizer:
    """ModelColorization.colorize_frame on a device batch: planar u8 RGB [B,3,S,S] in -> planar u8 RGB [B,3,S,S] out."""

    def __init__(self, sd: SD, name: str, batch: int, size: int, dtype=torch.float16, device="cuda", keep_taps=False,
                 precision: Optional[str] = None):
        self.B, self.S, self.dev = batch, size, torch.device(device)
        self.prog = ZhangProgram(sd, name, batch, dtype, device=self.dev, keep_taps=keep_taps, precision=precision)
        self.lib, self.hd = self.prog.lib, self.prog.hd
        u8 = dict(dtype=torch.uint8, device=self.dev)
        self.resize = size != NET_SIZE
        if self.resize:
            self.tab_h = _PilTable(size, NET_SIZE, "bicubic", self.dev)
            self.tab_v = self.tab_h
            self.tmp = torch.empty(batch, 3, size, NET_SIZE, **u8)           # after the horizontal pass
            self.rs = torch.empty(batch, 3, NET_SIZE, NET_SIZE, **u8)
        self.L = torch.empty(batch, size, size, dtype=torch.float32, device=self.dev)
        self.launches = 0

    def run(self, rgb: torch.Tensor, out: torch.Tensor, stream: int = 0):
        lib, B, S, chk = self.lib, self.B, self.S, _lib.check
        assert rgb.shape == (B, 3, S, S) and rgb.dtype == torch.uint8 and rgb.is_contiguous()
        src = rgb
        if self.resize:                                            # Pillow BICUBIC: horizontal pass, then vertical
            t = self.tab_h
            chk(lib.havc_pil_resample_u8(rgb.data_ptr(), self.tmp.data_ptr(), B * 3, S, S, NET_SIZE, 1, t.bounds.data_ptr(),
                                         t.coeffs.data_ptr(), t.ksize, stream), "zhang.resize_h")
            chk(lib.havc_pil_resample_u8(self.tmp.data_ptr(), self.rs.data_ptr(), B * 3, S, NET_SIZE, NET_SIZE, 0, t.bounds.data_ptr(),
                                         t.coeffs.data_ptr(), t.ksize, stream), "zhang.resize_v")
            src = self.rs
            chk(lib.havc_zhang_pre(rgb.data_ptr(), B, S * S, self.L.data_ptr(), None, self.hd, stream), "zhang.L_orig")
            chk(lib.havc_zhang_pre(src.data_ptr(), B, NET_SIZE * NET_SIZE, None, self.prog.x.data_ptr(), self.hd, stream), "zhang.L_rs")
        else:
            chk(lib.havc_zhang_pre(rgb.data_ptr(), B, S * S, self.L.data_ptr(), self.prog.x.data_ptr(), self.hd, stream), "zhang.L")
        self.prog.run(stream)
        chk(lib.havc_zhang_post(self.prog.ab.data_ptr(), NET_SIZE, NET_SIZE, self.L.data_ptr(), out.data_ptr(), B, S, S, stream),
            "zhang.post")


class _PilTable:
    def __init__(self, src: int, dst: int, filt: str, dev):
        bounds, coeffs = resample.pil_tables(src, dst, filt)
        self.ksize = int(coeffs.shape[1])
        self.bounds = torch.from_numpy(bounds).to(dev)
        self.coeffs = torch.from_numpy(coeffs).to(dev)
