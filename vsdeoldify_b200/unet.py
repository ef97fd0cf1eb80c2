"""DeOldify DynamicUnet (wide / deep) as a program of libhavc_b200 launches.

Host side of the network: reads a state-dict in the reference's schema (Learner.load,
vsdeoldify/fastai/basic_train.py:264-286), folds the re-parametrisations and the eval BatchNorms
(SURVEY.md Appendix C), packs 16-bit K-major weights, allocates the NHWC activation buffers for a fixed
(batch, S) and emits the ordered launch list.  The list is replayed through a CUDA graph by
`vsdeoldify_b200.engine`.  No arithmetic on activations happens in Python/torch.

Graph (reference: DynamicUnetWide unet.py:208-285, DynamicUnetDeep unet.py:94-166):
  x -> im2col+GEMM stem(+BN+ReLU) -> maxpool -> resnet blocks (BN folded, ReLU/residual in epilogues)
    -> BN+ReLU -> middle_conv x2 (ReLU->BN epilogue)
    -> 4 x [shuf 1x1 GEMM (BN folded, ReLU, PixelShuffle store) -> blur ; BN+ReLU(skip) ; 3x3 GEMM over two
            K sources (no concat) (ReLU->BN epilogue) ; (+ self-attention as 4 GEMMs + row soft-max)]
    -> shuf 1x1 (+bias, ReLU, PixelShuffle store) -> blur = u
    -> res_block on cat([u, x]): 2 x 3x3 GEMM whose K loop is 9 taps x u-channels + ONE im2col chunk holding the 3
       image channels of all 9 taps (37 instead of 45 K steps); the first stores its 3 extra output channels to a
       side tensor (column split), the second adds the identity from (u, x) and applies the 1x1 head in its
       epilogue, writing only fp32 logits -> head kernel (sigmoid, de-normalise, quantise, luma transplant).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch

import math
import os

from . import _lib, ops
from .ops import Pair, chan_storage, hi_of, lo_of, pad_to

# Operand precision policy (DESIGN.md "Precision"; tools/precision_emulator.py is the error budget it follows).  16-bit operand
# storage injects its error where nothing damps it: ~95 % of the logit error of the U-Nets comes from the ResNet encoder (7 % of
# the FLOPs of the wide net, 2.4 % of the deep one), > 90 % of eccv16's from its first four blocks.  Those parts can run
# split-precision (value = hi + lo, products hi*hi + lo*hi + hi*lo: three MMAs per K step, ~2^-20 operands).
#   fast     - 16-bit operands everywhere (one MMA per K step)
#   balanced - split-precision encoders for both U-Nets and eccv16's first four blocks; everything else as in `fast`
#   auto     - [default] split precision only where `fast` misses the north-star gate (mean dE00 <= 0.5) on the synthetic-weight
#              parity suite: the artistic (ResNet-34) generator and eccv16; the video / stable generator passes it 2x over as is
PRECISION = os.environ.get("HAVC_B200_PRECISION", "auto")
PRECISIONS = ("fast", "balanced", "auto")
FUSE_BLUR = os.environ.get("HAVC_B200_FUSE_BLUR", "1") != "0"          # A/B switch for profiling

SD = Dict[str, torch.Tensor]
BN_EPS = 1e-5


# ------------------------------------------------------------------------------------------------
# folding (load time, fp32/fp64 on the CPU)
# ------------------------------------------------------------------------------------------------
def bn_affine(sd: SD, p: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """eval BatchNorm as y = x*scale + shift."""
    scale = sd[p + ".weight"].double() / torch.sqrt(sd[p + ".running_var"].double() + BN_EPS)
    shift = sd[p + ".bias"].double() - sd[p + ".running_mean"].double() * scale
    return scale.float(), shift.float()


def folded_weight(sd: SD, p: str) -> torch.Tensor:
    """Effective conv weight in eval mode:
      spectral norm   W = weight_orig / sigma, sigma = u^T W_mat v with the STORED u, v (no power iteration in eval);
      legacy spectral-norm state-dicts (version < 1: keys `weight`, `weight_orig`, `weight_u`, no `weight_v`): torch's load
          hook (torch/nn/utils/spectral_norm.py `_load_from_state_dict` + `_solve_v_and_rescale`) sets
          sigma = mean(weight_orig / weight) and solves v so that u^T W_mat v == sigma exactly, hence W = weight_orig / sigma;
      weight norm     W = g * v / ||v|| (norm over all dims but 0);
      plain           W = weight."""
    if p + ".weight_orig" in sd:
        w = sd[p + ".weight_orig"].double()
        if p + ".weight_v" in sd:
            u, v = sd[p + ".weight_u"].double(), sd[p + ".weight_v"].double()
            sigma = torch.dot(u, torch.mv(w.flatten(1), v))
        elif p + ".weight" in sd:
            sigma = (w / sd[p + ".weight"].double()).mean()
        else:
            raise KeyError(f"{p}: spectral-norm entry without weight_v (current schema) or weight (legacy schema)")
        return (w / sigma).float()
    if p + ".weight_g" in sd:
        v, g = sd[p + ".weight_v"].double(), sd[p + ".weight_g"].double()
        n = v.flatten(1).norm(dim=1).view(-1, *[1] * (v.dim() - 1))
        return (v * (g / n)).float()
    return sd[p + ".weight"].float()


# ------------------------------------------------------------------------------------------------
@dataclass
class Op:
    name: str
    fn: Callable[[int], None]
    flops: float = 0.0       # algorithmic FLOPs (2*MAC, unpadded)
    bytes: float = 0.0       # algorithmic HBM bytes for memory-bound ops
    kind: str = "aux"


class LaunchProgram:
    """Ordered launch list over libhavc_b200 for one (state-dict, batch, size, dtype): buffers, folded parameters and
    the op emitters shared by the DeOldify U-Nets (below) and the Zhang colorizers (zhang.py)."""

    def __init__(self, sd: SD, batch: int, size: int, dtype: torch.dtype = torch.float16, device="cuda",
                 keep_taps: bool = False, precision: Optional[str] = None):
        self.sd, self.B, self.S, self.dtype, self.dev = sd, batch, size, dtype, torch.device(device)
        self.precision = precision or PRECISION
        if self.precision not in PRECISIONS:
            raise ValueError(f"unknown precision policy {self.precision!r} ({' | '.join(PRECISIONS)})")
        self.enc_x3 = self.precision == "balanced"          # UnetProgram refines this for 'auto'
        self.hd = ops.havc_dtype(dtype)
        self.lib = _lib.lib()
        self.ops: List[Op] = []
        self.keep: list = []
        self.taps: Dict[str, torch.Tensor] = {}
        self.keep_taps = keep_taps
        self.head_flops = 0.0

    # ---- buffers / params ---------------------------------------------------------------------
    def buf(self, *shape, dtype=None, zero=False) -> torch.Tensor:
        t = (torch.zeros if zero else torch.empty)(*shape, dtype=dtype or self.dtype, device=self.dev)
        self.keep.append(t)
        return t

    def dev_f32(self, v: torch.Tensor) -> torch.Tensor:
        t = v.detach().float().contiguous().to(self.dev)
        self.keep.append(t)
        return t

    def tap(self, name, t):
        if self.keep_taps:
            self.taps[name] = t
        return t

    # ---- op emitters --------------------------------------------------------------------------
    def conv(self, name: str, src0, w: torch.Tensor, *, ks=1, src1=None, cin_splits=None, stride=1, bias=None,
             relu1=False, scale=None, shift=None, residual=None, relu2=False, shuffle=False, out=None,
             out_c: Optional[int] = None, dilation: int = 1, leaky1: float = 0.0, out_dtype=None, taps=None,
             phase=None, flops: Optional[float] = None, head_w=None, head_out=None, x3: bool = False, blur: bool = False):
        """w: folded fp32 [Cout, Cin, ks, ks].  Returns the NHWC output tensor.
        taps: explicit (dh, dw, phase, weight tap) list (transposed convolutions); phase = (up, oy, ox): the output
        pixel of (h, w) is (h*up+oy, w*up+ox) of `out`.
        x3: split-precision launch - the weight is packed as hi + lo planes, sources / residual that are `Pair`s contribute
        their lo planes, and the result is a `Pair`.  The weight is pre-scaled by a power of two (undone by the epilogue's
        scale) so that its lo plane stays in fp16's normal range.  A launch without x3 reads only the hi plane of a Pair.
        blur (with shuffle): the ICNR blur runs in the epilogue of the PixelShuffle launch (havc_conv_desc.blur)."""
        Cout, Cin = w.shape[0], w.shape[1]
        h0, h1 = hi_of(src0), hi_of(src1) if src1 is not None else None
        storage = [h0.shape[-1]] + ([h1.shape[-1]] if h1 is not None else [])
        wl = None
        if x3:
            assert not shuffle and head_w is None and out is None
            k = max(-8, min(14, int(math.floor(math.log2(16384.0 / max(float(w.abs().max()), 1e-30))))))
            g = 2.0 ** k
            w32, meta = ops.pack_conv_weight(w.double() * g, cin_splits, dtype=None, shuffle=False, cin_storage=storage)
            wp, wl = ops.split_hi_lo(w32, self.dtype)
            wl = wl.to(self.dev)
            self.keep.append(wl)
            bias = None if bias is None else bias * g                      # (acc*g + bias*g) -> act -> * (scale/g) + shift
            scale = (torch.ones(Cout) if scale is None else scale.float().cpu()) / g
            shift = torch.zeros(Cout) if shift is None else shift
        else:
            wp, meta = ops.pack_conv_weight(w, cin_splits, dtype=self.dtype, shuffle=("blur" if (shuffle and blur) else shuffle),
                                            cin_storage=storage)
        wp = wp.to(self.dev)
        self.keep.append(wp)
        n_total = meta["rows"]
        pc = lambda v, fill: None if v is None else self.dev_f32(ops.pack_cols(v, n_total, fill, meta if shuffle else None))
        if stride == 1:
            B, H, W = h0.shape[0], h0.shape[1], h0.shape[2]
            if taps is None:
                taps = ops.taps_for(ks, dilation)
        else:  # src0 is phase-split [P,B,H/2,W/2,C]
            B, H, W = h0.shape[1], h0.shape[2], h0.shape[3]
            taps = ops.taps_stride2(ks) if ks > 1 else [(0, 0, 0, 0)]
        c_real = meta["cg"] if shuffle else (out_c or Cout)
        up, oy, ox = phase if phase is not None else (1, 0, 0)
        if out is None and head_w is None:     # zero-initialised: the pad channels beyond c_store are never written and stay zero
            if x3:
                out = Pair(self.buf(2, B, up * H, up * W, chan_storage(c_real), zero=True))
            else:
                out = self.buf(B, 2 * H, 2 * W, chan_storage(c_real), zero=True) if shuffle else \
                    self.buf(B, up * H, up * W, chan_storage(c_real), zero=True, dtype=out_dtype)
        lo = (lambda t: lo_of(t)) if x3 else (lambda t: None)
        op = ops.make_conv(h0, wp, hi_of(out), taps, src1=h1, w_c1_off=meta["c1_off"], n_total=n_total,
                           bias=pc(bias, 0.0), scale=pc(scale, 1.0), shift=pc(shift, 0.0), relu1=relu1, relu2=relu2,
                           residual=hi_of(residual), out_space=(B, H, W), shuffle=shuffle, group_n=meta.get("group_n", 0),
                           c_store=pad_to(c_real, 8), up=up, oy=oy, ox=ox, leaky1=leaky1, head_w=head_w, head_out=head_out,
                           bn=(4 * ops.BLUR_CW) if blur else (n_total if head_w is not None else None), name=name,
                           box=(16, 8, 1) if blur else None, blur=blur, src0_lo=lo(src0), src1_lo=lo(src1),
                           weight_lo=wl, out_lo=lo(out), residual_lo=lo(residual))
        if flops is None:
            flops = 2.0 * B * H * W * Cout * Cin * len(taps)
        self.ops.append(Op(name, op.launch, flops=flops, kind="gemm"))
        self.keep.append(op)
        return out

    def aux(self, name, fn, nbytes=0.0):
        self.ops.append(Op(name, fn, bytes=nbytes))

    def affine(self, name, src, scale, shift, relu, out=None, out_c_off=0, c=None):
        """src may be a Pair (read as hi + lo); the result is a plain 16-bit tensor."""
        sh_, sl_ = hi_of(src), lo_of(src)
        B, H, W, Cs = sh_.shape
        c = c or Cs
        if out is None:
            out = self.buf(B, H, W, Cs, zero=True)
        sc = self.dev_f32(torch.cat([scale, scale.new_ones(pad_to(c, 8) - scale.numel())]))
        sh = self.dev_f32(torch.cat([shift, shift.new_zeros(pad_to(c, 8) - shift.numel())]))
        in_ptr, out_ptr = sh_.data_ptr(), out.data_ptr() + 2 * out_c_off
        lo_ptr = sl_.data_ptr() if sl_ is not None else None
        n_pix, cc, istr, ostr, hd, lib = B * H * W, pad_to(c, 8), sh_.stride(2), out.stride(2), self.hd, self.lib

        def fn(stream):
            _lib.check(lib.havc_affine_act(in_ptr, lo_ptr, out_ptr, None, n_pix, cc, istr, ostr, sc.data_ptr(), sh.data_ptr(),
                                           int(relu), hd, stream), name)
        self.aux(name, fn, nbytes=(4.0 if sl_ is None else 6.0) * n_pix * cc)
        return out

    def blur(self, name, src, out=None):
        B, H, W, Cs = src.shape
        if out is None:
            out = self.buf(B, H, W, Cs, zero=True)
        ip, op_, ostr, hd, lib = src.data_ptr(), out.data_ptr(), out.stride(2), self.hd, self.lib

        def fn(stream):
            _lib.check(lib.havc_blur2x2(ip, op_, B, H, W, Cs, ostr, hd, stream), name)
        self.aux(name, fn, nbytes=4.0 * B * H * W * Cs)
        return out

    def phase_split(self, name, src, n_phases):
        """src: [B,H,W,C] tensor or Pair -> [P,B,H/2,W/2,C] (Pair: both planes, one launch each)."""
        B, H, W, Cs = src.shape
        pair = isinstance(src, Pair)
        shape = (n_phases, B, (H + 1) // 2, (W + 1) // 2, Cs)
        out = Pair(self.buf(2, *shape, zero=True)) if pair else self.buf(*shape, zero=True)
        lib = self.lib
        ptrs = [(src.hi.data_ptr(), out.hi.data_ptr()), (src.lo.data_ptr(), out.lo.data_ptr())] if pair else \
            [(src.data_ptr(), out.data_ptr())]

        def fn(stream):
            for ip, op_ in ptrs:
                _lib.check(lib.havc_phase_split(ip, op_, B, H, W, Cs, n_phases, stream), name)
        self.aux(name, fn, nbytes=2.0 * len(ptrs) * B * H * W * Cs * (1 + n_phases / 4))
        return out

    # ---- execution ----------------------------------------------------------------------------
    def run(self, stream: int = 0):
        for op in self.ops:
            op.fn(stream)

    @property
    def flops(self) -> float:
        return sum(o.flops for o in self.ops) + self.head_flops


class UnetProgram(LaunchProgram):
    """Compiled launch list of a DeOldify generator for one (state-dict, batch, S, dtype)."""

    def __init__(self, sd: SD, batch: int, size: int, dtype: torch.dtype = torch.float16, device="cuda",
                 keep_taps: bool = False, x: Optional[torch.Tensor] = None, precision: Optional[str] = None):
        if size % 16 != 0 or size < 32:
            raise ValueError(f"render size {size} must be render_factor * 16 (a multiple of 16, at least 32)")
        super().__init__(sd, batch, size, dtype, device, keep_taps, precision)
        self.bottleneck = "layers.0.4.0.conv3.weight" in sd
        if self.precision == "auto":
            self.enc_x3 = not self.bottleneck               # the artistic generator (ResNet-34 basic blocks)
        self._x_shared = x
        self._build()

    # ---- network ------------------------------------------------------------------------------
    def _build(self):
        sd, B, S, lib, hd = self.sd, self.B, self.S, self.lib, self.hd
        # network input: normalised image, NHWC [B,S,S,8] (3 real channels), written by the pre kernel
        self.x = self._x_shared if self._x_shared is not None else self.buf(B, S, S, 8, zero=True)

        # ---- encoder stem: 7x7/s2 conv as im2col + GEMM, BN folded, ReLU --------------------------
        # The three input channels are the same gray L (ColorizerFilter._transform, filters.py:92-93), normalised per channel:
        # x_c = (L/255 - mean_c)/std_c.  The normalisation is folded into the weights and the GEMM reads the EXACT integer
        # channels (L, 1) the pre kernel stores in x[..., 4:6] (the '1' channel carries the -mean/std term and, being zero
        # outside the image like every im2col tap, reproduces the zero padding of the normalised image):
        #   sum_c w_c x_c = (sum_c w_c / (255 std_c)) * L  +  (-sum_c w_c mean_c / std_c) * 1
        x3 = self.enc_x3
        w = sd["layers.0.0.weight"].double()
        sc, sh = bn_affine(sd, "layers.0.1")
        mean = torch.tensor([0.485, 0.456, 0.406], dtype=torch.float64).view(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225], dtype=torch.float64).view(1, 3, 1, 1)
        ws = w * sc.double().view(-1, 1, 1, 1)
        w2 = torch.stack([(ws / (255.0 * std)).sum(1), -(ws * mean / std).sum(1)], 1)          # [64, 2, 7, 7]
        # K order of havc_im2col_small: k = kh*16 + kw*2 + c  (each filter row padded from 14 to 16)
        Kp = chan_storage(7 * 16)
        wf = torch.cat([w2.permute(0, 2, 3, 1).reshape(64, 7, 14), w2.new_zeros(64, 7, 2)], 2).reshape(64, 112)
        wf = torch.cat([wf, wf.new_zeros(64, Kp - 112)], 1).view(64, Kp, 1, 1).float()
        H2 = S // 2
        col = self.buf(B, H2, H2, Kp, zero=True)
        xp, cp = self.x.data_ptr(), col.data_ptr()

        def im2col(stream):
            _lib.check(lib.havc_im2col_small(xp, cp, B, S, S, 8, 4, 2, 7, 2, 3, Kp, hd, stream), "stem.im2col")
        self.aux("stem.im2col", im2col, nbytes=2.0 * B * H2 * H2 * Kp)
        stem = self.conv("stem.conv", col, wf, bias=sh, relu1=True, x3=x3)
        self.ops[-1].flops = 2.0 * B * H2 * H2 * 64 * 147
        self.tap("enc.stem", stem)
        H4 = S // 4
        pool = Pair(self.buf(2, B, H4, H4, 64)) if x3 else self.buf(B, H4, H4, 64)
        sp, pp = hi_of(stem).data_ptr(), hi_of(pool).data_ptr()
        spl = stem.lo.data_ptr() if x3 else None
        ppl = pool.lo.data_ptr() if x3 else None

        def maxpool(stream):
            _lib.check(lib.havc_maxpool3x3s2(sp, spl, pp, ppl, B, H2, H2, 64, hd, stream), "stem.maxpool")
        self.aux("stem.maxpool", maxpool, nbytes=(4.0 if x3 else 2.0) * B * (H2 * H2 + H4 * H4) * 64)

        # ---- resnet layers --------------------------------------------------------------------------
        y = pool
        skips = []
        for li in (4, 5, 6, 7):
            bi = 0
            while f"layers.0.{li}.{bi}.conv1.weight" in sd:
                stride = 2 if (bi == 0 and li > 4) else 1
                y = (self._bottleneck if self.bottleneck else self._basic)(y, f"layers.0.{li}.{bi}", stride)
                bi += 1
            self.tap(f"enc.layer{li - 3}", y)
            skips.append(y)
        skips = [skips[2], skips[1], skips[0], stem]

        # ---- layers[1..2] BN + ReLU; layers[3] middle_conv ----------------------------------------
        sc, sh = bn_affine(sd, "layers.1")
        y = self.affine("enc.bn_relu", y, sc, sh, True)
        self.tap("enc.out", y)
        for j in (0, 1):
            sc, sh = bn_affine(sd, f"layers.3.{j}.2")
            y = self.conv(f"middle.{j}", y, folded_weight(sd, f"layers.3.{j}.0"), ks=3, relu1=True, scale=sc, shift=sh)
        self.tap("middle", y)

        # ---- U-Net blocks ---------------------------------------------------------------------------
        for i, s in enumerate(skips):
            y = self._unet_block(y, s, f"layers.{4 + i}")
            self.tap(f"block{i}", y)

        # ---- layers[8] PixelShuffle_ICNR; layers[9] MergeLayer(dense); layers[10] res_block; layers[11] head -------
        w8 = folded_weight(sd, "layers.8.conv.0")
        if self._fuse_blur(y):                                  # [B,S,S,cu_s]: the 'x' half of cat([x, x.orig])
            u = self.conv("shuf8.conv+blur", y, w8, bias=sd["layers.8.conv.0.bias"].float(), relu1=True, shuffle=True, blur=True)
        else:
            t8 = self.conv("shuf8.conv", y, w8, bias=sd["layers.8.conv.0.bias"].float(), relu1=True, shuffle=True)
            u = self.blur("shuf8.blur", t8)
        self.tap("shuf8", u)
        cu = w8.shape[0] // 4                                   # real channels of u (256 wide / 300 deep)
        cu_s = u.shape[-1]                                      # storage width (multiple of 64)
        split = pad_to(cu, 16)                                  # GEMM columns [0,split) = main, [split,split+16) = 3 image ch
        n_tot = split + 16
        ncat = cu + 3

        def im2col3(name, src):
            """3x3 neighbourhood of the 3 leading channels of an 8-channel tensor -> one 64-wide K chunk."""
            col = self.buf(B, S, S, 64, zero=True)
            ip, op_ = src.data_ptr(), col.data_ptr()

            def fn(stream):
                _lib.check(lib.havc_im2col_small(ip, op_, B, S, S, 8, 0, 3, 3, 1, 1, 64, hd, stream), name)
            self.aux(name, fn, nbytes=B * S * S * (16.0 + 96.0))
            return col

        def pack_tail(wt):
            """[ncat, ncat, 3, 3] -> [n_tot rows][9 taps][cu_s + 64]: main channels per tap, the 3 image channels of all
            9 taps as one im2col chunk (k = kh*16 + kw*3 + c) stored at weight tap 0, channels [cu_s, cu_s+64)."""
            wk = torch.zeros(n_tot, 9, cu_s + 64)
            rows = torch.cat([torch.arange(cu), split + torch.arange(3)])
            wk[rows, :, :cu] = wt[:, :cu].reshape(ncat, cu, 9).permute(0, 2, 1)
            for kh in range(3):
                for kw in range(3):
                    wk[rows, 0, cu_s + kh * 16 + kw * 3: cu_s + kh * 16 + kw * 3 + 3] = wt[:, cu:, kh, kw]
            return wk.to(self.dtype).contiguous().to(self.dev)

        def cols(v, fill=0.0):
            o = torch.full((n_tot,), fill)
            o[:cu] = v[:cu]
            o[split:split + 3] = v[cu:]
            return self.dev_f32(o)

        taps9 = ops.taps_for(3)
        xcol = im2col3("cat.im2col_x", self.x)
        w0 = pack_tail(folded_weight(sd, "layers.10.layers.0.0"))
        w1 = pack_tail(folded_weight(sd, "layers.10.layers.1.0"))
        self.keep += [w0, w1]
        r1 = self.buf(B, S, S, cu_s, zero=True)
        r1x = self.buf(B, S, S, 8, zero=True)
        op = ops.make_conv(u, w0, r1, taps9, src1=xcol, w_c1_off=cu_s, n_total=n_tot, bn=n_tot,
                           bias=cols(sd["layers.10.layers.0.0.bias"].float()), relu1=True, out_space=(B, S, S),
                           c_store=pad_to(cu, 8), src1_single_tap=True, src1_wi=0, split_n=split, out2=r1x, c_store2=8,
                           name="res.conv0")
        self.ops.append(Op("res.conv0", op.launch, flops=2.0 * B * S * S * ncat * ncat * 9, kind="gemm"))
        self.keep.append(op)
        r1col = im2col3("res.im2col_r1", r1x)
        # second conv: +bias, ReLU, + cat([u, x]) (the res_block's identity), then the fused 1x1 head -> fp32 logits
        w11 = folded_weight(sd, "layers.11.0").view(3, ncat)
        w11p = torch.zeros(3, n_tot)
        w11p[:, :cu] = w11[:, :cu]
        w11p[:, split:split + 3] = w11[:, cu:]
        self.w11 = self.dev_f32(w11p)
        self.b11 = self.dev_f32(sd["layers.11.0.bias"].float())
        self.logits = self.buf(B, S, S, 4, dtype=torch.float32, zero=True)
        op = ops.make_conv(r1, w1, None, taps9, src1=r1col, w_c1_off=cu_s, n_total=n_tot, bn=n_tot,
                           bias=cols(sd["layers.10.layers.1.0.bias"].float()), relu1=True, out_space=(B, S, S),
                           c_store=pad_to(cu, 8), src1_single_tap=True, src1_wi=0, split_n=split, c_store2=8,
                           residual=u, residual2=self.x, head_w=self.w11, head_out=self.logits, name="res.conv1")
        self.ops.append(Op("res.conv1+head", op.launch, flops=2.0 * B * S * S * ncat * ncat * 9 + 2.0 * B * S * S * ncat * 3,
                           kind="gemm"))
        self.keep.append(op)
        self.tap("logits", self.logits)
        self.head_flops = 0.0

    def _fuse_blur(self, src) -> bool:
        """Fuse the ICNR blur into the PixelShuffle launch (halo recompute: 15 x 7 useful pixels per 16 x 8 tile) where the
        tensor is big enough for the saved HBM round trip (write + read + write of the shuffled tensor) to outweigh the 22 % of
        recomputed 1x1-conv rows: the 48 x 48 inputs and up (96 % of the blur bytes of the wide net)."""
        return FUSE_BLUR and min(src.shape[1], src.shape[2]) >= 48

    def _fold_bn(self, p_conv, p_bn):
        w = self.sd[p_conv + ".weight"].float()
        sc, sh = bn_affine(self.sd, p_bn)
        return w * sc.view(-1, 1, 1, 1), sh

    def _bottleneck(self, x, p, stride):
        """torchvision Bottleneck, v1.5 (stride on conv2); BN folded into each conv."""
        x3 = self.enc_x3
        w1, b1 = self._fold_bn(p + ".conv1", p + ".bn1")
        w2, b2 = self._fold_bn(p + ".conv2", p + ".bn2")
        w3, b3 = self._fold_bn(p + ".conv3", p + ".bn3")
        c1 = self.conv(p + ".conv1", x, w1, bias=b1, relu1=True, x3=x3)
        if stride == 2:
            ph = self.phase_split(p + ".split", c1, 4)
            c2 = self.conv(p + ".conv2", ph, w2, ks=3, stride=2, bias=b2, relu1=True, x3=x3)
        else:
            c2 = self.conv(p + ".conv2", c1, w2, ks=3, bias=b2, relu1=True, x3=x3)
        idt = x
        if p + ".downsample.0.weight" in self.sd:
            wd, bd = self._fold_bn(p + ".downsample.0", p + ".downsample.1")
            if stride == 2:
                xs = self.phase_split(p + ".ds_split", x, 1)
                idt = self.conv(p + ".downsample", xs, wd, stride=2, bias=bd, x3=x3)
            else:
                idt = self.conv(p + ".downsample", x, wd, bias=bd, x3=x3)
        return self.conv(p + ".conv3", c2, w3, bias=b3, residual=idt, relu2=True, x3=x3)

    def _basic(self, x, p, stride):
        x3 = self.enc_x3
        w1, b1 = self._fold_bn(p + ".conv1", p + ".bn1")
        w2, b2 = self._fold_bn(p + ".conv2", p + ".bn2")
        idt = x
        if stride == 2:
            ph = self.phase_split(p + ".split", x, 4)
            c1 = self.conv(p + ".conv1", ph, w1, ks=3, stride=2, bias=b1, relu1=True, x3=x3)
            if p + ".downsample.0.weight" in self.sd:
                wd, bd = self._fold_bn(p + ".downsample.0", p + ".downsample.1")
                ph0 = ph.sub(slice(0, 1)) if isinstance(ph, Pair) else ph[0:1]
                idt = self.conv(p + ".downsample", ph0, wd, stride=2, bias=bd, x3=x3)
        else:
            c1 = self.conv(p + ".conv1", x, w1, ks=3, bias=b1, relu1=True, x3=x3)
            if p + ".downsample.0.weight" in self.sd:
                wd, bd = self._fold_bn(p + ".downsample.0", p + ".downsample.1")
                idt = self.conv(p + ".downsample", x, wd, bias=bd, x3=x3)
        return self.conv(p + ".conv2", c1, w2, ks=3, bias=b2, residual=idt, relu2=True, x3=x3)

    def _unet_block(self, up_in, skip, p):
        """UnetBlockWide / UnetBlockDeep (unet.py:170-205 / 55-91)."""
        sd = self.sd
        # shuf: conv1x1 (no bias) -> BN -> ReLU -> PixelShuffle -> blur; BN folds into the conv (exact: 1x1, no pad)
        ws = folded_weight(sd, p + ".shuf.conv.0")
        sc, sh = bn_affine(sd, p + ".shuf.conv.1")
        if self._fuse_blur(up_in):
            u = self.conv(p + ".shuf.conv+blur", up_in, ws * sc.view(-1, 1, 1, 1), bias=sh, relu1=True, shuffle=True, blur=True)
        else:
            t = self.conv(p + ".shuf.conv", up_in, ws * sc.view(-1, 1, 1, 1), bias=sh, relu1=True, shuffle=True)
            u = self.blur(p + ".shuf.blur", t)
        if tuple(u.shape[1:3]) != tuple(skip.shape[1:3]):
            # odd render factors: the encoder halves 16*odd/16 = odd with a ceil, so the up path is one pixel larger than the skip
            # and the reference resizes it with F.interpolate(mode='nearest') (unet.py:201-203 / 87-89).  For in = out + 1 the
            # nearest source index floor(d * (out + 1) / out) is d itself: the resize is a crop of the last row / column.
            Hs, Ws = int(skip.shape[1]), int(skip.shape[2])
            if (int(u.shape[1]), int(u.shape[2])) != (Hs + 1, Ws + 1):
                raise ValueError(f"up-path {tuple(u.shape[1:3])} vs skip {(Hs, Ws)}: only the one-pixel mismatch of odd render factors is handled")
            u = u[:, :Hs, :Ws, :]
        sc, sh = bn_affine(sd, p + ".bn")
        sb = self.affine(p + ".skip_bn_relu", skip, sc, sh, True)
        cu = ws.shape[0] // 4
        cs = sc.numel()
        convs = ["conv"] if p + ".conv.0.weight_orig" in sd else ["conv1", "conv2"]
        y = None
        for k, cv in enumerate(convs):
            w = folded_weight(sd, f"{p}.{cv}.0")
            bsc, bsh = bn_affine(sd, f"{p}.{cv}.2")
            if k == 0:
                y = self.conv(f"{p}.{cv}", u, w, ks=3, src1=sb, cin_splits=[cu, cs], relu1=True, scale=bsc, shift=bsh)
            else:
                y = self.conv(f"{p}.{cv}", y, w, ks=3, relu1=True, scale=bsc, shift=bsh)
            if f"{p}.{cv}.3.gamma" in sd:
                y = self._attention(y, f"{p}.{cv}.3")
        return y

    def _attention(self, x, p):
        """SelfAttention (fastai/layers.py:81-96) as GEMMs + a row soft-max; N x N logits in fp32."""
        sd, B = self.sd, self.B
        _, H, W, Cc = x.shape
        N = H * W
        wq, wk, wv = (folded_weight(sd, f"{p}.{n}") for n in ("query", "key", "value"))
        d = wq.shape[0]
        gamma = float(sd[p + ".gamma"].flatten()[0])
        xt = x.view(B, 1, N, Cc)
        q = self.conv(p + ".query", xt, wq)          # f  [B,1,N,d]
        k = self.conv(p + ".key", xt, wk)            # g  [B,1,N,d]
        dp = q.shape[-1]
        # Ht[b, c, i] = sum_k Wv[c,k] x[b,i,k]  (A = Wv shared, B = tokens per image)
        wv16 = self.buf(1, 1, Cc, Cc, zero=True)              # [rows = out channel (padded), K = in channel (padded)]
        wv16[0, 0, :wv.shape[0], :wv.shape[1]].copy_(wv.view(wv.shape[0], wv.shape[1]).to(self.dtype))
        Ns = chan_storage(N)                                   # K storage of the P.V GEMM
        Cr = wv.shape[0]
        Ht = self.buf(B, 1, Cc, Ns, zero=True)
        op = ops.make_conv(wv16, x.view(B, N, 1, Cc), Ht, [(0, 0, 0, 0)], n_total=pad_to(N, 16), a_batched=False,
                           b_batched=True, out_space=(B, 1, Cc), c_store=pad_to(N, 8), name=p + ".value_t")
        self.ops.append(Op(p + ".value_t", op.launch, flops=2.0 * B * N * Cc * Cc, kind="gemm"))
        self.keep.append(op)
        # S[b, j, i] = sum_c g[b,j,c] f[b,i,c]
        # fp16 path: the N x N logits make their HBM round trip as fp16 through the TMA-store epilogue (|S| ~ 40 at most; the 2^-11
        # rounding disappears in the soft-max: tools/precision_emulator.py); bf16's 8 mantissa bits would not do, it keeps fp32
        s16 = self.dtype == torch.float16 and N % 64 == 0
        # rows padded to Ns either way: with N % 8 != 0 (odd render factors) the epilogue stores up to 7 pad columns per row
        Sx = self.buf(B, 1, N, Ns, zero=True) if s16 else self.buf(B, 1, N, Ns, dtype=torch.float32, zero=True)
        op = ops.make_conv(k, q.view(B, N, 1, dp), Sx, [(0, 0, 0, 0)], n_total=pad_to(N, 16), b_batched=True,
                           out_space=(B, 1, N), c_store=pad_to(N, 8), name=p + ".logits")
        self.ops.append(Op(p + ".logits", op.launch, flops=2.0 * B * N * N * d, kind="gemm"))
        self.keep.append(op)
        P = self.buf(B, 1, N, Ns, zero=True)
        sp, pp, hd, lib = Sx.data_ptr(), P.data_ptr(), self.hd, self.lib
        s_dt, s_stride = (_lib.HAVC_F16 if s16 else _lib.HAVC_F32), Ns

        def softmax(stream):
            _lib.check(lib.havc_softmax_rows(sp, s_dt, pp, B * N, N, s_stride, Ns, hd, stream), p + ".softmax")
        self.aux(p + ".softmax", softmax, nbytes=(4.0 if s16 else 6.0) * B * N * N)
        # out[b, j, c] = x[b,j,c] + gamma * sum_i P[b,j,i] Ht[b,c,i]
        out = self.buf(B, H, W, Cc, zero=True)
        gs = self.dev_f32(torch.full((pad_to(Cc, 16),), gamma))
        gz = self.dev_f32(torch.zeros(pad_to(Cc, 16)))
        op = ops.make_conv(P, Ht.view(B, Cc, 1, Ns), out.view(B, 1, N, Cc), [(0, 0, 0, 0)], n_total=pad_to(Cc, 16),
                           b_batched=True, out_space=(B, 1, N), scale=gs, shift=gz, residual=xt,
                           c_store=pad_to(Cr, 8), name=p + ".pv")
        self.ops.append(Op(p + ".pv", op.launch, flops=2.0 * B * N * N * Cc, kind="gemm"))
        self.keep.append(op)
        return out

