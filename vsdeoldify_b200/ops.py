"""Host-side launch helpers over the C ABI: build `havc_conv_desc` records from torch tensor handles.

torch is used here only for device memory (tensor handles) and dtype bookkeeping; all arithmetic is
done by libhavc_b200.so.  Layout conventions:

  * activations: NHWC `[B, H, W, C]` (or phase-split `[P, B, H, W, C]`), C and strides multiples of 8;
  * packed conv weights: `[rows, taps, cin]` with `rows` = GEMM N (padded to 16), `cin` = concatenated
    K sources each padded to 8; pad rows / channels are zero.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import HAVC_BF16, HAVC_F16, HAVC_F32, ActView, ConvDesc


TMA_STORE_DEFAULT = os.environ.get("HAVC_B200_TMA_STORE", "1") != "0"


def pad_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def havc_dtype(t: torch.dtype) -> int:
    if t == torch.float16:
        return HAVC_F16
    if t == torch.bfloat16:
        return HAVC_BF16
    if t == torch.float32:
        return HAVC_F32
    raise ValueError(f"unsupported dtype {t}")


def act_view(t: torch.Tensor) -> ActView:
    """View of an NHWC tensor [B,H,W,C] or [P,B,H,W,C] whose last dim is contiguous."""
    assert t.stride(-1) == 1, "channel dim must be contiguous"
    v = ActView()
    v.ptr = t.data_ptr()
    if t.dim() == 5:
        P, B, H, W, Cc = t.shape
        sp, sb, sh, sw, _ = t.stride()
    else:
        assert t.dim() == 4
        B, H, W, Cc = t.shape
        sb, sh, sw, _ = t.stride()
        P, sp = 1, 0
    v.C, v.W, v.H, v.B, v.P = Cc, W, H, B, P
    v.stride_w, v.stride_h, v.stride_b, v.stride_p = sw, sh, sb, sp
    return v


def choose_box(W: int, H: int, B: int) -> Tuple[int, int, int]:
    """Pick the (box_w, box_h, box_b) with product 128 that wastes the fewest padded pixels."""
    best = None
    for lw in range(8):
        for lh in range(8 - lw):
            bw, bh = 1 << lw, 1 << lh
            bb = 128 // (bw * bh)
            tiles = -(-W // bw) * -(-H // bh) * -(-B // bb)
            key = (tiles, bb, -bw)  # fewest tiles, then least batch folding, then widest rows
            if best is None or key < best[0]:
                best = (key, (bw, bh, bb))
    return best[1]


def taps_for(ks: int, dilation: int = 1) -> List[Tuple[int, int, int, int]]:
    """(dh, dw, phase, weight_tap_index) for a stride-1 'same' convolution."""
    r = (ks - 1) // 2
    return [((i - r) * dilation, (j - r) * dilation, 0, i * ks + j) for i in range(ks) for j in range(ks)]


def taps_stride2(ks: int) -> List[Tuple[int, int, int, int]]:
    """Taps of a stride-2 'same' (pad=(ks-1)//2) conv over a phase-split input [4,B,H/2,W/2,C].

    Input pixel (2i+kh-r, 2j+kw-r) lives in phase ((kh-r)&1, (kw-r)&1) at offset floor((kh-r)/2)."""
    r = (ks - 1) // 2
    out = []
    for i in range(ks):
        for j in range(ks):
            a, b = i - r, j - r
            out.append((a >> 1, b >> 1, (a & 1) * 2 + (b & 1), i * ks + j))
    return out


def chan_storage(c: int) -> int:
    """Channel storage width of an activation tensor: multiples of 64 above 64 channels (a TMA box whose
    64-channel extent is partly outside the tensor is served ~3x slower than an in-bounds one - measured),
    multiples of 8 (16 B) below.  Pad channels are zero and never written."""
    return pad_to(c, 64) if c > 64 else pad_to(c, 8)


class Pair:
    """A split-precision ("x3") tensor: value = hi + lo, two 16-bit planes of one allocation `t` = [2, ...] (plane 0 = the value
    rounded to 16 bit, plane 1 = the rounded remainder).  havc_conv_gemm multiplies such operands as hi*hi + lo*hi + hi*lo."""

    def __init__(self, t: torch.Tensor):
        assert t.shape[0] == 2
        self.t, self.hi, self.lo = t, t[0], t[1]

    @property
    def shape(self):
        return self.hi.shape

    @property
    def dtype(self):
        return self.hi.dtype

    def dim(self):
        return self.hi.dim()

    def float(self) -> torch.Tensor:
        return self.hi.float() + self.lo.float()

    def view(self, *shape) -> "Pair":
        return Pair(self.t.view(2, *shape))

    def sub(self, index) -> "Pair":
        """Pair of a slice of the leading (non-plane) dimension, e.g. one phase of a phase-split tensor."""
        return Pair(self.t[:, index])


def hi_of(t):
    return t.hi if isinstance(t, Pair) else t


def lo_of(t):
    return t.lo if isinstance(t, Pair) else None


def split_hi_lo(w: torch.Tensor, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 tensor -> (hi, lo) 16-bit tensors with hi + lo ~= w to ~2^-22 relative."""
    hi = w.to(dtype)
    lo = (w - hi.float()).to(dtype)
    return hi, lo


BLUR_CW = 64      # channels per N tile of a fused PixelShuffle + blur launch (BN = 4 * BLUR_CW = 256)


def pack_conv_weight(w: torch.Tensor, cin_splits: Optional[Sequence[int]] = None, dtype=torch.float16,
                     shuffle: bool = False, row_pad: int = 16,
                     cin_storage: Optional[Sequence[int]] = None) -> Tuple[torch.Tensor, dict]:
    """[Cout, Cin, kh, kw] (or [Cout, Cin]) fp32 -> packed [rows, taps, cin_storage] 16-bit (CPU tensor).

    cin_splits: channel counts of the K sources (torch.cat order); each is zero-padded to its storage width
    (cin_storage, default: next multiple of 8).
    shuffle: reorder rows for the PixelShuffle(2) store: row g*group_n + c <- out channel c*4 + g."""
    if w.dim() == 2:
        w = w[:, :, None, None]
    if w.dim() == 3:  # conv1d, kernel size 1
        w = w[:, :, :, None]
    Cout, Cin, kh, kw = w.shape
    cin_splits = list(cin_splits) if cin_splits else [Cin]
    assert sum(cin_splits) == Cin
    cin_storage = list(cin_storage) if cin_storage else [pad_to(c, 8) for c in cin_splits]
    parts, off, c1_off = [], 0, 0
    for i, c in enumerate(cin_splits):
        blk = w[:, off:off + c]
        cp = cin_storage[i]
        assert cp >= c and cp % 8 == 0
        if cp != c:
            blk = torch.cat([blk, blk.new_zeros(Cout, cp - c, kh, kw)], 1)
        parts.append(blk)
        off += c
        if i == 0:
            c1_off = cp
    wcat = torch.cat(parts, 1)  # [Cout, cin_storage, kh, kw]
    cin_storage = wcat.shape[1]
    wk = wcat.permute(0, 2, 3, 1).reshape(Cout, kh * kw, cin_storage)
    meta = {"c1_off": c1_off, "taps": kh * kw, "cin": cin_storage, "cout": Cout}
    if shuffle == "blur":        # fused PixelShuffle + blur launch: N tile t = [4 sub-pixel groups][BLUR_CW channels]
        assert Cout % 4 == 0
        cg, cw = Cout // 4, BLUR_CW
        nt = -(-cg // cw)
        rows = wk.new_zeros(nt * 4 * cw, kh * kw, cin_storage)
        for t in range(nt):
            c0, c1 = t * cw, min((t + 1) * cw, cg)
            for g in range(4):
                rows[t * 4 * cw + g * cw: t * 4 * cw + g * cw + (c1 - c0)] = wk[c0 * 4 + g: c1 * 4: 4]
        meta.update(blur_cw=cw, cg=cg, rows=nt * 4 * cw)
        wk = rows
    elif shuffle:
        assert Cout % 4 == 0
        cg = Cout // 4
        gn = pad_to(cg, row_pad)
        rows = wk.new_zeros(4 * gn, kh * kw, cin_storage)
        for g in range(4):
            rows[g * gn:g * gn + cg] = wk[g::4]
        meta.update(group_n=gn, cg=cg, rows=4 * gn)
        wk = rows
    else:
        rp = pad_to(Cout, row_pad)
        if rp != Cout:
            wk = torch.cat([wk, wk.new_zeros(rp - Cout, kh * kw, cin_storage)], 0)
        meta.update(rows=rp)
    if dtype is None:        # keep fp32: the caller splits it into hi / lo planes
        return wk.float().contiguous(), meta
    return wk.to(dtype).contiguous(), meta


def pack_cols(v: Optional[torch.Tensor], n_alloc: int, fill: float, shuffle_meta: Optional[dict] = None) -> Optional[torch.Tensor]:
    """Per-output-channel fp32 vector -> per-GEMM-column vector of length n_alloc (row order of the packed weight)."""
    if v is None:
        return None
    v = v.detach().float().cpu()
    out = torch.full((n_alloc,), fill, dtype=torch.float32)
    if shuffle_meta and "blur_cw" in shuffle_meta:
        cw, cg = shuffle_meta["blur_cw"], shuffle_meta["cg"]
        for t in range(-(-cg // cw)):
            c0, c1 = t * cw, min((t + 1) * cw, cg)
            for g in range(4):
                out[t * 4 * cw + g * cw: t * 4 * cw + g * cw + (c1 - c0)] = v[c0 * 4 + g: c1 * 4: 4]
    elif shuffle_meta and "group_n" in shuffle_meta:
        gn, cg = shuffle_meta["group_n"], shuffle_meta["cg"]
        for g in range(4):
            out[g * gn:g * gn + cg] = v[g::4]
    else:
        out[:v.numel()] = v
    return out


def choose_bn(n_total: int, m_tiles: int, sms: int = 148, ksteps: int = 36, max_bn: int = 256) -> int:
    """N tile (UMMA N: multiple of 16, <= 256; 272..320 = 256 + rest for the res_block tail) from a small cost
    model fitted to measurements on B200: a tile costs ksteps * t_k(BN) + a fixed prologue/epilogue share, the
    launch costs ceil(tiles / SMs) waves of that.  Wide tiles re-read the activations least; narrow tiles fill the
    machine when the layer has few pixels."""
    if 256 < n_total <= 320 and max_bn >= 256:
        return n_total
    cands = [bn for bn in range(max_bn, 63, -16) if n_total % bn == 0] or [min(n_total, max_bn)]
    if n_total <= max_bn and n_total not in cands:
        cands = [n_total] + cands
    best, best_cost = cands[0], None
    for bn in cands:
        tiles = m_tiles * -(-n_total // bn)
        waves = -(-tiles // sms)
        t_k = max(0.22, 0.5 * bn / 256.0)                 # us per 64-deep K step (smem/L2 floor below BN~112)
        cost = waves * (ksteps * t_k + 2.0 + 3.0 * bn / 256.0)
        if best_cost is None or cost < best_cost * 0.98:  # prefer the wider tile on (near) ties
            best, best_cost = bn, cost
    return best


@dataclass
class ConvOp:
    """A fully specified havc_conv_gemm launch (descriptor + the tensors it points at)."""
    desc: ConvDesc
    keep: list = field(default_factory=list)
    flops: float = 0.0   # algorithmic (unpadded) FLOPs, filled by the planner
    name: str = ""

    def launch(self, stream: Optional[int] = None):
        rc = _lib.lib().havc_conv_gemm(C.byref(self.desc), C.c_void_p(stream or 0))
        _lib.check(rc, f"havc_conv_gemm[{self.name}]")


def make_conv(src0: torch.Tensor, weight: torch.Tensor, out: torch.Tensor, taps, *, src1: Optional[torch.Tensor] = None,
              w_c1_off: int = 0, n_total: Optional[int] = None, bn: Optional[int] = None,
              bias=None, scale=None, shift=None, relu1=False, relu2=False, residual=None,
              out_space: Optional[Tuple[int, int, int]] = None, box=None, a_batched=True, b_batched=False,
              up=1, oy=0, ox=0, shuffle=False, group_n=0, c_store: Optional[int] = None,
              src1_single_tap: bool = False, src1_wi: int = 0, split_n: int = 0, out2: Optional[torch.Tensor] = None,
              c_store2: int = 0, residual2: Optional[torch.Tensor] = None, head_w: Optional[torch.Tensor] = None,
              head_out: Optional[torch.Tensor] = None, tma_store: Optional[bool] = None, leaky1: float = 0.0,
              pair: int = 0, name: str = "", src0_lo: Optional[torch.Tensor] = None, src1_lo: Optional[torch.Tensor] = None,
              weight_lo: Optional[torch.Tensor] = None, out_lo: Optional[torch.Tensor] = None,
              residual_lo: Optional[torch.Tensor] = None, blur: bool = False) -> ConvOp:
    """src0/src1: NHWC (or [P,B,H,W,C]) 16-bit device tensors; weight: packed [rows,taps,cin] or
    [batches,rows,taps,cin]; out: NHWC tensor written at pixel (h*up+oy, w*up+ox)."""
    d = ConvDesc()
    d.dtype = havc_dtype(src0.dtype)
    d.src0 = act_view(src0)
    if src1 is not None:
        d.src1 = act_view(src1)
    if weight.dim() == 3:
        wb, (rows, wt, cin) = 1, weight.shape
    else:
        wb, rows, wt, cin = weight.shape
    assert weight.is_contiguous()
    d.weight = weight.data_ptr()
    d.w_rows, d.w_taps, d.w_cin, d.w_batches = rows, wt, cin, wb
    d.w_c1_off = w_c1_off
    d.n_taps = len(taps)
    for i, (dh, dw, p, wi) in enumerate(taps):
        d.tap_dh[i], d.tap_dw[i], d.tap_p[i], d.tap_wi[i] = dh, dw, p, wi
    if out_space is None:
        B, H, W = src0.shape[-4], src0.shape[-3], src0.shape[-2]
    else:
        B, H, W = out_space
    d.out_B, d.out_H, d.out_W = B, H, W
    if box is None:
        box = choose_box(W, H, B if (a_batched and not b_batched) else 1)
        if not a_batched or b_batched:
            box = choose_box(W, H, 1)
    d.box_w, d.box_h, d.box_b = box
    d.a_batched, d.b_batched = int(a_batched), int(b_batched)
    d.N_total = n_total if n_total is not None else rows
    m_tiles = -(-W // box[0]) * -(-H // box[1]) * -(-B // box[2])
    ksteps = len(taps) * (-(-d.src0.C // 64) + (-(-d.src1.C // 64) if src1 is not None and not src1_single_tap else 0))
    d.BN = bn if bn is not None else choose_bn(d.N_total, m_tiles, ksteps=ksteps, max_bn=128 if out_lo is not None else 256)
    n_alloc = -(-d.N_total // d.BN) * d.BN
    keep = [src0, src1, weight, out, residual]
    for nm, v in (("bias", bias), ("scale", scale), ("shift", shift)):
        if v is not None:
            assert v.dtype == torch.float32 and v.numel() >= d.N_total, nm
            setattr(d, nm, v.data_ptr())
            keep.append(v)
    d.relu1, d.relu2 = int(relu1), int(relu2)
    d.leaky1 = float(leaky1)
    if residual is not None:
        assert residual.dtype == src0.dtype and residual.dim() == 4
        d.residual = residual.data_ptr()
        d.res_stride_b, d.res_stride_h, d.res_stride_w = residual.stride()[:3]
    if out is not None:
        d.out = out.data_ptr()
        d.out_dtype = havc_dtype(out.dtype)
        assert out.dim() == 4 and out.stride(-1) == 1
        d.out_stride_b, d.out_stride_h, d.out_stride_w = out.stride()[:3]
    else:
        assert head_w is not None, "a launch without an output tensor needs the fused head"
        d.out_dtype = d.dtype
    d.up, d.oy, d.ox = up, oy, ox
    d.shuffle, d.group_n = int(bool(shuffle)), group_n
    d.blur = int(blur)
    d.c_store = c_store if c_store is not None else out.shape[-1]
    d.src1_single_tap, d.src1_wi, d.split_n = int(src1_single_tap), src1_wi, split_n
    if out2 is not None:
        assert out2.dim() == 4 and out2.stride(-1) == 1 and out2.dtype == out.dtype
        d.out2 = out2.data_ptr()
        d.out2_stride_b, d.out2_stride_h, d.out2_stride_w = out2.stride()[:3]
    d.c_store2 = c_store2
    if residual2 is not None:
        assert residual2.dtype == src0.dtype and residual2.dim() == 4
        d.residual2 = residual2.data_ptr()
        d.res2_stride_b, d.res2_stride_h, d.res2_stride_w = residual2.stride()[:3]
    if head_w is not None:
        assert head_w.dtype == torch.float32 and head_w.numel() == 3 * d.N_total and head_out.dtype == torch.float32
        assert head_out.dim() == 4 and head_out.shape[-1] == 4
        d.head_w, d.head_out = head_w.data_ptr(), head_out.data_ptr()
        d.head_stride_b, d.head_stride_h, d.head_stride_w = head_out.stride()[:3]
    d.tma_store = int(TMA_STORE_DEFAULT if tma_store is None else tma_store)
    d.pair = int(pair)     # 0 = library default (CTA pairs on), 1 = force, -1 = single-CTA tiles
    # split-precision planes: same geometry as their hi tensors
    for nm, lo, hi in (("src0_lo", src0_lo, src0), ("src1_lo", src1_lo, src1), ("weight_lo", weight_lo, weight),
                       ("out_lo", out_lo, out), ("residual_lo", residual_lo, residual)):
        if lo is not None:
            assert hi is not None and lo.shape == hi.shape and lo.stride() == hi.stride() and lo.dtype == hi.dtype, nm
            setattr(d, nm, lo.data_ptr())
    keep += [out2, residual2, head_w, head_out, src0_lo, src1_lo, weight_lo, out_lo, residual_lo]
    return ConvOp(d, keep, name=name)
