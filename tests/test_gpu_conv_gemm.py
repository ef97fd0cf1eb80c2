"""GPU parity tests of the tcgen05 implicit-GEMM kernel (through the C ABI) against a plain torch fp32
reference of the same op (F.conv2d / matmul with TF32 disabled) on identical 16-bit-rounded inputs.

Tolerance: the kernel accumulates in fp32 and rounds ONCE to the 16-bit output type, so
|err| <= 2^-10 * |y| (fp16) or 2^-7 * |y| (bf16) plus fp32 summation-order noise.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)


def _tol(dtype):
    return 2.0 ** -9 if dtype == torch.float16 else 2.0 ** -6


def _nhwc(x):  # NCHW fp32 -> NHWC
    return x.permute(0, 2, 3, 1).contiguous()


def _check(got, ref, dtype, what=""):
    got = got.float()
    err = (got - ref).abs()
    bound = _tol(dtype) * ref.abs() + _tol(dtype) * ref.abs().mean() + 1e-4
    bad = (err > bound).sum().item()
    assert bad == 0, f"{what}: {bad}/{err.numel()} mismatches, max err {err.max().item():.4g}, ref absmax {ref.abs().max().item():.4g}"


def _run_conv(B, H, W, cins, Cout, ks, dtype, *, bias=False, relu1=False, affine=False, residual=False, relu2=False,
              shuffle=False, dilation=1, bn=None, box=None, pair=0):
    from vsdeoldify_b200 import ops
    _setup()
    dev = "cuda"
    Cin = sum(cins)
    xs = [torch.randn(B, c, H, W, device=dev) for c in cins]
    w = torch.randn(Cout, Cin, ks, ks, device=dev) / (Cin * ks * ks) ** 0.5
    xs16 = [x.to(dtype) for x in xs]
    w16 = w.to(dtype)
    xcat = torch.cat([x.float() for x in xs16], 1)
    ref = F.conv2d(xcat, w16.float(), padding=dilation * (ks - 1) // 2, dilation=dilation)
    bvec = torch.randn(Cout, device=dev) if bias else None
    svec = (torch.rand(Cout, device=dev) + 0.5) if affine else None
    tvec = torch.randn(Cout, device=dev) if affine else None
    if bias:
        ref = ref + bvec[None, :, None, None]
    if relu1:
        ref = ref.relu()
    if affine:
        ref = ref * svec[None, :, None, None] + tvec[None, :, None, None]
    res16 = None
    if residual:
        res = torch.randn(B, Cout, H, W, device=dev)
        res16 = res.to(dtype)
        ref = ref + res16.float()
    if relu2:
        ref = ref.relu()
    if shuffle:
        ref = F.pixel_shuffle(ref, 2)

    # device-side operands
    srcs = []
    for x16, c in zip(xs16, cins):
        cp = ops.pad_to(c, 8)
        t = torch.zeros(B, H, W, cp, device=dev, dtype=dtype)
        t[..., :c] = _nhwc(x16)
        srcs.append(t)
    wp, meta = ops.pack_conv_weight(w16.float().cpu(), cins, dtype=dtype, shuffle=shuffle)
    wp = wp.to(dev)
    n_total = meta["rows"]
    pc = lambda v, fill: None if v is None else ops.pack_cols(v, n_total, fill, meta if shuffle else None).to(dev)
    if shuffle:
        cg = Cout // 4
        out = torch.full((B, 2 * H, 2 * W, ops.pad_to(cg, 8)), float("nan"), device=dev, dtype=dtype)
    else:
        out = torch.full((B, H, W, ops.pad_to(Cout, 8)), float("nan"), device=dev, dtype=dtype)
    rs = None
    if residual:
        rs = torch.zeros(B, H, W, ops.pad_to(Cout, 8), device=dev, dtype=dtype)
        rs[..., :Cout] = _nhwc(res16)
    op = ops.make_conv(srcs[0], wp, out, ops.taps_for(ks, dilation), src1=srcs[1] if len(srcs) > 1 else None,
                       w_c1_off=meta["c1_off"], n_total=n_total, bn=bn, box=box,
                       bias=pc(bvec, 0.0), scale=pc(svec, 1.0), shift=pc(tvec, 0.0),
                       relu1=relu1, relu2=relu2, residual=rs, shuffle=shuffle, group_n=meta.get("group_n", 0), pair=pair)
    op.launch()
    torch.cuda.synchronize()
    cvalid = Cout // 4 if shuffle else Cout
    got = out[..., :cvalid].permute(0, 3, 1, 2)
    _check(got, ref, dtype, f"conv {cins}->{Cout} k{ks} {H}x{W}x{B}")
    pad = out[..., cvalid:]
    if pad.numel():
        assert torch.isfinite(pad.float()).all() and (pad == 0).all(), "pad channels must be written as zeros"


def test_conv1x1_minimal():
    _run_conv(1, 8, 16, [64], 64, 1, torch.float16)


def test_conv1x1_k256():
    _run_conv(2, 16, 16, [256], 128, 1, torch.float16)


def test_conv3x3_basic():
    _run_conv(2, 16, 16, [64], 64, 3, torch.float16)


def test_conv3x3_bias_relu_bf16():
    _run_conv(2, 24, 24, [128], 128, 3, torch.bfloat16, bias=True, relu1=True)


def test_conv3x3_two_sources_affine():
    _run_conv(2, 24, 24, [96, 40], 64, 3, torch.float16, relu1=True, affine=True)


def test_conv3x3_residual_relu():
    _run_conv(2, 16, 16, [64], 256, 3, torch.float16, bias=True, residual=True, relu2=True)


def test_conv3x3_cout259_bn272():
    _run_conv(1, 32, 32, [256, 3], 259, 3, torch.float16, bias=True, relu1=True)


def test_conv1x1_shuffle():
    _run_conv(2, 12, 12, [128], 256, 1, torch.float16, bias=True, relu1=True, shuffle=True)


def test_conv1x1_shuffle_odd_group():
    _run_conv(1, 16, 16, [64], 4 * 75, 1, torch.float16, bias=True, relu1=True, shuffle=True)


def test_conv3x3_small_spatial_batch_box():
    _run_conv(8, 12, 12, [128], 512, 3, torch.float16, relu1=True, affine=True)


def test_conv3x3_dilated():
    _run_conv(1, 32, 32, [64], 64, 3, torch.float16, dilation=2)


def test_conv3x3_many_tiles_persistent():
    # 4*96*96/128 = 288 m-tiles x 2 n-tiles > 148 SMs: exercises the persistent loop, both
    # accumulator stages and the smem ring wrap-around.
    _run_conv(4, 96, 96, [256], 512, 3, torch.float16, relu1=True, affine=True)


def test_conv_stride2_phase_split():
    from vsdeoldify_b200 import ops
    _setup()
    dev, dtype = "cuda", torch.float16
    B, H, W, Cin, Cout = 2, 32, 32, 64, 128
    x = torch.randn(B, Cin, H, W, device=dev).to(dtype)
    w = (torch.randn(Cout, Cin, 3, 3, device=dev) / (Cin * 9) ** 0.5).to(dtype)
    ref = F.conv2d(x.float(), w.float(), stride=2, padding=1)
    xn = _nhwc(x)  # [B,H,W,C]
    ph = torch.stack([xn[:, a::2, b::2] for a in range(2) for b in range(2)], 0).contiguous()  # [4,B,H/2,W/2,C]
    wp, meta = ops.pack_conv_weight(w.float().cpu(), dtype=dtype)
    out = torch.zeros(B, H // 2, W // 2, Cout, device=dev, dtype=dtype)
    op = ops.make_conv(ph, wp.to(dev), out, ops.taps_stride2(3), n_total=meta["rows"])
    op.launch()
    torch.cuda.synchronize()
    _check(out.permute(0, 3, 1, 2), ref, dtype, "stride-2 conv")


def test_batched_gemm_fp32_out():
    """S[b, j, i] = sum_c g[b, j, c] f[b, i, c]  (the attention logits: both operands per-image)."""
    from vsdeoldify_b200 import ops
    _setup()
    dev, dtype = "cuda", torch.float16
    B, N, d = 3, 320, 64
    g = torch.randn(B, N, d, device=dev).to(dtype)
    f = torch.randn(B, N, d, device=dev).to(dtype)
    ref = torch.einsum("bjc,bic->bji", g.float(), f.float())
    out = torch.zeros(B, 1, N, N, device=dev, dtype=torch.float32)
    op = ops.make_conv(g.view(B, 1, N, d), f.view(B, N, 1, d), out, [(0, 0, 0, 0)], n_total=N, b_batched=True,
                       out_space=(B, 1, N))
    op.launch()
    torch.cuda.synchronize()
    err = (out.view(B, N, N) - ref).abs().max().item()
    assert err < 1e-3, err


def test_gemm_shared_a_batched_b():
    """Ht[b, c, i] = sum_k Wv[c, k] x[b, i, k]  (value projection written channel-major)."""
    from vsdeoldify_b200 import ops
    _setup()
    dev, dtype = "cuda", torch.float16
    B, N, Cc = 2, 192, 128
    wv = (torch.randn(Cc, Cc, device=dev) / Cc ** 0.5).to(dtype)
    x = torch.randn(B, N, Cc, device=dev).to(dtype)
    ref = torch.einsum("ck,bik->bci", wv.float(), x.float())
    out = torch.zeros(B, 1, Cc, N, device=dev, dtype=dtype)
    op = ops.make_conv(wv.view(1, 1, Cc, Cc), x.view(B, N, 1, Cc), out, [(0, 0, 0, 0)], n_total=N,
                       a_batched=False, b_batched=True, out_space=(B, 1, Cc))
    op.launch()
    torch.cuda.synchronize()
    _check(out.view(B, Cc, N), ref, dtype, "shared-A batched-B gemm")


def test_conv_k24_single_kstep_many_tiles():
    """Zhang model1.0 shape: K = 24 (one partial 64-chunk), BN = 64, ~7 tiles per CTA."""
    _run_conv(2, 256, 256, [24], 64, 1, torch.float16, bias=True, relu1=True)


def test_conv3x3_c64_large_map():
    _run_conv(2, 256, 256, [64], 64, 3, torch.float16, bias=True, relu1=True, affine=True)


def test_conv3x3_c64_medium_map():
    _run_conv(2, 128, 128, [64], 64, 3, torch.float16, bias=True, relu1=True, affine=True)


# ---- CTA pairs (tcgen05.mma.cta_group::2): forced on (the library's own policy only pairs long launches) ------------------
@pytest.mark.parametrize("pair", [1, -1])
def test_pair_conv3x3_many_tiles(pair):
    _run_conv(4, 96, 96, [256], 512, 3, torch.float16, relu1=True, affine=True, pair=pair)


def test_pair_cout259_bn272_staggered_two_part_mma():
    # BN = 272 = 144 + 128: two cta_group::2 MMAs per K step, 72 + 64 weight rows per CTA, staggered accumulators
    _run_conv(2, 32, 32, [256, 3], 259, 3, torch.float16, bias=True, relu1=True, pair=1)
    _run_conv(3, 40, 24, [256, 3], 259, 3, torch.bfloat16, bias=True, relu1=True, residual=True, pair=1)


def test_pair_odd_tile_count_and_ragged_edges():
    # 3*20*20 = 1200 pixels -> 10 M tiles of a 20x... box grid with clipped edges; odd counts leave the peer CTA an empty tile
    _run_conv(3, 20, 20, [64], 64, 3, torch.float16, bias=True, relu1=True, pair=1)
    _run_conv(1, 24, 40, [128], 128, 3, torch.float16, residual=True, relu2=True, pair=1)
    _run_conv(5, 12, 12, [128], 512, 3, torch.float16, relu1=True, affine=True, pair=1)


def test_pair_shuffle_and_residual_fast_epilogue():
    _run_conv(2, 24, 24, [128], 256, 1, torch.float16, bias=True, relu1=True, shuffle=True, pair=1)
    _run_conv(2, 48, 48, [64], 256, 1, torch.float16, bias=True, residual=True, relu2=True, pair=1)
    _run_conv(2, 64, 64, [64], 64, 3, torch.float16, bias=True, relu1=True, affine=True, pair=1)


def test_pair_batched_gemm_same_batch_only():
    """b_batched GEMMs pair only when both M tiles of a pair lie in the same batch (2 tiles per batch here)."""
    from vsdeoldify_b200 import ops
    _setup()
    dev = "cuda"
    B, N, d = 3, 256, 64
    f = torch.randn(B, N, d, device=dev).half()
    g = torch.randn(B, N, d, device=dev).half()
    ref = torch.bmm(g.float(), f.float().transpose(1, 2))
    for pair in (1, -1):
        out = torch.zeros(B, 1, N, N, device=dev, dtype=torch.float32)
        op = ops.make_conv(g.view(B, 1, N, d), f.view(B, N, 1, d), out, [(0, 0, 0, 0)], n_total=N, b_batched=True,
                           out_space=(B, 1, N), pair=pair)
        op.launch()
        torch.cuda.synchronize()
        err = (out.view(B, N, N) - ref).abs().max().item()
        assert err < 1e-3, (pair, err)


# ---- fast epilogue with the TMA-loaded residual tile (staging buffer doubles as the residual source) ---------------------
@pytest.mark.parametrize("pair", [-1, 1])
def test_residual_tma_fast_epilogue_shapes(pair):
    # ResNet bottleneck conv3 shapes: 1x1 + BN + residual + ReLU; BN tile 64 / 128 / 256, several tiles per CTA, ragged edges
    _run_conv(2, 48, 48, [64], 256, 1, torch.float16, affine=True, residual=True, relu2=True, pair=pair)
    _run_conv(4, 96, 96, [64], 256, 1, torch.float16, affine=True, residual=True, relu2=True, pair=pair)      # > 148 tiles
    _run_conv(3, 20, 28, [128], 512, 1, torch.bfloat16, affine=True, residual=True, relu2=True, pair=pair)    # clipped boxes
    _run_conv(2, 24, 24, [256], 64, 1, torch.float16, bias=True, residual=True, relu2=True, pair=pair)
    _run_conv(2, 24, 24, [64], 128, 3, torch.float16, bias=True, residual=True, pair=pair)                      # no final ReLU


# ---- split-precision ("x3") launches: value = hi + lo planes, products hi*hi + lo*hi + hi*lo -------------------------------
def _run_conv_x3(B, H, W, cins, Cout, ks, dtype, *, bias=True, relu1=True, affine=False, residual=False, relu2=False,
                 stride=1, pair=0, a_lo=True, w_lo=True):
    """fp32 inputs / weights, compared with an fp64 reference: the result (hi + lo of the output Pair) must be fp32-class."""
    from vsdeoldify_b200 import ops
    _setup()
    dev = "cuda"
    Cin = sum(cins)
    xs = [torch.randn(B, c, H, W, device=dev) for c in cins]
    w = torch.randn(Cout, Cin, ks, ks, device=dev) / (Cin * ks * ks) ** 0.5
    q = lambda t: t if a_lo else t.to(dtype).float()
    qw = lambda t: t if w_lo else t.to(dtype).float()
    xcat = torch.cat([q(x) for x in xs], 1).double()
    ref = F.conv2d(xcat, qw(w).double(), padding=(ks - 1) // 2, stride=stride)
    bvec = torch.randn(Cout, device=dev) if bias else None
    svec = (torch.rand(Cout, device=dev) + 0.5) if affine else None
    tvec = torch.randn(Cout, device=dev) if affine else None
    if bias:
        ref = ref + bvec.double()[None, :, None, None]
    if relu1:
        ref = ref.relu()
    if affine:
        ref = ref * svec.double()[None, :, None, None] + tvec.double()[None, :, None, None]
    OH, OW = ref.shape[2], ref.shape[3]
    res = None
    if residual:
        res = torch.randn(B, Cout, OH, OW, device=dev)
        ref = ref + res.double()
    if relu2:
        ref = ref.relu()

    def pair_of(x, c):          # NCHW fp32 -> Pair [2,B,H,W,cp]
        cp = ops.pad_to(c, 8)
        t = torch.zeros(2, x.shape[0], x.shape[2], x.shape[3], cp, device=dev, dtype=dtype)
        hi, lo = ops.split_hi_lo(_nhwc(x), dtype)
        t[0, ..., :c], t[1, ..., :c] = hi, lo
        return ops.Pair(t)
    srcs = [pair_of(x, c) for x, c in zip(xs, cins)]
    g = 1024.0                                           # power-of-two weight pre-scale, undone by the epilogue scale
    w32, meta = ops.pack_conv_weight((w.double() * g).cpu(), cins, dtype=None)
    whi, wlo = ops.split_hi_lo(w32, dtype)
    whi, wlo = whi.to(dev), wlo.to(dev)
    n_total = meta["rows"]
    pc = lambda v, fill: None if v is None else ops.pack_cols(v, n_total, fill).to(dev)
    scale = (svec if affine else torch.ones(Cout, device=dev)) / g
    shift = tvec if affine else torch.zeros(Cout, device=dev)
    out = ops.Pair(torch.full((2, B, OH, OW, ops.pad_to(Cout, 8)), float("nan"), device=dev, dtype=dtype))
    rs = pair_of(res, Cout) if residual else None
    if stride == 2:
        # phase-split inputs [P,B,H/2,W/2,C] per plane
        def split(t):
            P = torch.zeros(2, 4, B, (H + 1) // 2, (W + 1) // 2, t.shape[-1], device=dev, dtype=dtype)
            for a in (0, 1):
                for b in (0, 1):
                    v = t.t[:, :, a::2, b::2]
                    P[:, a * 2 + b, :, :v.shape[2], :v.shape[3]] = v
            return ops.Pair(P)
        srcs = [split(s) for s in srcs]
        taps = ops.taps_stride2(ks) if ks > 1 else [(0, 0, 0, 0)]
    else:
        taps = ops.taps_for(ks)
    op = ops.make_conv(srcs[0].hi, whi, out.hi, taps, src1=srcs[1].hi if len(srcs) > 1 else None, w_c1_off=meta["c1_off"],
                       n_total=n_total, bias=pc(bvec * g if bias else None, 0.0), scale=pc(scale, 1.0), shift=pc(shift, 0.0),
                       relu1=relu1, relu2=relu2, residual=rs.hi if rs is not None else None, out_space=(B, OH, OW), pair=pair,
                       src0_lo=srcs[0].lo if a_lo else None, src1_lo=srcs[1].lo if (a_lo and len(srcs) > 1) else None,
                       weight_lo=wlo if w_lo else None, out_lo=out.lo, residual_lo=rs.lo if rs is not None else None)
    op.launch()
    torch.cuda.synchronize()
    got = out.float()[..., :Cout].permute(0, 3, 1, 2).double()
    err = (got - ref).abs()
    scale_ref = float(ref.abs().mean()) + 1e-6
    # hi + lo of fp16 carries ~2^-21 per operand and the three-product contraction drops lo*lo (2^-22 relative per term): the
    # error of a K-term sum grows like sqrt(K) * 2^-21 (measured 4.5e-6 of the output range at K = 1152); plain fp16 operands
    # sit at ~5e-4, so the bound still separates the two by 30x
    tol = 1.5e-5 if dtype == torch.float16 else 3e-4
    assert float(err.max()) <= tol * (float(ref.abs().max()) + scale_ref), \
        f"x3 conv {cins}->{Cout} k{ks} s{stride}: max err {float(err.max()):.3e} (ref absmax {float(ref.abs().max()):.3g})"
    pad = out.t[..., Cout:]
    if pad.numel():
        assert (pad == 0).all(), "pad channels must be written as zeros"


def test_x3_conv1x1_bias_relu():
    _run_conv_x3(2, 16, 16, [64], 64, 1, torch.float16)


def test_x3_conv1x1_wide_residual_relu2():
    _run_conv_x3(2, 24, 24, [256], 1024, 1, torch.float16, relu1=False, residual=True, relu2=True)


def test_x3_conv3x3_pair_and_single():
    _run_conv_x3(4, 24, 24, [128], 128, 3, torch.float16, pair=1)
    _run_conv_x3(4, 24, 24, [128], 128, 3, torch.float16, pair=-1)


def test_x3_conv3x3_stride2():
    _run_conv_x3(2, 32, 32, [64], 128, 3, torch.float16, stride=2)


def test_x3_conv3x3_affine_two_sources():
    _run_conv_x3(2, 16, 16, [64, 64], 128, 3, torch.float16, affine=True)


def test_x3_partial_planes():
    """Only one of the lo planes present: the launch equals the contraction of the operands it was given."""
    _run_conv_x3(2, 16, 16, [64], 64, 3, torch.float16, a_lo=False)
    _run_conv_x3(2, 16, 16, [64], 64, 3, torch.float16, w_lo=False)


def test_x3_bf16():
    _run_conv_x3(2, 16, 16, [64], 64, 3, torch.bfloat16)


# ---- fused PixelShuffle + ICNR blur epilogue ---------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,Cin,Cout,dtype", [(2, 48, 48, 64, 4 * 64, torch.float16), (1, 50, 37, 128, 4 * 100, torch.float16),
                                                  (3, 96, 96, 64, 4 * 256, torch.bfloat16)])
def test_shuffle_blur_fused_equals_unfused_and_torch(B, H, W, Cin, Cout, dtype):
    """conv1x1 + bias + ReLU -> PixelShuffle(2) -> ReplicationPad2d((1,0,1,0)) + AvgPool2d(2, 1) (CustomPixelShuffle_ICNR,
    unet.py:24-52) as ONE launch (havc_conv_desc.blur: halo recompute, blur in the epilogue) against the two-launch path
    (PixelShuffle store + havc_blur2x2: bit-identical, same roundings in the same order) and against torch."""
    from vsdeoldify_b200 import _lib, ops
    _setup()
    dev = "cuda"
    x = torch.randn(B, Cin, H, W, device=dev)
    w = torch.randn(Cout, Cin, 1, 1, device=dev) / Cin ** 0.5
    bias = torch.randn(Cout, device=dev)
    x16, w16 = x.to(dtype), w.to(dtype)
    y = F.pixel_shuffle(F.relu(F.conv2d(x16.float(), w16.float()) + bias[None, :, None, None]), 2)
    ref = F.avg_pool2d(F.pad(y.to(dtype).float(), (1, 0, 1, 0), mode="replicate"), 2, stride=1)
    cg = Cout // 4
    src = _nhwc(x16)
    hd = ops.havc_dtype(dtype)
    # two launches
    wp, meta = ops.pack_conv_weight(w16.float().cpu(), None, dtype=dtype, shuffle=True)
    t = torch.zeros(B, 2 * H, 2 * W, ops.chan_storage(cg), device=dev, dtype=dtype)
    ops.make_conv(src, wp.to(dev), t, ops.taps_for(1), n_total=meta["rows"], bias=ops.pack_cols(bias, meta["rows"], 0.0, meta).to(dev),
                  relu1=True, shuffle=True, group_n=meta["group_n"], c_store=ops.pad_to(cg, 8)).launch()
    unfused = torch.zeros_like(t)
    _lib.check(_lib.lib().havc_blur2x2(t.data_ptr(), unfused.data_ptr(), B, 2 * H, 2 * W, t.shape[-1], unfused.stride(2), hd, None))
    # one launch
    wpb, mb = ops.pack_conv_weight(w16.float().cpu(), None, dtype=dtype, shuffle="blur")
    fused = torch.zeros_like(t)
    for pair in (-1, 1):
        fused.zero_()
        ops.make_conv(src, wpb.to(dev), fused, ops.taps_for(1), n_total=mb["rows"], bn=4 * ops.BLUR_CW, box=(16, 8, 1),
                      bias=ops.pack_cols(bias, mb["rows"], 0.0, mb).to(dev), relu1=True, shuffle=True, blur=True,
                      c_store=ops.pad_to(cg, 8), pair=pair).launch()
        torch.cuda.synchronize()
        assert torch.equal(fused, unfused), f"pair={pair}: {(fused != unfused).sum().item()} of {fused.numel()} values differ"
    _check(fused[..., :cg].permute(0, 3, 1, 2), ref, dtype, "shuffle+blur")
    assert (fused[..., cg:] == 0).all()


def test_fp16_outputs_saturate_instead_of_overflowing():
    """fp16 tops out at 65504: an epilogue value beyond it is stored as +-65504 (F2FP.SATFINITE), not as inf - an inf would become
    NaN for the whole receptive field in the next layer (0 * inf, inf - inf).  Values inside the range are untouched."""
    from vsdeoldify_b200 import ops
    _setup()
    dev = "cuda"
    B, H, W, C = 1, 16, 16, 64
    x = torch.full((B, H, W, C), 30.0, device=dev, dtype=torch.float16)
    w = torch.zeros(64, C, 1, 1)
    w[0, :, 0, 0] = 40.0            # 64 * 30 * 40 = 76800 > 65504
    w[1, :, 0, 0] = -40.0
    w[2, :, 0, 0] = 1.0             # 1920: representable
    wp, meta = ops.pack_conv_weight(w, None, dtype=torch.float16)
    out = torch.zeros(B, H, W, 64, device=dev, dtype=torch.float16)
    ops.make_conv(x, wp.to(dev), out, ops.taps_for(1), n_total=meta["rows"]).launch()
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert (out[..., 0] == 65504).all() and (out[..., 1] == -65504).all() and (out[..., 2] == 1920).all() and (out[..., 3:] == 0).all()
