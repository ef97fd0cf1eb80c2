"""CPU tests of the host-side logic: weight packing / tap tables / concat layout (checked by emulating the
implicit GEMM in plain torch and comparing with F.conv2d), tile/box heuristics, resampling tables, BN folding,
frame partitioning (incl. a world_size-2 gloo run)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def emulate_conv(srcs, wp, taps, c1_off, out_hw, phase_split=False):
    """Reference semantics of havc_conv_gemm's mainloop in torch: srcs NHWC (or [P,B,H,W,C]) fp32, wp [rows,taps,cin]."""
    rows = wp.shape[0]
    B = srcs[0].shape[-4]
    H, W = out_hw
    acc = torch.zeros(B, H, W, rows)
    for (dh, dw, p, wi) in taps:
        off = 0
        for si, s in enumerate(srcs):
            x = s[p] if s.dim() == 5 else s
            C = x.shape[-1]
            xp = F.pad(x, (0, 0, 8, 8, 8, 8))                                   # zero fill outside the image (TMA OOB)
            win = xp[:, 8 + dh:8 + dh + H, 8 + dw:8 + dw + W, :]
            wsl = wp[:, wi, off:off + C] if si == 0 else wp[:, wi, c1_off:c1_off + C]
            acc += torch.einsum("bhwc,nc->bhwn", win, wsl)
            off += C
    return acc


@pytest.mark.parametrize("cins,ks,dil", [([64], 3, 1), ([96, 40], 3, 1), ([24], 1, 1), ([16], 3, 2), ([300, 3], 3, 1)])
def test_pack_and_taps_stride1(cins, ks, dil):
    from vsdeoldify_b200 import ops
    torch.manual_seed(0)
    B, H, W, Cout = 2, 9, 11, 20
    xs = [torch.randn(B, c, H, W) for c in cins]
    w = torch.randn(Cout, sum(cins), ks, ks)
    ref = F.conv2d(torch.cat(xs, 1), w, padding=dil * (ks - 1) // 2, dilation=dil)
    wp, meta = ops.pack_conv_weight(w, cins, dtype=torch.float32)
    srcs = []
    for x, c in zip(xs, cins):
        t = torch.zeros(B, H, W, ops.pad_to(c, 8))
        t[..., :c] = x.permute(0, 2, 3, 1)
        srcs.append(t)
    got = emulate_conv(srcs, wp, ops.taps_for(ks, dil), meta["c1_off"], (H, W))
    assert torch.allclose(got[..., :Cout].permute(0, 3, 1, 2), ref, atol=1e-4)
    assert (got[..., Cout:] == 0).all()


@pytest.mark.parametrize("ks", [1, 3])
def test_taps_stride2_phase_split(ks):
    from vsdeoldify_b200 import ops
    torch.manual_seed(1)
    B, H, W, Cin, Cout = 2, 12, 10, 16, 8
    x = torch.randn(B, Cin, H, W)
    w = torch.randn(Cout, Cin, ks, ks)
    ref = F.conv2d(x, w, stride=2, padding=(ks - 1) // 2)
    xn = x.permute(0, 2, 3, 1)
    ph = torch.stack([xn[:, a::2, b::2] for a in range(2) for b in range(2)], 0)
    wp, meta = ops.pack_conv_weight(w, dtype=torch.float32)
    taps = ops.taps_stride2(ks) if ks > 1 else [(0, 0, 0, 0)]
    got = emulate_conv([ph], wp, taps, 0, (H // 2, W // 2))
    assert torch.allclose(got[..., :Cout].permute(0, 3, 1, 2), ref, atol=1e-4)


def test_shuffle_row_order_matches_pixel_shuffle():
    from vsdeoldify_b200 import ops
    torch.manual_seed(2)
    B, H, W, Cin, Cout = 1, 4, 5, 8, 4 * 6
    x = torch.randn(B, Cin, H, W)
    w = torch.randn(Cout, Cin, 1, 1)
    ref = F.pixel_shuffle(F.conv2d(x, w), 2)                         # [B, 6, 2H, 2W]
    wp, meta = ops.pack_conv_weight(w, dtype=torch.float32, shuffle=True)
    acc = emulate_conv([x.permute(0, 2, 3, 1).contiguous()], wp, [(0, 0, 0, 0)], 0, (H, W))
    gn, cg = meta["group_n"], meta["cg"]
    out = torch.zeros(B, 2 * H, 2 * W, cg)
    for g in range(4):                                               # the kernel's store rule
        out[:, (g >> 1)::2, (g & 1)::2, :] = acc[..., g * gn:g * gn + cg]
    assert torch.allclose(out.permute(0, 3, 1, 2), ref, atol=1e-5)
    v = torch.arange(Cout, dtype=torch.float32)
    pv = ops.pack_cols(v, meta["rows"], -1.0, meta)
    assert pv[0] == 0 and pv[1] == 4 and pv[gn] == 1                  # row g*gn + c <- channel c*4 + g


def test_choose_box_and_bn():
    from vsdeoldify_b200 import ops
    for (W, H, B) in [(384, 384, 8), (12, 12, 8), (24, 24, 8), (15, 15, 4), (2, 2, 2), (48, 48, 1)]:
        bw, bh, bb = ops.choose_box(W, H, B)
        assert bw * bh * bb == 128
        tiles = -(-W // bw) * -(-H // bh) * -(-B // bb)
        assert tiles * 128 >= W * H * B
    assert ops.choose_box(12, 12, 8) == (4, 4, 8)                     # 3x3 boxes of 4x4 over 8 images: no waste
    assert ops.choose_bn(272, 9216) == 272 and ops.choose_bn(320, 9216) == 320   # res_block tails: one 256+rest tile
    assert ops.choose_bn(256, 2000) == 256                            # plenty of pixels: widest tile
    assert ops.choose_bn(256, 72) == 128                              # 72 pixel tiles on 148 SMs: split N to fill them
    for n, m in [(4096, 9), (512, 18), (1024, 4608), (2304, 18)]:
        bn = ops.choose_bn(n, m)
        assert 64 <= bn <= 256 and bn % 16 == 0 and n % bn == 0


def test_resample_tables_match_oracle_matrix():
    from oracle import pixel_oracle as px
    from vsdeoldify_b200 import resample
    for src, dst in [(1920, 384), (384, 1920), (1080, 384), (384, 1080), (64, 64), (100, 37)]:
        start, w = resample.build_tables(src, dst)
        dense = np.zeros((dst, src))
        for o in range(dst):
            for t in range(w.shape[1]):
                if start[o] + t < src:
                    dense[o, start[o] + t] += w[o, t]
        assert np.abs(dense - px.resize_matrix(src, dst)).max() < 1e-6
        assert (start >= 0).all() and (start + w.shape[1] <= src + w.shape[1]).all()


def test_bn_and_spectral_fold_match_torch():
    from oracle import synth_weights, unet_oracle
    from vsdeoldify_b200 import unet
    sd = synth_weights.make_unet_state_dict("deep", 7, calibrate=False)
    for p in ("layers.3.0.0", "layers.8.conv.0", "layers.10.layers.0.0", "layers.0.0"):
        assert torch.allclose(unet.folded_weight(sd, p), unet_oracle.conv_weight(sd, p), rtol=1e-5, atol=1e-7)
    sd["layers.1.running_mean"] = torch.randn(512)
    sd["layers.1.running_var"] = torch.rand(512) + 0.1
    x = torch.randn(2, 512, 3, 3)
    sc, sh = unet.bn_affine(sd, "layers.1")
    ref = F.batch_norm(x, sd["layers.1.running_mean"], sd["layers.1.running_var"], sd["layers.1.weight"], sd["layers.1.bias"],
                       False, 0.0, 1e-5)
    assert torch.allclose(x * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1), ref, atol=1e-5)


def test_block_partition_properties():
    from vsdeoldify_b200 import partition
    for n in (0, 1, 7, 2000, 2001):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s, e = partition.block_range(n, r, world)
                assert 0 <= s <= e <= n
                seen += list(range(s, e))
                for f in range(s, e):
                    assert partition.owner_of(f, n, world) == r
            assert seen == list(range(n))                             # disjoint, complete, in order
    assert list(partition.batches(3, 20, 8)) == [(3, 11), (11, 19), (19, 20)]


def _gloo_worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from vsdeoldify_b200 import partition
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = partition.block_range(n_frames, rank, world)
    mine = torch.arange(s, e, dtype=torch.int64) * 10 + 1             # stand-in for "rendered frame n"
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([e - s]))
    bufs = [torch.zeros(int(z), dtype=torch.int64) for z in sizes]
    dist.all_gather(bufs, mine) if len(set(int(z) for z in sizes)) == 1 else [dist.broadcast(bufs[r], r) if r != rank else dist.broadcast(mine, r) for r in range(world)]
    if len(set(int(z) for z in sizes)) != 1:
        bufs[rank] = mine
    if rank == 0:
        q.put(torch.cat(bufs).tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_gloo():
    """world_size 2 over gloo: each rank renders its block; concatenation in rank order is frame order."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_frames, world, port = 11, 2, 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [n * 10 + 1 for n in range(n_frames)]


def test_clip_adapter_recycles_only_unreferenced_result_arrays():
    """The clip adapter hands out frames whose planes are views of big result arrays; an array may be reused for a later batch
    only when no frame (view) references it any more."""
    from vsdeoldify_b200 import havc

    class FakeClip:
        height, width, num_frames = 4, 6, 10
    c = havc._ColorizedClip.__new__(havc._ColorizedClip)
    c.B, c.clip, c._bufs = 2, FakeClip(), []
    v = c._result_buf()[0, 1]                 # a frame plane keeps the first array alive and busy
    first = id(v.base)
    w = c._result_buf()[:1]
    assert id(w.base) != first and len(c._bufs) == 2
    del v
    assert id(c._result_buf()) == first       # free again -> recycled, no new allocation
    assert len(c._bufs) == 2


def test_legacy_spectral_norm_state_dict_folds_like_torch():
    """Spectral-norm state-dicts written before torch 1.0 (version < 1) hold `weight`, `weight_orig`, `weight_u` and no
    `weight_v`; torch's load hook (spectral_norm.py `_load_from_state_dict` -> `_solve_v_and_rescale`) derives v so that the
    eval weight is weight_orig / mean(weight_orig / weight).  folded_weight must agree with what torch itself loads."""
    import torch
    from torch import nn
    from vsdeoldify_b200.unet import folded_weight
    torch.manual_seed(5)
    conv = nn.utils.spectral_norm(nn.Conv2d(12, 20, 3, bias=False))
    conv.train()
    for _ in range(6):
        conv(torch.randn(2, 12, 8, 8))                       # power iterations: u, v converge, weight = weight_orig / sigma
    conv.eval()
    w_eff = conv(torch.zeros(1, 12, 8, 8)) is not None and conv.weight.detach().clone()
    sd_new = {"c." + k: v.clone() for k, v in conv.state_dict().items()}
    assert torch.allclose(folded_weight(sd_new, "c"), w_eff, atol=1e-6)
    # legacy schema: drop weight_v, add the materialised weight
    legacy = {k: v for k, v in sd_new.items() if not k.endswith("weight_v")}
    legacy["c.weight"] = w_eff.clone()
    got = folded_weight(legacy, "c")
    # what torch loads from the same legacy dict (no version metadata -> the version < 1 branch)
    conv2 = nn.utils.spectral_norm(nn.Conv2d(12, 20, 3, bias=False))
    conv2.load_state_dict({k[2:]: v for k, v in legacy.items()})
    conv2.eval()
    conv2(torch.zeros(1, 12, 8, 8))
    assert torch.allclose(got, conv2.weight.detach(), atol=1e-5, rtol=1e-4)
    assert torch.allclose(got, w_eff, atol=1e-5, rtol=1e-4)


def test_sharded_renderer_orders_frames_and_props_with_cpu_stand_in_engines():
    """sharded.ShardedRenderer (the in-process multi-GPU clip renderer) with CPU stand-ins for the engines: block and interleaved
    partitions, bounded and unbounded windows, in-order and out-of-order requests - every frame comes back under its own number
    with its own props and the bytes its engine produced."""
    import time
    import numpy as np
    from vsdeoldify_b200 import sharded, vs_shim

    class FakeEngine:
        n_slots = 2

        def __init__(self, B, H, W):
            self.bufs = [np.zeros((B, 3, H, W), np.uint8) for _ in range(2)]
            self.nxt, self.busy = 0, [False, False]

        def next_input(self):
            assert not self.busy[self.nxt]
            return self.bufs[self.nxt]

        def submit(self, frames, skip=None, n=None):
            s = self.nxt
            self.busy[s] = True
            self.nxt = (s + 1) % 2
            return (s, n, None if skip is None else np.asarray(skip).copy())

        def collect(self, t, out=None, pool=None):
            s, n, skip = t
            time.sleep(0.001)
            self.busy[s] = False
            out[:n] = 255 - self.bufs[s][:n]
            if skip is not None:
                out[:n][skip[:n]] = self.bufs[s][:n][skip[:n]]
            return out[:n]

    n, H, W, B = 53, 8, 12, 4
    frames = np.random.default_rng(0).integers(0, 256, (n, 3, H, W), dtype=np.uint8)
    clip = vs_shim.array_clip(frames, props=[{"_SceneChangePrev": int(i % 5 == 0), "id": i} for i in range(n)])
    assert sharded.plan_jobs(10, 2, 4, "block") == [(0, 4, 0), (4, 5, 0), (5, 9, 1), (9, 10, 1)]
    assert sharded.plan_jobs(10, 2, 4, "interleaved") == [(0, 4, 0), (4, 8, 1), (8, 10, 0)]
    for mode, win, sc in (("block", None, False), ("interleaved", None, False), ("interleaved", 3, False), ("interleaved", None, True)):
        r = sharded.ShardedRenderer(clip, [FakeEngine(B, H, W) for _ in range(3)], B, scenechange=sc, partition=mode, window=win)
        order = list(range(n)) if win is None else [0, 1, 2, 40, 41, 7, 8, 52, 3]
        for i in order:
            f = r(i)
            want = 255 - frames[i]
            if sc and not (i == 0 or i % 5 == 0):
                want = frames[i]                              # not a scene change: the gate leaves the frame uncoloured
            assert f.props["id"] == i and np.array_equal(np.stack([f[p] for p in range(3)]), want), (mode, win, sc, i)
        r.close()
    # a frame without the scene-detection prop is an error, not a silently uncoloured frame
    bare = vs_shim.array_clip(frames[:8], props=[{"id": i} for i in range(8)])
    r = sharded.ShardedRenderer(bare, [FakeEngine(B, H, W)], B, scenechange=True)
    with pytest.raises(KeyError):
        r(1)
    r.close()


def test_periodic_resampling_plans_are_bit_identical_to_the_tables():
    """resample.periodic_plan_down / _up (the phase-periodic horizontal kernels' host side): on the interior range the plan's
    weights and offsets reproduce the table formula out[o] = sum_t w[o, t] * in[start[o] + t] term by term (zero taps are
    exact no-ops of the fmaf chain), for the integer ratios of the BASELINE configs; other ratios have no plan."""
    import numpy as np
    from vsdeoldify_b200 import resample
    rng = np.random.default_rng(0)

    def fma_chain(ws, xs):
        acc = np.float32(0)
        for wv, xv in zip(ws, xs):
            acc = np.float32(np.float64(wv) * np.float64(xv) + np.float64(acc))      # an fp32 product is exact in fp64: this is fmaf
        return acc
    for src, dst in [(1920, 384), (1920, 480), (3840, 640)]:
        st, w = resample.build_tables(src, dst, "spline64")
        p = resample.periodic_plan_down(st, w, src, dst)
        assert p is not None and p["hi"] - p["lo"] >= dst // 8 - 4
        x = rng.uniform(0, 255, src).astype(np.float32)
        R, kTP = p["ratio"], resample.PERIODIC_DOWN_TAPS[p["ratio"]]
        for o in list(range(8 * p["lo"], 8 * p["lo"] + 16)) + list(range(8 * p["hi"] - 16, 8 * p["hi"])):
            a0 = p["offset"] + R * 8 * (o // 8) + R * (o % 8)
            xs = [x[a0 + t] if 0 <= a0 + t < src else np.float32(0) for t in range(kTP)]
            assert fma_chain(p["w"][:kTP], xs).view(np.int32) == fma_chain(w[o], x[st[o]:st[o] + w.shape[1]]).view(np.int32), (src, dst, o)
        st, w = resample.build_tables(dst, src, "spline64")
        p = resample.periodic_plan_up(st, w, dst, src)
        assert p is not None and p["taps"] in resample.PERIODIC_UP_TAPS and p["hi"] - p["lo"] >= dst // 4 - 4
        x = rng.uniform(-5, 260, dst).astype(np.float32)
        for o in list(range(4 * R * p["lo"], 4 * R * p["lo"] + 3 * R)) + list(range(4 * R * p["hi"] - 3 * R, 4 * R * p["hi"])):
            i, ph = divmod(o, R)
            xs = [x[i - p["offset"] + t] if 0 <= i - p["offset"] + t < dst else np.float32(0) for t in range(p["taps"])]
            got = fma_chain(p["w"][ph * p["taps"]:(ph + 1) * p["taps"]], xs)
            assert got.view(np.int32) == fma_chain(w[o], x[st[o]:st[o] + w.shape[1]]).view(np.int32), (dst, src, o)
    for src, dst in [(1920, 256), (1280, 384), (1000, 384)]:
        assert resample.periodic_plan_down(*resample.build_tables(src, dst, "spline64"), src, dst) is None
        assert resample.periodic_plan_up(*resample.build_tables(dst, src, "spline64"), dst, src) is None


def test_temporal_clip_adapter_feeds_halo_batches_and_scene_weights():
    """havc._TemporalClip (scope row N3, host side) with a recording stand-in engine: every request renders the aligned batch
    that holds the frame, the batch carries nh halo frames on either side clamped to the clip's ends (std.AverageFrames'
    min(max(n + d, 0), last)), the scene-change props of the window fold the weights per frame, frames come back under their
    own number with their own props and are cached."""
    import numpy as np
    from vsdeoldify_b200 import havc, vs_shim
    from vsdeoldify_b200.filters import scene_folded_weights, stab_weight_list
    n, H, W, B, nh = 11, 4, 6, 4, 2
    frames = np.arange(n * 3 * H * W, dtype=np.uint8).reshape(n, 3, H, W)
    props = [{"_SceneChangePrev": int(i in (0, 6)), "_SceneChangeNext": int(i == 5), "id": i} for i in range(n)]
    clip = vs_shim.array_clip(frames, props=props)
    calls = []

    class Engine:
        out_B = B

        def __init__(self):
            self.nh = nh
            self.temporal = type("T", (), {"wl": stab_weight_list(5, "A")})()

        def process_sequence(self, seq, n0, weights):
            calls.append((n0, seq[:, 0, 0, 0].copy(), None if weights is None else weights.copy()))
            return 255 - seq[nh:nh + B]
    fn = havc._TemporalClip(clip, Engine(), scene_weights=True)
    f = fn(9)                                                   # last, partial batch [8, 11): halo 6..7 before, clamped to 10 after
    assert f.props == props[9] and np.array_equal(np.stack([f[p] for p in range(3)]), 255 - frames[9])
    n0, first_px, w = calls[-1]
    assert n0 == 8 and list(first_px) == [frames[min(max(i, 0), n - 1), 0, 0, 0] for i in range(6, 14)]
    wl = stab_weight_list(5, "A")
    for b in range(B):
        idx = [min(max(8 + b + d, 0), n - 1) for d in range(-nh, nh + 1)]
        assert list(w[b]) == scene_folded_weights(wl, [props[i]["_SceneChangePrev"] for i in idx], [props[i]["_SceneChangeNext"] for i in idx])
    assert fn(10).props == props[10] and len(calls) == 1       # same batch: cached
    f = fn(1)                                                   # first batch: halo clamped to frame 0
    assert calls[-1][0] == 0 and list(calls[-1][1][:3]) == [frames[0, 0, 0, 0]] * 3
    # frame 5 ends a scene (_SceneChangeNext) and frame 6 starts one (_SceneChangePrev): the window of frame 5 keeps nothing beyond it
    fn(5)
    w5 = calls[-1][2][1]                                        # batch [4, 8): frame 5 is slot 1
    assert list(w5[3:]) == [0, 0] and sum(w5) == 100 and w5[2] == wl[2] + wl[3] + wl[4]
    assert scene_folded_weights([20] * 5, [0, 0, 0, 0, 0], [0, 0, 0, 0, 0]) == [20] * 5


def test_resize_min_hw_sizes_follow_the_reference_rules():
    """havc._min_hw_size (vsslib/vsresize.py:30-99): landscape clips above 480 lines go to height 480 with an even width (odd widths
    round down), portrait clips above 512 columns go to width 512 with an even height (odd heights round up), others stay."""
    from vsdeoldify_b200 import havc
    from oracle import pipeline_oracle
    cases = {(1920, 1080): (852, 480), (1280, 720): (852, 480), (720, 576): (600, 480), (640, 480): None, (644, 484): (638, 480),
             (1080, 1920): (512, 910), (600, 800): (512, 684), (500, 900): None, (3840, 2160): (852, 480)}
    for (w, h), want in cases.items():
        assert havc._min_hw_size(w, h) == want == pipeline_oracle.min_hw_size(w, h), (w, h)
