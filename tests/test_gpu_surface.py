"""GPU tests through the plugin surface (HAVC_colorizer / HAVC_main on clips): frame order under out-of-order
requests, bit-exact property pass-through, scene-change gating, and the stable/artistic S x S blend against the
CPU oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _register():
    from oracle import synth_weights
    from vsdeoldify_b200 import havc
    for name, (arch, seed) in {"ColorizeVideo_gen": ("wide", 1234), "ColorizeStable_gen": ("wide", 4321),
                               "ColorizeArtistic_gen": ("deep", 1234)}.items():
        if name not in havc._REGISTERED:        # another test module may have registered only some of them
            havc.register_state_dict(name, synth_weights.make_unet_state_dict(arch, seed))
    return havc


def _clip(n, h, w, seed=50):
    from oracle import synth_weights
    from vsdeoldify_b200 import vs_shim
    fr = np.stack([np.stack([synth_weights.make_test_frame(seed + 3 * i + c, h, w).numpy() for c in range(3)]) for i in range(n)])
    props = [{"_SceneChangePrev": int(i in (0, 3)), "_SceneChangeNext": int(i == 3), "sc_threshold": 0.1, "sc_frequency": 0,
              "sc_luma": 0.5, "sc_ratio": 1.0, "_Matrix": 1, "idx": i} for i in range(n)]
    return vs_shim.array_clip(fr, props=props), fr, props


def test_colorizer_clip_order_props_and_parity():
    from oracle import metrics, pipeline_oracle
    havc = _register()
    H, W, rf, n = 90, 160, 10, 5
    clip, fr, props = _clip(n, H, W)
    out = havc.HAVC_colorizer(clip, method=0, deoldify_p=[0, rf, 1.0, 0.0], ddcolor_p=[1, rf, 1.0, 0.0, True])
    assert (out.num_frames, out.width, out.height) == (n, W, H)
    order = [3, 0, 4, 2, 2, 1]                                        # out-of-order and repeated requests
    got = {i: out.get_frame(i) for i in order}
    sd = havc._REGISTERED["ColorizeVideo_gen"]
    for i in range(n):
        f = got[i]
        assert f.props == props[i], "frame properties must pass through bit-exactly"
        ref = pipeline_oracle.havc_colorizer_frame(sd, np.transpose(fr[i], (1, 2, 0)), rf)
        img = np.stack([np.asarray(f[p]) for p in range(3)], -1)
        m = metrics.frame_parity(img, ref)
        assert m["mean_de00"] <= 0.5, (i, m)                                 # also proves frame i is frame i
    # the source clip is untouched
    assert np.array_equal(np.asarray(clip.get_frame(3)[1]), fr[3, 1])


def test_scenechange_gating():
    """sc_min_freq > 0 -> scenechange=True: only n == 0 or _SceneChangePrev == 1 frames are colourised
    (vsslib/vsmodels.py:221-224); the others take the uncoloured squeeze/un-squeeze path."""
    from oracle import metrics, pipeline_oracle
    havc = _register()
    H, W, rf, n = 90, 160, 10, 5
    clip, fr, props = _clip(n, H, W, seed=90)
    out = havc.HAVC_colorizer(clip, method=0, deoldify_p=[0, rf, 1.0, 0.0], ddcolor_p=[1, rf, 1.0, 0.0, True], sc_min_freq=1)
    sd = havc._REGISTERED["ColorizeVideo_gen"]
    for i in range(n):
        colourised = i in (0, 3)
        ref = pipeline_oracle.havc_colorizer_frame(sd, np.transpose(fr[i], (1, 2, 0)), rf, skip=not colourised)
        f = out.get_frame(i)
        img = np.stack([np.asarray(f[p]) for p in range(3)], -1)
        d = np.abs(img.astype(int) - ref.astype(int))
        if colourised:
            assert metrics.frame_parity(img, ref)["mean_de00"] <= 0.5
        else:   # no network involved: integer pixel math + float resampling only
            assert d.max() <= 1 and (d > 0).mean() < 0.01, (i, int(d.max()), float((d > 0).mean()))
        assert f.props == props[i]


def test_havc_main_preset_path():
    """HAVC_main(Preset='VeryFast', ColorModel='DeOldify(Video)') == HAVC_colorizer(method=0, rf=16) followed by the preset's
    HAVC_stabilizer step (fast presets: colormap only, vsdeoldify/__init__.py:896-897)."""
    havc = _register()
    clip, fr, props = _clip(2, 144, 256, seed=170)
    a = havc.HAVC_main(clip, Preset="VeryFast", ColorModel="DeOldify(Video)", ColorMap="red->brown")
    b = havc.HAVC_stabilizer(havc.HAVC_deoldify(clip, model=0, render_factor=16, ddcolor_p=[1, 16, 1.0, 0.0, True]),
                             colormap="320:360|+50,0.90")
    # every preset built here runs with chroma_resize (vsdeoldify/__init__.py:492-494): resize_min_HW is the identity on a clip this
    # small, resize_to_chroma (restore_format, havc_utils.py:183-184) still takes the result's chroma through YUV420P8
    b = havc.resize_to_chroma(clip, b)
    for i in range(2):
        fa, fb = a.get_frame(i), b.get_frame(i)
        assert all(np.array_equal(np.asarray(fa[p]), np.asarray(fb[p])) for p in range(3))
        assert fa.props == props[i]


def test_read_ahead_sequential_equals_random_access():
    """Sequential readers are served through the read-ahead pipeline (two batches in flight), random readers batch by batch
    from wherever they land: same bytes, same props (needs a frame's result to be independent of its batch slot)."""
    havc = _register()
    n, H, W, rf = 21, 90, 160, 10
    clip, fr, props = _clip(n, H, W, seed=230)
    kw = dict(method=0, deoldify_p=[0, rf, 1.0, 0.0], ddcolor_p=[1, rf, 1.0, 0.0, True], sc_min_freq=1)
    a = havc.HAVC_colorizer(clip, **kw)
    seq = [a.get_frame(i) for i in range(n)]
    b = havc.HAVC_colorizer(clip, **kw)
    for i in [20, 3, 11, 0, 19, 8, 16, 7, 12, 1, 5, 14, 2, 18, 9, 4, 13, 6, 15, 10, 17]:
        f = b.get_frame(i)
        assert f.props == props[i] == seq[i].props
        for p in range(3):
            assert np.array_equal(np.asarray(f[p]), np.asarray(seq[i][p])), (i, p)
    # and a second sequential pass over a fresh graph reproduces the first one exactly
    c = havc.HAVC_colorizer(clip, **kw)
    for i in range(n):
        assert all(np.array_equal(np.asarray(c.get_frame(i)[p]), np.asarray(seq[i][p])) for p in range(3))


def _near(img, ref, what):
    """integer pixel math is exact; the two float resampling passes may differ in the last bit before rounding"""
    d = np.abs(img.astype(int) - ref.astype(int))
    assert d.max() <= 2 and (d > 0).mean() < 0.02, (what, int(d.max()), float((d > 0).mean()))


def test_stabilizer_clip_vs_oracle():
    """HAVC_stabilizer (per-frame stages) on a colour clip against the CPU restatement of the whole path; props pass through."""
    from oracle import pipeline_oracle
    from vsdeoldify_b200 import havc, vs_shim
    H, W, n = 120, 272, 3                        # rf 16 -> 256 x 256 working size (W >= 256)

    def colour_frame(seed, luma):                # smooth gradients + saturated 8x8 patches of every hue + a gray band
        rng = np.random.default_rng(seed)
        yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
        base = np.stack([0.5 + 0.5 * np.sin(xx / (9 + 3 * c) + yy / (17 - 2 * c) + c) for c in range(3)], -1)
        patch = np.kron(rng.uniform(0, 1, (H // 8 + 1, W // 8 + 1, 3)).astype(np.float32), np.ones((8, 8, 1), np.float32))[:H, :W]
        img = 0.6 * base + 0.4 * patch
        img[: H // 6] = img[: H // 6].mean(-1, keepdims=True)
        return (np.clip(img * (luma / img.mean()), 0, 1) * 255).astype(np.uint8)
    fr = np.stack([np.transpose(colour_frame(300 + i, luma), (2, 0, 1)) for i, luma in enumerate((0.1, 0.35, 0.7))])
    props = [{"_SceneChangePrev": int(i == 0), "idx": i} for i in range(n)]
    clip = vs_shim.array_clip(np.ascontiguousarray(fr), props=props)
    cases = [dict(dark=True, dark_p=[0.2, 0.8], smooth=True, smooth_p=[0.3, 0.7, 0.9, 0.0, "none"], colormap="red->brown"),
             dict(colormap="blue->green"), dict(smooth=True, smooth_p=[0.25, 0.6, 0.7, 0.1, "180:280|0.5,0.2"]), dict()]
    for kw in cases:
        out = havc.HAVC_stabilizer(clip, render_factor=16, **kw)
        okw = dict(kw)
        cm = okw.pop("colormap", "none")
        okw["colormap_adjust"] = havc._get_colormap(cm) if cm != "none" else "none"
        for i in (2, 0, 1):
            f = out.get_frame(i)
            assert f.props == props[i]
            img = np.stack([np.asarray(f[p]) for p in range(3)], -1)
            ref = pipeline_oracle.havc_stabilizer_frame(np.transpose(fr[i], (1, 2, 0)), render_factor=16, **okw)
            _near(img, ref, (kw, i))


@pytest.mark.parametrize("model,name", [(1, "ColorizeStable_gen"), (2, "ColorizeArtistic_gen")])
def test_stable_and_artistic_blend(model, name):
    from oracle import metrics, pipeline_oracle
    havc = _register()
    H, W, rf, n = 96, 160, 10, 2
    clip, fr, props = _clip(n, H, W, seed=130)
    out = havc.HAVC_colorizer(clip, method=0, deoldify_p=[model, rf, 1.0, 0.0], ddcolor_p=[1, rf, 1.0, 0.0, True])
    for i in range(n):
        ref = pipeline_oracle.havc_colorizer_frame(havc._REGISTERED["ColorizeVideo_gen"], np.transpose(fr[i], (1, 2, 0)), rf,
                                                   sd_other=havc._REGISTERED[name], video_weight=0.5)
        f = out.get_frame(i)
        img = np.stack([np.asarray(f[p]) for p in range(3)], -1)
        m = metrics.frame_parity(img, ref)
        assert m["mean_de00"] <= 0.5, (model, i, m)


def _register_zhang(havc):
    from oracle import zhang_oracle
    for name, file in havc._ZHANG_FILES.items():
        if file not in havc._REGISTERED:
            havc.register_state_dict(file, zhang_oracle.make_zhang_state_dict(name, 1234))


@pytest.mark.parametrize("method", [1, 2, 3, 4, 5, 6, 7])
def test_colorizer_two_models_merge_methods(method):
    """HAVC_colorizer with the Zhang siggraph17 model as second colour model (ddcolor_p model 2) and every merge
    method of vs_sc_combine_models, default ddtweak_p hue adjustment included, against the CPU oracle of the whole
    path (DeOldify + Zhang + hue adjust + merge + Spline64 back + luma transplant)."""
    from oracle import metrics, pipeline_oracle, zhang_oracle
    havc = _register()
    _register_zhang(havc)
    H, W, rf, n = 96, 176, 10, 2
    clip, fr, props = _clip(n, H, W, seed=130)
    out = havc.HAVC_colorizer(clip, method=method, mweight=0.45, deoldify_p=[0, rf, 1.0, 0.0], ddcolor_p=[2, rf, 1.0, 0.0, True])
    sd = havc._REGISTERED["ColorizeVideo_gen"]
    sdz = havc._REGISTERED[havc._ZHANG_FILES["siggraph17"]]
    for i in (1, 0):
        f = out.get_frame(i)
        assert f.props == props[i]
        ref = pipeline_oracle.havc_colorizer_frame(sd, np.transpose(fr[i], (1, 2, 0)), rf, zhang=("siggraph17", sdz), method=method,
                                                   merge_weight=0.45, hue_adjust="300:360|0.8,0.1")
        img = np.stack([np.asarray(f[p]) for p in range(3)], -1)
        m = metrics.frame_parity(img, ref)
        assert m["mean_de00"] <= 0.5, (method, i, m)


def test_colorizer_second_model_errors():
    from vsdeoldify_b200 import vs_shim
    havc = _register()
    clip, _, _ = _clip(2, 64, 96)
    with pytest.raises(vs_shim.Error):       # DDColor itself is out of scope
        havc.HAVC_colorizer(clip, method=2, deoldify_p=[0, 4, 1.0, 0.0], ddcolor_p=[1, 10, 1.0, 0.0, True])
    with pytest.raises(vs_shim.Error):
        havc.HAVC_colorizer(clip, method=9, deoldify_p=[0, 4, 1.0, 0.0], ddcolor_p=[2, 10, 1.0, 0.0, True])


@pytest.mark.parametrize("model,rf", [("video", 4), ("stable", 4), ("artistic", 6)])
def test_model_image_render_vs_reference_golden(model, rf):
    """ModelImageRender.get_transformed_image (the reference's per-image entry point, BASELINE cfg1) against outputs of
    the REAL reference (tests/golden/model_image_render.npz): Pillow BILINEAR squeeze / un-squeeze on non-square
    images, identity on square ones, the 50/50 blend of the second generator."""
    import os
    from PIL import Image
    from oracle import metrics, synth_weights
    havc = _register()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "model_image_render.npz"))
    color = lambda seed, h, w: np.stack([synth_weights.make_test_frame(seed + c, h, w).numpy() for c in range(3)], -1)
    S = rf * 16
    inputs = {"color": color(7, 120, 160), "gray": np.repeat(synth_weights.make_test_frame(9, 90, 144).numpy()[..., None], 3, -1),
              "square": color(21, S, S)}
    r = havc.ModelImageRender(package_dir=None, modelname=model, render_factor=rf, video_weight=0.5)
    for name, img in inputs.items():
        got = np.asarray(r.get_transformed_image(Image.fromarray(img)))
        m = metrics.frame_parity(got, g[f"{model}_rf{rf}_{name}"])
        assert m["mean_de00"] <= 0.5, (model, name, m)
        assert m["n_err_gt2"] <= 0.02 * m["n_values"], (model, name, m)


@pytest.mark.parametrize("zmodel,zname,gate", [(2, "siggraph17", 0.5), (3, "eccv16", 0.5)])
def test_colorizer_ddtweak_inverted_merge(zmodel, zname, gate):
    """cfg4-style call: ddtweak=[True, False, False] with the default luma-constrained tweak (luma_adjusted_levels on the
    second model's input, vs_recover_clip_luma on its output), clips swapped (cmb_sw).  North-star mean gate for both."""
    from oracle import metrics, pipeline_oracle
    from vsdeoldify_b200.constants import DEF_TWEAK_p
    havc = _register()
    _register_zhang(havc)
    H, W, rf, n = 96, 176, 10, 2
    clip, fr, props = _clip(n, H, W, seed=170)
    fr[1] //= 3                                                     # a dark frame: the luma floor / gamma branch is taken
    out = havc.HAVC_colorizer(clip, method=5, mweight=0.5, deoldify_p=[0, rf, 1.0, 0.0], ddcolor_p=[zmodel, rf, 1.0, 0.0, True],
                              ddtweak=[True, False, False], cmb_sw=True)
    sd = havc._REGISTERED["ColorizeVideo_gen"]
    sdz = havc._REGISTERED[havc._ZHANG_FILES[zname]]
    t = DEF_TWEAK_p
    tw = dict(bright=t[0], cont=t[1], gamma=t[2], luma_min=t[4], gamma_luma_min=t[5], gamma_alpha=t[6], gamma_min=t[7])
    for i in range(n):
        f = out.get_frame(i)
        assert f.props == props[i]
        ref = pipeline_oracle.havc_colorizer_frame(sd, np.transpose(fr[i], (1, 2, 0)), rf, zhang=(zname, sdz), method=5,
                                                   merge_weight=0.5, hue_adjust="300:360|0.8,0.1", invert=True, ddtweak=tw)
        img = np.stack([np.asarray(f[p]) for p in range(3)], -1)
        m = metrics.frame_parity(img, ref)
        assert m["mean_de00"] <= gate, (zname, i, m)


def test_multi_gpu_sharded_clip_matches_single_gpu():
    """HAVC_colorizer(device_index=[0, 1]) (vsdeoldify_b200/sharded.py): one engine + worker per GPU, the clip's frames
    partitioned over them, delivered in order with their props; bytes equal the single-GPU clip for every frame, for the block
    and the interleaved partition and for out-of-order requests.  Needs two visible GPUs."""
    import os
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    havc = _register()
    H, W, rf, n = 96, 128, 4, 23
    old_batch = havc._BATCH
    havc._BATCH = 4
    try:
        clip, fr, props = _clip(n, H, W, seed=310)
        single = havc.HAVC_colorizer(clip, method=0, deoldify_p=[0, rf, 1.0, 0.0])
        want = [np.stack([np.asarray(single.get_frame(i)[p]) for p in range(3)]) for i in range(n)]
        for part in ("interleaved", "block"):
            os.environ["HAVC_B200_PARTITION"] = part
            multi = havc.HAVC_colorizer(clip, method=0, deoldify_p=[0, rf, 1.0, 0.0], device_index=[0, 1])
            for i in list(range(n)) + [5, 22, 0, 13]:
                f = multi.get_frame(i)
                assert f.props == props[i], (part, i)
                assert np.array_equal(np.stack([np.asarray(f[p]) for p in range(3)]), want[i]), (part, i)
    finally:
        havc._BATCH = old_batch
        os.environ.pop("HAVC_B200_PARTITION", None)


def test_havc_merge_with_clip_luma():
    """HAVC_merge(clipa, clipb, clip_luma) (vsdeoldify/__init__.py:2633-2675): methods 3..7 squeeze both clips to
    frame_size x frame_size (from 0.4 * clip_luma.width), merge there and go through _clip_chroma_resize(clip_luma, .);
    methods 0 / 1 are _clip_chroma_resize of one clip.  Output frames are clip_luma's frames (props included) with the
    scene-detection props of the merged clip (CopySCDetect)."""
    from oracle import filters_oracle as fo, metrics, pixel_oracle as px, synth_weights
    from vsdeoldify_b200 import havc, vs_shim
    n, (Ha, Wa), (H, W) = 3, (120, 176), (288, 704)
    color = lambda seed, h, w: np.stack([synth_weights.make_test_frame(seed + c, h, w).numpy() for c in range(3)])
    fa = np.stack([color(600 + 5 * i, Ha, Wa) for i in range(n)])
    fb = np.stack([color(700 + 5 * i, Ha, Wa) for i in range(n)])
    fl = np.stack([color(800 + 5 * i, H, W) for i in range(n)])
    pa = [{"_SceneChangePrev": int(i == 1), "who": "a"} for i in range(n)]
    pl = [{"who": "luma", "idx": i} for i in range(n)]
    ca, cb, cl = vs_shim.array_clip(fa, props=pa), vs_shim.array_clip(fb), vs_shim.array_clip(fl, props=pl)
    hwc = lambda x: np.ascontiguousarray(np.transpose(x, (1, 2, 0)))
    fs = min(min(max(int(0.4 * W / 16), 16), 32) * 16, W)
    assert fs == 272
    for method, weight in ((3, 0.5), (5, 0.4), (0, 0.5), (1, 0.5)):
        out = havc.HAVC_merge(ca, cb, clip_luma=cl, weight=weight, method=method)
        assert (out.width, out.height) == (W, H)
        for i in range(n):
            f = out.get_frame(i)
            assert f.props["who"] == "luma" and f.props["idx"] == i
            if method != 1:                                       # CopySCDetect from the merged clip (= clipa's props)
                assert f.props.get("_SceneChangePrev") == pa[i]["_SceneChangePrev"]
            if method in (0, 1):
                low = hwc((fa if method == 0 else fb)[i])
            else:
                low = fo.combine_models(px.resize_plane_u8(hwc(fa[i]), fs, fs), px.resize_plane_u8(hwc(fb[i]), fs, fs), method, weight)
            want = px.chroma_post_process(px.resize_plane_u8(low, W, H), hwc(fl[i]))
            got = np.stack([np.asarray(f[p]) for p in range(3)], -1)
            m = metrics.frame_parity(got, want)
            assert m["mean_de00"] < 0.02 and m["n_err_gt2"] <= 2e-4 * m["n_values"], (method, i, m)


@pytest.mark.parametrize("method", [0, 2])
def test_colorizer_sat_hue_tweak(method):
    """deoldify_p / ddcolor_p saturation and hue (vs_tweak inside vs_sc_combine_models, mcomb.py:154-169) against the oracle of the
    whole path with the restated zimg round trip."""
    from oracle import metrics, pipeline_oracle
    havc = _register()
    _register_zhang(havc)
    H, W, rf, n = 96, 176, 10, 2
    clip, fr, props = _clip(n, H, W, seed=190)
    out = havc.HAVC_colorizer(clip, method=method, mweight=0.4, deoldify_p=[0, rf, 0.85, 8.0], ddcolor_p=[2, rf, 1.2, -5.0, True])
    sd = havc._REGISTERED["ColorizeVideo_gen"]
    sdz = havc._REGISTERED[havc._ZHANG_FILES["siggraph17"]]
    for i in range(n):
        f = out.get_frame(i)
        assert f.props == props[i]
        ref = pipeline_oracle.havc_colorizer_frame(sd, np.transpose(fr[i], (1, 2, 0)), rf, zhang=("siggraph17", sdz) if method else None,
                                                   method=method, merge_weight=0.4, hue_adjust="300:360|0.8,0.1", sat=(0.85, 1.2),
                                                   hue=(8.0, -5.0))
        img = np.stack([np.asarray(f[p]) for p in range(3)], -1)
        m = metrics.frame_parity(img, ref)
        assert m["mean_de00"] <= 0.5, (method, i, m)


def _yuv_clip(n, H, W, seed, fmt, props_extra=None):
    """A YUV420P8 (limited range) or GRAY8 clip of the stand-in VapourSynth with its plane arrays."""
    from oracle import synth_weights
    from vsdeoldify_b200 import vs_shim
    frames = []
    for i in range(n):
        y = (16 + synth_weights.make_test_frame(seed + 3 * i, H, W).numpy().astype(np.float32) * (219.0 / 255.0)).astype(np.uint8)
        if fmt == "gray8":
            frames.append([y])
        else:
            u = (128 + (synth_weights.make_test_frame(seed + 3 * i + 1, H // 2, W // 2).numpy().astype(np.int32) - 128) // 4).astype(np.uint8)
            v = (128 + (synth_weights.make_test_frame(seed + 3 * i + 2, H // 2, W // 2).numpy().astype(np.int32) - 128) // 5).astype(np.uint8)
            frames.append([y, u, v])
    vfmt = vs_shim.GRAY8 if fmt == "gray8" else vs_shim.YUV420P8
    props = [dict({"_SceneChangePrev": int(i == 0), "idx": i}, **(props_extra or {})) for i in range(n)]
    fn = lambda i: vs_shim.VideoFrame(frames[i], vfmt, props[i])
    return vs_shim.VideoNode(n, W, H, vfmt, fn), frames, props


@pytest.mark.parametrize("fmt,extra,matrix,out_limited", [("yuv420p8", {"_Matrix": 1}, "709", True), ("yuv420p8", {"_Matrix": 6, "_ColorRange": 0}, "601", False),
                                                          ("gray8", {}, "709", True)])
def test_colorizer_on_yuv420p8_and_gray8_clips(fmt, extra, matrix, out_limited):
    """convert_format_RGB24 / restore_format (havc_utils.py:57-237) on the device: an 8-bit YUV 4:2:0 or GRAY clip goes in, a
    YUV420P8 clip comes out (same format / props for YUV; a GRAY clip comes back as BT.709 YUV420P8), planes compared with the
    oracle of the whole path (restated zimg conversions around the RGB24 pipeline)."""
    from oracle import metrics, pipeline_oracle, zimg_oracle as zo
    from vsdeoldify_b200 import vs_shim
    havc = _register()
    H, W, rf, n = 96, 160, 10, 3
    clip, frames, props = _yuv_clip(n, H, W, 1200, fmt, extra)
    out = havc.HAVC_colorizer(clip, method=0, deoldify_p=[0, rf, 1.0, 0.0], ddcolor_p=[1, rf, 1.0, 0.0, True])
    assert out.format == vs_shim.YUV420P8 and (out.width, out.height) == (W, H)
    sd = havc._REGISTERED["ColorizeVideo_gen"]
    out_matrix = matrix if fmt == "yuv420p8" else "709"
    for i in (2, 0, 1):
        f = out.get_frame(i)
        assert f.props == props[i] and f.format == vs_shim.YUV420P8
        want = pipeline_oracle.havc_colorizer_yuv_frame(sd, frames[i], rf, matrix=matrix, out_limited=out_limited)
        got = [np.asarray(f[p]) for p in range(3)]
        assert all(g.shape == w.shape for g, w in zip(got, want))
        assert np.abs(got[0].astype(int) - want[0].astype(int)).max() <= 2                 # luma: transplanted, dither ties only
        rgb_g = zo.yuv420p8_to_rgb24(*got, dither=False, matrix=out_matrix, limited=out_limited)
        rgb_w = zo.yuv420p8_to_rgb24(*want, dither=False, matrix=out_matrix, limited=out_limited)
        m = metrics.frame_parity(rgb_g, rgb_w)
        assert m["mean_de00"] <= 0.5, (fmt, i, m)


def test_zimg_conversions_bit_exact_vs_restatement():
    """RGB24 <-> YUV420P8 through the C ABI for both matrices / ranges with error-diffusion dither in both directions, and GRAY8 ->
    RGB24, against oracle/zimg_oracle.py: bit-exact."""
    import torch
    from oracle import synth_weights, zimg_oracle as zo
    from vsdeoldify_b200.engine import _FormatIO
    B, H, W = 2, 48, 80
    imgs = [np.stack([synth_weights.make_test_frame(1300 + 5 * i + c, H, W).numpy() for c in range(3)], -1) for i in range(B)]
    rgb = torch.from_numpy(np.ascontiguousarray(np.stack([np.transpose(x, (2, 0, 1)) for x in imgs]))).cuda()
    for matrix in ("709", "601"):
        for limited in (True, False):
            io = _FormatIO("yuv420p8", B, H, W, torch.device("cuda:0"), matrix, limited)
            raw = torch.empty(io.out_bytes, dtype=torch.uint8, device="cuda")
            io.from_rgb(rgb, raw, 0)
            torch.cuda.synchronize()
            arr = raw.cpu().numpy()
            for j in range(B):
                want = zo.rgb24_to_yuv420p8(imgs[j], matrix=matrix, limited=limited, dither=True)
                for g, w in zip(io.planes(arr, j, out=True), want):
                    assert np.array_equal(g, w), (matrix, limited, j, int((g != w).sum()))
            if limited:                                    # the way in always reads limited-range YUV (havc_utils.py:139)
                back = torch.empty_like(rgb)
                io.to_rgb(raw, back, 0)
                torch.cuda.synchronize()
                for j in range(B):
                    y, u, v = io.planes(arr, j, out=True)
                    want = zo.yuv420p8_to_rgb24(y, u, v, dither=True, matrix=matrix, limited=True)
                    assert np.array_equal(np.transpose(back[j].cpu().numpy(), (1, 2, 0)), want), (matrix, j)
    io = _FormatIO("gray8", B, H, W, torch.device("cuda:0"))
    g = torch.from_numpy(np.stack([x[..., 0] for x in imgs]).reshape(-1)).cuda()
    back = torch.empty_like(rgb)
    io.to_rgb(g, back, 0)
    torch.cuda.synchronize()
    for j in range(B):
        assert np.array_equal(np.transpose(back[j].cpu().numpy(), (1, 2, 0)), zo.gray8_to_rgb24(imgs[j][..., 0]))


GT = np.load(os.path.join(os.path.dirname(__file__), "golden", "vsslib_temporal.npz"))


@pytest.mark.parametrize("name,kw", [
    ("stab_default", dict(nframes=5, mode="A", sat=1.0, tht=15, weight=0.2, tht_scen=0.8)),
    ("stab_w3", dict(nframes=3, mode="W", sat=0.8, tht=30, weight=-0.4, tht_scen=0.8, hue_adjust="0:60|0.8,0.1")),
    ("stab_tht0", dict(nframes=7, mode="A", tht=0))])
def test_temporal_stabilizer_bit_exact_vs_reference_graph(name, kw):
    """Scope row N3: vs_chroma_stabilizer_ex on the device against the golden of the REAL reference graph
    (tests/golden/make_golden.py::golden_temporal; zimg / std.AverageFrames underneath restated: unpinned).  Every byte equal:
    halo frames at the clip's ends and across batch boundaries, the n < 15 pass-through, the luma gate, the scene-change
    folding of the weights, props untouched, any request order."""
    from vsdeoldify_b200 import havc, vs_shim
    clip_np = GT["clip"]
    T = clip_np.shape[0]
    props = [{"_SceneChangePrev": int(n in (0, 9)), "_SceneChangeNext": int(n in (8, T - 1)), "idx": n} for n in range(T)]
    clip = vs_shim.array_clip(np.ascontiguousarray(np.transpose(clip_np, (0, 3, 1, 2))), props=props)
    out = havc.vs_chroma_stabilizer_ex(clip, **kw)
    frames = [int(n) for n in GT["frames"]]
    for n in frames[::-1]:
        f = out.get_frame(n)
        assert f.props == props[n]
        got = np.dstack([np.asarray(f[p]) for p in range(3)])
        want = GT[f"{name}_{n}"]
        assert np.array_equal(got, want), (name, n, int((got != want).sum()), int(np.abs(got.astype(int) - want).max()))


def test_havc_stabilizer_temporal_chain_with_host_plugin():
    """HAVC_stabilizer(stab=True): squeeze -> per-frame stages -> temporal stabiliser -> (host-provided ReduceFlicker) ->
    _clip_chroma_resize, against the CPU restatement of the chain.  The plugin is external (vsplugins.py:263-272): the test
    registers a pass-through `core.rdfl` on the stand-in and checks that it is what the chain ends in."""
    from oracle import pipeline_oracle
    from vsdeoldify_b200 import havc, vs_shim
    import types
    fr = np.ascontiguousarray(np.kron(GT["clip"], np.ones((1, 2, 2, 1), np.uint8)))       # the golden clip, 2x: 20 x 48 x 64 x 3
    T, H, W = fr.shape[:3]
    props = [{"_SceneChangePrev": int(n == 0), "idx": n} for n in range(T)]
    clip = vs_shim.array_clip(np.ascontiguousarray(np.transpose(fr, (0, 3, 1, 2))), props=props)
    with pytest.raises(vs_shim.Error, match="ReduceFlicker"):
        havc.HAVC_stabilizer(clip, stab=True, render_factor=16)
    calls = []

    def reduce_flicker(clip=None, strength=None, aggressive=None):
        calls.append((strength, aggressive))
        return clip
    vs_shim.core.rdfl = types.SimpleNamespace(ReduceFlicker=reduce_flicker)
    try:
        kw = dict(dark=True, dark_p=[0.2, 0.8], smooth=True, smooth_p=[0.3, 0.7, 0.9, 0.0, "none"])
        out = havc.HAVC_stabilizer(clip, stab=True, stab_p=[5, 'A', 1, 15, 0.2, 0.8], render_factor=16, **kw)
        assert calls == [(2, 0)]                                                      # vs_reduce_flicker's defaults
        only = (16, 3, 19)
        ref = pipeline_oracle.havc_stabilizer_clip(fr, only, render_factor=16, **kw)
        for n in only:
            f = out.get_frame(n)
            assert f.props == props[n]
            img = np.stack([np.asarray(f[p]) for p in range(3)], -1)
            d = np.abs(img.astype(int) - ref[n].astype(int))
            # the Spline64 squeeze may differ in the last bit before rounding; error diffusion then places its +-1 decisions
            # elsewhere: a tolerance on the distribution, the temporal stage itself is bit-exact (test above)
            assert d.mean() < 0.35 and np.percentile(d, 99) <= 2, (n, float(d.mean()), int(d.max()))
            assert not np.array_equal(img, fr[n])
    finally:
        del vs_shim.core.rdfl


def test_merge_and_temporal_stabilizer_on_yuv_clips():
    """The format glue of the other entry points under the stand-in (convert_format_RGB24 / restore_format on the device,
    engine.FormatEngine): HAVC_merge converts clipa / clipb and restores clipa's format (vsdeoldify/__init__.py:2651-2675);
    vs_chroma_stabilizer_ex / HAVC_stabilizer accept a YUV420P8 clip.  Bit-exact against the restated conversions around the
    RGB24 oracles (the conversions and the merges are integer / fixed-order float32 arithmetic on both sides)."""
    from oracle import filters_oracle as fo, temporal_oracle as to, zimg_oracle as zo
    from vsdeoldify_b200 import havc, vs_shim
    H, W, n = 48, 64, 18
    ca, fa, pa = _yuv_clip(n, H, W, 2100, "yuv420p8", {"_Matrix": 6, "_ColorRange": 0})
    cb, fb, _ = _yuv_clip(n, H, W, 2200, "yuv420p8", {"_Matrix": 1})
    rgb_a = [zo.yuv420p8_to_rgb24(*f, dither=True, matrix="601", limited=True) for f in fa]     # the input range is read as limited (:133-143)
    rgb_b = [zo.yuv420p8_to_rgb24(*f, dither=True, matrix="709", limited=True) for f in fb]
    out = havc.HAVC_merge(ca, cb, weight=0.4, method=3)
    assert out.format == vs_shim.YUV420P8
    for i in (5, 0):
        f = out.get_frame(i)
        assert f.props == pa[i]
        want = zo.rgb24_to_yuv420p8(fo.combine_models(rgb_a[i], rgb_b[i], 3, 0.4), "601", False, True)    # clipa's matrix / range
        for p in range(3):
            assert np.array_equal(np.asarray(f[p]), want[p]), (i, p, int((np.asarray(f[p]) != want[p]).sum()))
    st = havc.vs_chroma_stabilizer_ex(ca, nframes=3, mode="A", sat=1.0, tht=20, weight=0.2, tht_scen=0.8)
    assert st.format == vs_shim.YUV420P8
    ref = to.chroma_stabilizer_ex(np.stack(rgb_a), nframes=3, mode="A", sat=1.0, tht=20, weight=0.2, tht_scen=0.8, only=[16, 3])
    for i in (16, 3):
        f = st.get_frame(i)
        assert f.props == pa[i]
        want = zo.rgb24_to_yuv420p8(ref[i], "601", False, True)
        for p in range(3):
            assert np.array_equal(np.asarray(f[p]), want[p]), (i, p, int((np.asarray(f[p]) != want[p]).sum()))
    g, fg, pg = _yuv_clip(3, H, 80, 2300, "gray8")
    sg = havc.HAVC_stabilizer(g, dark=True, render_factor=16)                                      # GRAY8 in -> YUV420P8 out (:208-222)
    assert sg.format == vs_shim.YUV420P8 and sg.get_frame(1).props == pg[1] and np.asarray(sg.get_frame(1)[1]).shape == (H // 2, 40)


def test_resize_min_hw_and_resize_to_chroma_vs_oracle():
    """HAVC_main's chroma_resize detour (vsslib/vsresize.py:30-127): Spline36 reduction to height 480 and the way back with the luma
    of the full-size clip, against the CPU restatement; and HAVC_main == the same steps chained by hand (one frame, 644 x 484)."""
    from oracle import pipeline_oracle
    havc = _register()
    H, W = 484, 644
    clip, fr, props = _clip(2, H, W, seed=910)
    small = havc.resize_min_HW(clip)
    assert (small.width, small.height) == pipeline_oracle.min_hw_size(W, H) == (638, 480)
    src = np.ascontiguousarray(np.transpose(fr[1], (1, 2, 0)))
    got = np.stack([np.asarray(small.get_frame(1)[p]) for p in range(3)], -1)
    want = pipeline_oracle.resize_min_hw(src)
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.01, (int(d.max()), float((d > 0).mean()))   # float passes: last-bit ties only
    assert small.get_frame(1).props == props[1]
    # the way back: the low-resolution frame of the oracle goes in on both sides, so only the float resize can differ
    from vsdeoldify_b200 import vs_shim
    low_clip = vs_shim.array_clip(np.ascontiguousarray(np.transpose(want, (2, 0, 1)))[None].repeat(2, 0),
                                  props=[{"_SceneChangePrev": 1, "sc_threshold": 0.25}, {"_SceneChangePrev": 0, "sc_threshold": 0.25}])
    back = havc.resize_to_chroma(clip, low_clip)
    f = back.get_frame(1)
    assert f.props == dict(props[1], _SceneChangePrev=0, sc_threshold=0.25)               # CopyFrameProps of the SC props (:124-125)
    got = np.stack([np.asarray(f[p]) for p in range(3)], -1)
    want_back = pipeline_oracle.resize_to_chroma(src, want)
    d = np.abs(got.astype(int) - want_back.astype(int))
    assert d.mean() < 0.05 and np.percentile(d, 99.9) <= 2, (float(d.mean()), int(d.max()))   # error diffusion moves +-1 decisions
    a = havc.HAVC_main(clip, Preset="VeryFast", ColorModel="DeOldify(Video)")
    b = havc.resize_to_chroma(clip, havc.HAVC_stabilizer(havc.HAVC_deoldify(small, model=0, render_factor=16, ddcolor_p=[1, 16, 1.0, 0.0, True])))
    fa, fb = a.get_frame(0), b.get_frame(0)
    assert (a.width, a.height) == (W, H) and all(np.array_equal(np.asarray(fa[p]), np.asarray(fb[p])) for p in range(3))
