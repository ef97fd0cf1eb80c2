#!/usr/bin/env python
"""Generates the golden fixtures in this directory by running the REAL reference (imported from
/root/reference through oracle/refshim.py) on seeded inputs and seeded synthetic weights.

Run in the build container only (the reference tree does not travel to the GPU box):
    python tests/golden/make_golden.py
The fixtures are small (.npz, < 1 MB total); weights are not stored — both sides regenerate them from the seed
with oracle/synth_weights.py (the generator's schema is checked here against the reference's own state_dict).
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim, synth_weights  # noqa: E402


def color_test_image(seed, h, w):
    return np.stack([synth_weights.make_test_frame(seed + c, h, w).numpy() for c in range(3)], -1)


def golden_unets():
    for arch in ("wide", "deep"):
        sd = synth_weights.make_unet_state_dict(arch, 1234)
        m = refshim.build_unet(arch)
        ref_sd = m.state_dict()
        assert list(ref_sd.keys()) == list(sd.keys()), "synthetic state-dict schema differs from the reference"
        for k in sd:
            assert ref_sd[k].shape == sd[k].shape and ref_sd[k].dtype == sd[k].dtype, k
        m.load_state_dict(sd, strict=True)
        x = synth_weights.calibration_batch(5, 64, 1)
        with torch.no_grad():
            y = m(x)
        np.savez_compressed(os.path.join(HERE, f"unet_{arch}_s64.npz"), y=y.numpy(), x_seed=5, size=64, weight_seed=1234)
        print(arch, "unet golden: out mean/std", float(y.mean()), float(y.std()))


def golden_render():
    """vsdeoldify.deoldify.visualize.ModelImageRender — the reference's own per-frame entry point
    (vsslib/vsmodels.py:196-233 constructs exactly this) — on CPU with synthetic weights."""
    refshim.install()
    import vsdeoldify.deoldify.generators as gen
    from vsdeoldify.fastai.vision.learner import create_body as _cb
    gen.create_body = lambda arch, pretrained=True, cut=None: _cb(arch, False, cut)   # no network
    from vsdeoldify.deoldify import device
    from vsdeoldify.deoldify.device_id import DeviceId
    device.set(device=DeviceId.CPU)
    from PIL import Image
    from vsdeoldify.deoldify.visualize import ModelImageRender
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "models"))
        torch.save(synth_weights.make_unet_state_dict("wide", 1234), os.path.join(tmp, "models", "ColorizeVideo_gen.pth"))
        torch.save(synth_weights.make_unet_state_dict("wide", 4321), os.path.join(tmp, "models", "ColorizeStable_gen.pth"))
        torch.save(synth_weights.make_unet_state_dict("deep", 1234), os.path.join(tmp, "models", "ColorizeArtistic_gen.pth"))
        cwd = os.getcwd()
        os.chdir(tmp)   # get_dummy_databunch() reads ./dummy/
        try:
            out = {}
            img = color_test_image(7, 120, 160)
            gray = np.repeat(synth_weights.make_test_frame(9, 90, 144).numpy()[..., None], 3, -1)
            for model, rf in (("video", 4), ("stable", 4), ("artistic", 6)):
                r = ModelImageRender(package_dir=tmp, modelname=model, render_factor=rf, video_weight=0.5)
                out[f"{model}_rf{rf}_color"] = np.asarray(r.get_transformed_image(Image.fromarray(img)))
                out[f"{model}_rf{rf}_gray"] = np.asarray(r.get_transformed_image(Image.fromarray(gray)))
                # square input of exactly S x S: the resize is the identity (the HAVC_colorizer situation)
                S = rf * 16
                sq = color_test_image(21, S, S)
                out[f"{model}_rf{rf}_square"] = np.asarray(r.get_transformed_image(Image.fromarray(sq)))
                print(model, rf, "done")
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "model_image_render.npz"), **out)


def golden_render_narrow():
    """ModelImageRender on SQUARE images whose side F differs from render_factor*16 - what vs_sc_deoldify hands it when the clip
    is narrower than render_factor*16 or when ddcolor_rf > deoldify_rf (frame_size, vsdeoldify/__init__.py:2502): the filter
    stretches F -> rf*16 with Pillow BILINEAR, renders, and resizes back (deoldify/filters.py:37-41,70-73)."""
    refshim.install()
    import vsdeoldify.deoldify.generators as gen
    from vsdeoldify.fastai.vision.learner import create_body as _cb
    gen.create_body = lambda arch, pretrained=True, cut=None: _cb(arch, False, cut)   # no network
    from vsdeoldify.deoldify import device
    from vsdeoldify.deoldify.device_id import DeviceId
    device.set(device=DeviceId.CPU)
    from PIL import Image
    from vsdeoldify.deoldify.visualize import ModelImageRender
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "models"))
        torch.save(synth_weights.make_unet_state_dict("wide", 1234), os.path.join(tmp, "models", "ColorizeVideo_gen.pth"))
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            out = {}
            r = ModelImageRender(package_dir=tmp, modelname="video", render_factor=4, video_weight=0.5)
            for F in (48, 80):
                sq = color_test_image(23 + F, F, F)
                out[f"video_rf4_sq{F}"] = np.asarray(r.get_transformed_image(Image.fromarray(sq)))
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "model_image_render_narrow.npz"), **out)


def golden_cfg1():
    """BASELINE.json configs[0]: the reference's ModelImageRender('video', render_factor=24) on the 23 test_images/*.jpg stills
    (1090 x 767), torch CPU, synthetic weights seed 1234.  The fixture holds the original JPEG bytes (the GPU box has no
    /root/reference) and the reference's output sampled at every third pixel in both directions (1/9 of the values: the full
    outputs would be 57 MB): tests/golden/cfg1_stills.npz."""
    refshim.install()
    import glob
    import io
    import vsdeoldify.deoldify.generators as gen
    from vsdeoldify.fastai.vision.learner import create_body as _cb
    gen.create_body = lambda arch, pretrained=True, cut=None: _cb(arch, False, cut)   # no network
    from vsdeoldify.deoldify import device
    from vsdeoldify.deoldify.device_id import DeviceId
    device.set(device=DeviceId.CPU)
    from PIL import Image
    from vsdeoldify.deoldify.visualize import ModelImageRender
    files = sorted(glob.glob("/root/reference/test_images/Image_*_test.jpg"))
    assert len(files) == 23
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "models"))
        torch.save(synth_weights.make_unet_state_dict("wide", 1234), os.path.join(tmp, "models", "ColorizeVideo_gen.pth"))
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            r = ModelImageRender(package_dir=tmp, modelname="video", render_factor=24, video_weight=0.5)
            for k, f in enumerate(files):
                raw = open(f, "rb").read()
                img = Image.open(io.BytesIO(raw)).convert("RGB")
                res = np.asarray(r.get_transformed_image(img))
                out[f"jpeg_{k:02d}"] = np.frombuffer(raw, np.uint8)
                out[f"ref_{k:02d}"] = np.ascontiguousarray(res[::3, ::3])
                print(os.path.basename(f), res.shape, "done")
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "cfg1_stills.npz"), **out)


def golden_pixels():
    """Reference vsslib helpers that import without VapourSynth (imfilters/nputils)."""
    refshim.install()
    import types
    if "vapoursynth" not in sys.modules:
        sys.modules["vapoursynth"] = types.ModuleType("vapoursynth")
    from PIL import Image
    from vsdeoldify.vsslib import imfilters
    a, b = color_test_image(31, 72, 96), color_test_image(41, 72, 96)
    out = {"a": a, "b": b}
    out["chroma_post_process"] = np.asarray(imfilters.chroma_post_process(Image.fromarray(a), Image.fromarray(b)))
    for w in (0.15, 0.4, 0.5, 0.6):
        out[f"weighted_merge_{w}"] = np.asarray(imfilters.image_weighted_merge(Image.fromarray(a), Image.fromarray(b), w))
    np.savez_compressed(os.path.join(HERE, "vsslib_pixels.npz"), **out)


def filter_test_pair(seed, h, w, luma):
    """Two colourful uint8 RGB images with frame-mean luma near `luma` (0..1): `a` plays the DeOldify result,
    `b` the second model; saturated patches of every hue, gray areas and hard edges included."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(2):
        base = color_test_image(seed * 10 + k * 3, h, w).astype(np.float32) / 255
        hue_patch = rng.uniform(0, 1, (h // 8 + 1, w // 8 + 1, 3)).astype(np.float32)
        hue_patch = np.kron(hue_patch, np.ones((8, 8, 1), np.float32))[:h, :w]
        img = 0.6 * base + 0.4 * hue_patch
        img[: h // 6] = img[: h // 6].mean(-1, keepdims=True)            # a gray band (low saturation)
        img = img * (luma / max(img.mean(), 1e-3))
        out.append((np.clip(img, 0, 1) * 255).astype(np.uint8))
    return out


def golden_filters():
    """vsslib merges and chroma-adjust filters, run through the REAL reference code: mcomb.vs_sc_combine_models on
    clips of the in-repo VapourSynth stand-in (the reference's selectors, frame_to_image / image_to_frame and
    imfilters / restcolor / nputils execute unmodified), plus the image-level helpers called directly."""
    refshim.install()
    from vsdeoldify_b200 import vs_shim
    sys.modules["vapoursynth"] = vs_shim
    from PIL import Image
    from vsdeoldify.vsslib import imfilters, mcomb, restcolor, vsfilters
    out = {}
    H, W = 64, 80
    lumas = (0.05, 0.15, 0.25, 0.5, 0.85)
    for li, luma in enumerate(lumas):
        a, b = filter_test_pair(50 + li, H, W, luma)
        out[f"a_{li}"], out[f"b_{li}"] = a, b
        clip_a = vs_shim.array_clip(np.ascontiguousarray(np.transpose(a, (2, 0, 1)))[None])
        clip_b = vs_shim.array_clip(np.ascontiguousarray(np.transpose(b, (2, 0, 1)))[None])
        for method in (2, 3, 4, 5, 7):
            for wi, w in enumerate((0.4, 0.7)):
                c = mcomb.vs_sc_combine_models(clip_a, clip_b, method=method, clipb_weight=w, scenechange=False)
                f = c.get_frame(0)
                out[f"combine_m{method}_w{wi}_{li}"] = np.dstack([np.asarray(f[p]) for p in range(3)])
        # non-default parameters: hard luma mask (limit == white), no red fix, adaptive alpha 2
        c = mcomb.vs_sc_combine_models(clip_a, clip_b, method=4, clipb_weight=0.6, LMM_p=[0.3, 0.3, 1.0], scenechange=False)
        out[f"combine_m4_hard_{li}"] = np.dstack([np.asarray(c.get_frame(0)[p]) for p in range(3)])
        c = mcomb.vs_sc_combine_models(clip_a, clip_b, method=3, clipb_weight=0.5, CMC_p=[0.3, False, 20, 24], scenechange=False)
        out[f"combine_m3_noredfix_{li}"] = np.dstack([np.asarray(c.get_frame(0)[p]) for p in range(3)])
        c = mcomb.vs_sc_combine_models(clip_a, clip_b, method=5, clipb_weight=0.5, ALM_p=[0.6, 2.0, 0.1], scenechange=False)
        out[f"combine_m5_alpha2_{li}"] = np.dstack([np.asarray(c.get_frame(0)[p]) for p in range(3)])
        # method 6 up to (not including) std.Merge: vs_sc_recover_gradient_color on the clips
        for algo in (0, 1, 2):
            c = vsfilters.vs_sc_recover_gradient_color(clip=clip_a, clip_color=clip_b, sat=0.8, tht=30, weight=0, alpha=2.0,
                                                       scenechange=False, algo=algo)
            out[f"recover_gradient_algo{algo}_{li}"] = np.dstack([np.asarray(c.get_frame(0)[p]) for p in range(3)])
        out[f"luma_{li}"] = np.float64(imfilters.get_image_luma(Image.fromarray(a), 255))
        ia = Image.fromarray(b)
        out[f"hue_adjust_default_{li}"] = np.asarray(restcolor.adjust_hue_range(ia, hue_adjust="300:360|0.8,0.1"))
        out[f"hue_adjust_shift_{li}"] = np.asarray(restcolor.adjust_hue_range(ia, hue_adjust="blue,cyan|+40,0.3"))
        out[f"hue_adjust_neg_{li}"] = np.asarray(restcolor.adjust_hue_range(ia, hue_adjust="0:60,200:260|0.5,-0.4"))
        out[f"tweak_bcg_{li}"] = np.asarray(imfilters.image_tweak(ia, bright=12, cont=1.1))
        out[f"tweak_sat_range_{li}"] = np.asarray(imfilters.image_tweak(ia, sat=0.6, hue_range="280:360,0:30"))
        out[f"tweak_sat_up_{li}"] = np.asarray(imfilters.image_tweak(ia, sat=1.4, bright=-20))
        out[f"levels_{li}"] = np.asarray(imfilters.luma_adjusted_levels(ia, luma_min=0.3, gamma=1.3, gamma_luma_min=0.4,
                                                                         gamma_alpha=0.5, gamma_min=0.5))
        out[f"levels_plain_{li}"] = np.asarray(imfilters.luma_adjusted_levels(ia, luma_min=0.2, gamma=0.8, gamma_luma_min=0.6))
        out[f"stab_adaptive_{li}"] = np.asarray(imfilters.chroma_stabilizer_adaptive(Image.fromarray(a), ia, 14, 18, 1.0))
        out[f"stab_{li}"] = np.asarray(imfilters.chroma_stabilizer(Image.fromarray(a), ia, 0.1, 1.0))
        out[f"restore_w_{li}"] = np.asarray(restcolor.restore_color_gradient(ia, Image.fromarray(a), sat=0.7, tht=40, weight=0.3,
                                                                             alpha=3.0, algo=0))
        out[f"restore_wneg_{li}"] = np.asarray(restcolor.restore_color_gradient(ia, Image.fromarray(a), sat=1.0, tht=20, weight=-0.5,
                                                                                alpha=4.0, algo=1))
    np.savez_compressed(os.path.join(HERE, "vsslib_filters.npz"), **out)
    print("filters golden:", len(out), "arrays")


def golden_zhang():
    """ECCVGenerator / SIGGRAPHGenerator (colorization/colorizers/*.py), the real modules with synthetic weights
    loaded strict=True, on a seeded L input."""
    from oracle import zhang_oracle
    for name in ("eccv16", "siggraph17"):
        sd = zhang_oracle.make_zhang_state_dict(name, 1234)
        m = refshim.build_zhang(name)
        ref_sd = m.state_dict()
        assert set(ref_sd.keys()) == set(sd.keys()), (set(ref_sd) ^ set(sd))
        m.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(99)
        low = torch.rand(1, 1, 12, 12, generator=g)
        l = torch.nn.functional.interpolate(low, size=(96, 96), mode="bicubic", align_corners=False).clamp(0, 1) * 100.0
        with torch.no_grad():
            y = m(l)
        np.savez_compressed(os.path.join(HERE, f"zhang_{name}_96.npz"), l=l.numpy(), ab=y.numpy(), weight_seed=1234)
        print(name, "golden: ab mean/std/min/max", float(y.mean()), float(y.std()), float(y.min()), float(y.max()))


def golden_stabilizer():
    """Per-frame stages of HAVC_stabilizer (vs_dark_tweak, vs_chroma_bright_tweak, vs_colormap) run through the REAL
    reference selectors on clips of the VapourSynth stand-in."""
    refshim.install()
    from vsdeoldify_b200 import vs_shim
    sys.modules["vapoursynth"] = vs_shim
    from vsdeoldify.vsslib import vsfilters
    from vsdeoldify import havc_utils
    out = {}
    H, W = 48, 72
    for li, luma in enumerate((0.08, 0.3, 0.6)):
        a, _ = filter_test_pair(80 + li, H, W, luma)
        out[f"img_{li}"] = a
        clip = vs_shim.array_clip(np.ascontiguousarray(np.transpose(a, (2, 0, 1)))[None])
        get = lambda c: np.dstack([np.asarray(c.get_frame(0)[p]) for p in range(3)])
        out[f"dark_{li}"] = get(vsfilters.vs_dark_tweak(clip, dark_threshold=0.2, dark_amount=0.8))
        out[f"dark_hue_{li}"] = get(vsfilters.vs_dark_tweak(clip, dark_threshold=0.35, dark_amount=0.5, dark_hue_adjust="0:60,300:360"))
        out[f"smooth_{li}"] = get(vsfilters.vs_chroma_bright_tweak(clip, black_threshold=0.3, white_threshold=0.7, dark_sat=0.9, dark_bright=-0.0))
        out[f"smooth_adj_{li}"] = get(vsfilters.vs_chroma_bright_tweak(clip, black_threshold=0.25, white_threshold=0.25, dark_sat=0.7,
                                                                       dark_bright=-0.15, chroma_adjust="180:280|0.5,0.2"))
        for name in ("blue->brown", "red->blue", "yellow->rose"):
            adj = havc_utils._get_colormap(name)
            out[f"colormap_{name}_{li}"] = get(vsfilters.vs_colormap(clip, colormap=adj))
    np.savez_compressed(os.path.join(HERE, "vsslib_stabilizer.npz"), **out)
    print("stabilizer golden:", len(out), "arrays")


def temporal_test_clip(seed, T, h, w):
    """A short colourful clip with motion (a pattern drifting over a fixed background), gray areas and a luma ramp over time so
    that both branches of the frame-luma gate and the gray mask of restore_color are exercised."""
    rng = np.random.default_rng(seed)
    a, b = filter_test_pair(seed, h, w + T, 0.45)
    frames = []
    for t in range(T):
        img = a[:, t:t + w].astype(np.float32) * 0.7 + b[:, T - t:T - t + w].astype(np.float32) * 0.3
        img[h // 3: h // 2, w // 4: w // 2] = img[h // 3: h // 2, w // 4: w // 2].mean(-1, keepdims=True)   # a gray box
        gain = 0.35 + 1.1 * t / max(T - 1, 1)                     # frame luma from ~0.15 to ~0.65
        img = img * gain + rng.normal(0, 1.5, img.shape)
        frames.append(np.clip(img, 0, 255).astype(np.uint8))
    return np.stack(frames)


def _install_vs_primitives(vs_shim):
    """Give the VapourSynth stand-in the two primitives the reference's temporal stabiliser calls — `clip.resize.Bicubic` between
    RGB24 and YUV420P8 and `std.AverageFrames` — backed by the restatements in oracle/ (zimg_oracle, temporal_oracle).  Golden
    generation only: this pins the reference's GRAPH (which frame / plane goes where), not the primitives (unpinned)."""
    from oracle import temporal_oracle as to, zimg_oracle as zo

    class _Resize:
        def __init__(self, clip):
            self.clip = clip

        def Bicubic(self, format=None, matrix_s=None, matrix_in_s=None, range_s=None, dither_type="none", **kw):
            src = self.clip
            fmt = format if isinstance(format, vs_shim.VideoFormat) else {1: vs_shim.RGB24, 3: vs_shim.YUV420P8}[int(format)]
            dither = dither_type == "error_diffusion"
            assert range_s == "full"

            def fn(n):
                f = src.get_frame(n)
                if src.format == vs_shim.RGB24 and fmt == vs_shim.YUV420P8:
                    assert matrix_s == "709"
                    planes = zo.rgb24_to_yuv420p8(np.dstack([np.asarray(f[p]) for p in range(3)]), "709", False, dither)
                elif src.format == vs_shim.YUV420P8 and fmt == vs_shim.RGB24:
                    assert matrix_in_s == "709"
                    rgb = zo.yuv420p8_to_rgb24(np.asarray(f[0]), np.asarray(f[1]), np.asarray(f[2]), dither, "709", False)
                    planes = [np.ascontiguousarray(rgb[..., p]) for p in range(3)]
                else:
                    raise AssertionError("conversion not needed by the temporal stabiliser")
                return vs_shim.VideoFrame(list(planes), fmt, dict(f.props))
            return vs_shim.VideoNode(src.num_frames, src.width, src.height, fmt, fn, src.fps_num, src.fps_den)

    vs_shim.VideoNode.resize = property(lambda self: _Resize(self))

    def AverageFrames(self, clips=None, weights=None, scale=None, scenechange=None, planes=None):
        clips = self._clip if clips is None else clips
        single = isinstance(clips, vs_shim.VideoNode) or len(clips) == 1
        lst = [clips] if isinstance(clips, vs_shim.VideoNode) else list(clips)
        base = lst[0]
        planes_ = [0, 1, 2] if planes is None else list(planes)
        sc = sum(weights) if scale is None else scale

        def fn(n):
            if single:
                r = len(weights) // 2
                fr = [base.get_frame(min(max(n + d, 0), base.num_frames - 1)) for d in range(-r, r + 1)]
                w = list(weights)
                if scenechange:
                    w = to.scene_folded_weights(w, [f.props.get("_SceneChangePrev", 0) for f in fr], [f.props.get("_SceneChangeNext", 0) for f in fr])
                center = fr[r]
            else:
                fr, w, center = [c.get_frame(n) for c in lst], list(weights), lst[0].get_frame(n)
            out = [to.average_planes([np.asarray(f[p]) for f in fr], w, sc) if p in planes_ else np.asarray(center[p]).copy()
                   for p in range(base.format.num_planes)]
            return vs_shim.VideoFrame(out, base.format, dict(center.props))
        return vs_shim.VideoNode(base.num_frames, base.width, base.height, base.format, fn, base.fps_num, base.fps_den)

    vs_shim._Std.AverageFrames = lambda self, clips=None, weights=None, scale=None, scenechange=None, planes=None: \
        AverageFrames(self, clips, weights, scale, scenechange, planes)


def golden_temporal():
    """Scope row N3: the REAL vs_chroma_stabilizer_ex / vs_clip_color_stabilizer / vs_sc_recover_clip_color / restore_color /
    weight-table code of the reference on the stand-in (see _install_vs_primitives for what that does and does not pin)."""
    refshim.install()
    from vsdeoldify_b200 import vs_shim
    sys.modules["vapoursynth"] = vs_shim
    _install_vs_primitives(vs_shim)
    from PIL import Image
    from vsdeoldify.vsslib import restcolor, vsfilters
    out = {}
    T, H, W = 20, 24, 32
    clip_np = temporal_test_clip(91, T, H, W)
    out["clip"] = clip_np
    for n in range(3, 16):
        out[f"avg_arith_{n}"] = np.array(vsfilters._build_avg_arithmetic(n))
        out[f"avg_weighted_{n}"] = np.array(vsfilters._build_avg_weighted(n))
    # restore_color, image level
    for k, (i, j, kw) in enumerate([(16, 17, dict(sat=1.0, tht=15, weight=0.2, tht_scen=0.8)),
                                    (5, 18, dict(sat=0.8, tht=40, weight=-0.5, tht_scen=0.8)),
                                    (19, 2, dict(sat=1.0, tht=30, weight=0.0, tht_scen=0.9, hue_adjust="0:60|0.8,0.1")),
                                    (10, 11, dict(sat=1.0, tht=250, weight=0.2, tht_scen=0.5))]):     # last: the tht_scen bypass
        out[f"restore_{k}"] = np.asarray(restcolor.restore_color(Image.fromarray(clip_np[i]), Image.fromarray(clip_np[j]), **kw))
    props = [{"_SceneChangePrev": int(n in (0, 9)), "_SceneChangeNext": int(n in (8, T - 1))} for n in range(T)]
    clip = vs_shim.array_clip(np.ascontiguousarray(np.transpose(clip_np, (0, 3, 1, 2))), props=props)
    get = lambda c, n: np.dstack([np.asarray(c.get_frame(n)[p]) for p in range(3)])
    # the selector of vs_sc_recover_clip_color (n < 15 passes through; the luma gate flips the weight on dark frames)
    other = vs_shim.array_clip(np.ascontiguousarray(np.transpose(clip_np[::-1], (0, 3, 1, 2))))
    rec = vsfilters.vs_recover_clip_color(clip=other, clip_color=clip, sat=0.9, tht=25, weight=0.3, tht_scen=0.8)
    for n in (3, 15, 19):
        out[f"recover_{n}"] = get(rec, n)
    frames = (0, 1, 8, 9, 14, 15, 16, 18, 19)
    out["frames"] = np.array(frames)
    st = vsfilters.vs_chroma_stabilizer_ex(clip, nframes=5, mode="A", sat=1.0, tht=15, weight=0.2, tht_scen=0.8, hue_adjust="none", algo=0)
    for n in frames:
        out[f"stab_default_{n}"] = get(st, n)
    st = vsfilters.vs_chroma_stabilizer_ex(clip, nframes=3, mode="W", sat=0.8, tht=30, weight=-0.4, tht_scen=0.8,
                                           hue_adjust="0:60|0.8,0.1", algo=0)
    for n in frames:
        out[f"stab_w3_{n}"] = get(st, n)
    st = vsfilters.vs_chroma_stabilizer_ex(clip, nframes=7, mode="A", tht=0)              # vs_clip_color_stabilizer, scenechange=True
    for n in frames:
        out[f"stab_tht0_{n}"] = get(st, n)
    np.savez_compressed(os.path.join(HERE, "vsslib_temporal.npz"), **out)
    print("temporal golden:", len(out), "arrays")


if __name__ == "__main__":
    if len(sys.argv) > 1:        # regenerate selected fixtures only: python make_golden.py golden_render_narrow ...
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    golden_render_narrow()
    golden_unets()
    golden_pixels()
    golden_filters()
    golden_zhang()
    golden_stabilizer()
    golden_temporal()
    golden_render()
    print("golden fixtures written to", HERE)
