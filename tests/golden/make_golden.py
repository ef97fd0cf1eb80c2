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


if __name__ == "__main__":
    golden_unets()
    golden_pixels()
    golden_render()
    print("golden fixtures written to", HERE)
