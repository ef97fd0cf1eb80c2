"""CPU tests of the plugin surface: names/signatures mirror the reference, graph-build errors are vs.Error, the
VapourSynth stand-in honours the selector protocol (props survive f.copy())."""
import inspect

import numpy as np
import pytest


def test_public_names_and_signatures():
    import vsdeoldify_b200 as pkg
    from vsdeoldify_b200 import havc
    for name in ("HAVC_main", "HAVC_colorizer", "HAVC_deoldify", "HAVC_ddeoldify"):
        assert callable(getattr(pkg, name))
    # positional order of the reference (vsdeoldify/__init__.py:2290-2298)
    params = list(inspect.signature(havc.HAVC_colorizer).parameters)
    assert params == ["clip", "method", "mweight", "deoldify_p", "ddcolor_p", "ddtweak", "ddtweak_p", "cmc_p", "lmm_p",
                      "alm_p", "crt_p", "cmb_sw", "sc_threshold", "sc_tht_offset", "sc_min_freq", "sc_tht_ssim",
                      "sc_normalize", "sc_min_int", "sc_tht_white", "sc_tht_black", "device_index", "torch_dir",
                      "debug_level"]
    sig = inspect.signature(havc.HAVC_colorizer)
    assert sig.parameters["method"].default == 2 and sig.parameters["mweight"].default == 0.4
    assert tuple(sig.parameters["deoldify_p"].default) == (0, 24, 1.0, 0.0)
    assert tuple(sig.parameters["ddcolor_p"].default) == (1, 24, 1.0, 0.0, True)
    # vsdeoldify/__init__.py:3612-3628
    dd = list(inspect.signature(havc.HAVC_ddeoldify).parameters)
    assert dd[:8] == ["clip", "method", "mweight", "deoldify_p", "ddcolor_p", "ddtweak", "ddtweak_p", "cmc_tresh"] and dd[-1] == "sc_debug"
    mp = inspect.signature(havc.HAVC_main).parameters
    assert mp["Preset"].default == "Medium" and mp["ColorModel"].default == "Video+Artistic"


def _clip(n=3, h=32, w=48):
    from vsdeoldify_b200 import vs_shim
    fr = np.random.default_rng(0).integers(0, 256, (n, 3, h, w), dtype=np.uint8)
    return vs_shim.array_clip(fr, props=[{"_SceneChangePrev": int(i == 0), "sc_threshold": 0.1, "x": i} for i in range(n)])


def test_graph_build_errors_are_vs_error():
    import torch
    from vsdeoldify_b200 import havc, vs_shim
    clip = _clip()
    with pytest.raises(vs_shim.Error, match="CPU mode"):
        havc.HAVC_colorizer(clip, method=0, device_index=99)
    if not torch.cuda.is_available():
        with pytest.raises(vs_shim.Error, match="CUDA is not available"):        # vsdeoldify/__init__.py:2441
            havc.HAVC_colorizer(clip, method=0)
    with pytest.raises(vs_shim.Error, match="Preset choice is invalid"):           # havc_utils.py:347
        havc.HAVC_main(clip, Preset="warp9", ColorModel="DeOldify(Video)")
    with pytest.raises(vs_shim.Error):
        havc.HAVC_main(clip, ColorModel="Video+Artistic")                          # DDColor side: not built
    with pytest.raises(vs_shim.Error):
        havc.HAVC_main(clip, ColorModel="DeOldify(Video)", EnableDeepEx=True)


def test_preset_table_matches_reference():
    from vsdeoldify_b200 import havc
    # havc_utils.py:338-340
    assert [havc._get_render_factor(p) for p in ("Placebo", "VerySlow", "Slower", "Slow", "Medium", "Fast", "Faster", "VeryFast")] \
        == [32, 32, 32, 28, 24, 22, 20, 16]


def test_shim_selector_protocol_keeps_props():
    from vsdeoldify_b200 import vs_shim
    clip = _clip()

    def sel(n, f):
        g = f.copy()
        np.copyto(np.asarray(g[0]), 255 - np.asarray(f[0]))
        return g
    out = clip.std.ModifyFrame(clips=[clip], selector=sel)
    assert out.num_frames == clip.num_frames and (out.width, out.height) == (48, 32)
    f2 = out.get_frame(2)
    assert f2.props == {"_SceneChangePrev": 0, "sc_threshold": 0.1, "x": 2}
    assert np.array_equal(np.asarray(f2[0]), 255 - np.asarray(clip.get_frame(2)[0]))
    assert np.array_equal(np.asarray(f2[1]), np.asarray(clip.get_frame(2)[1]))
    with pytest.raises(vs_shim.Error):
        out.get_frame(3)
    tagged = out.std.SetFrameProp(prop="sc_frequency", intval=1).std.CopyFrameProps(prop_src=clip, props=["x"])
    assert tagged.get_frame(1).props["sc_frequency"] == 1 and tagged.get_frame(1).props["x"] == 1


def test_havc_main_preset_tables_match_reference():
    """The string -> number tables behind HAVC_main (havc_utils.py:335-517), compared with the REAL reference functions when
    the reference tree is present (skipped on the GPU box, where it is not)."""
    import sys
    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference tree not present")
    refshim.install()
    from vsdeoldify_b200 import havc, vs_shim
    saved = sys.modules.get("vapoursynth")
    sys.modules["vapoursynth"] = vs_shim
    try:
        from vsdeoldify import havc_utils as ref
    finally:
        if saved is None:
            sys.modules.pop("vapoursynth", None)
        else:
            sys.modules["vapoursynth"] = saved
    for preset in ['Placebo', 'VerySlow', 'Slower', 'Slow', 'Medium', 'Fast', 'Faster', 'VeryFast']:
        assert havc._get_render_factor(preset) == ref._get_render_factors(preset)[1]
    for vt in ['VeryStable', 'MoreStable', 'Stable', 'Balanced', 'Vivid', 'MoreVivid', 'VeryVivid']:
        assert havc._VIDEO_TUNE[vt.lower()] == ref._get_mweight(vt)
    for cm in ['Simple', 'Constrained-Chroma', 'Luma-Masked', 'Adaptive-Luma', 'Chroma-Retention', 'ChromaBound Adaptive']:
        assert havc._COMB_METHOD[cm.lower()] == ref._get_comb_method(cm)
    for model in ['Video+Siggraph17', 'Stable+ECCV16', 'Artistic+Artistic', 'Video+ModelScope', 'DeOldify(Video)', 'DeOldify(Stable)',
                  'DeOldify(Artistic)', 'Zhang(Siggraph17)', 'Zhang(ECCV16)', 'DDColor(Artistic)']:
        assert havc._get_color_model(model) == ref._get_color_model(model), model
    for tune in ['None', 'Light', 'Medium', 'Strong']:
        for fix in ['None', 'Magenta', 'Magenta/Violet', 'Violet', 'Violet/Red', 'Blue/Magenta', 'Yellow', 'Yellow/Orange',
                    'Yellow/Green', 'Retinex/Red']:
            for dd in (0, 1, 2, 3):
                want = ref._get_color_tune(tune, fix, 'None', dd)
                assert havc._get_color_tune(tune, fix, dd) == (want[0], want[1]), (tune, fix, dd)
    # ColorMap -> "chroma adjustment" (havc_utils.py:519-548 inside _get_color_tune, :552-581 _get_colormap)
    maps = ['blue->brown', 'blue->red', 'blue->green', 'green->brown', 'green->red', 'green->blue', 'redrose->brown', 'redrose->blue',
            'red->brown', 'red->blue', 'yellow->rose', '30:90|+300,0.5']
    for tune in ['none', 'light', 'medium', 'strong']:
        for cm in maps:
            assert havc._get_colormap(cm, tune) == ref._get_color_tune(tune, 'None', cm, 0)[3], (tune, cm)
            assert havc._get_colormap(cm, tune) == ref._get_colormap(cm, tune), (tune, cm)
    assert havc._get_colormap("teal->pink") == ref._get_colormap("teal->pink")     # an unknown name passes through (fails later)
    with pytest.raises(vs_shim.Error, match="ColorMap choice is invalid"):
        havc._get_colormap("a|b|c")


def test_stabilizer_surface_and_errors():
    """HAVC_stabilizer mirrors vsdeoldify/__init__.py:2748-2751 (names, order, defaults) and rejects what the reference rejects."""
    from vsdeoldify_b200 import havc, vs_shim
    params = list(inspect.signature(havc.HAVC_stabilizer).parameters)
    assert params[:9] == ["clip", "dark", "dark_p", "smooth", "smooth_p", "stab", "stab_p", "colormap", "render_factor"]
    sig = inspect.signature(havc.HAVC_stabilizer).parameters
    assert tuple(sig["dark_p"].default) == (0.2, 0.8) and tuple(sig["smooth_p"].default) == (0.3, 0.7, 0.9, 0.0, "none")
    assert tuple(sig["stab_p"].default) == (5, 'A', 1, 15, 0.2, 0.8) and sig["render_factor"].default == 24
    clip = _clip()
    with pytest.raises(vs_shim.Error, match="render_factor must be between: 16-64"):       # :2796
        havc.HAVC_stabilizer(clip, dark=True, render_factor=12)
    # stab=True ends in the external ReduceFlicker plugin (vsplugins.py:263-272): without it, the reference's own error
    with pytest.raises(vs_shim.Error, match="ReduceFlicker.dll' not properly loaded/installed"):
        havc.HAVC_stabilizer(clip, stab=True)
    with pytest.raises(vs_shim.Error, match="algo=1"):
        havc.vs_chroma_stabilizer_ex(clip, tht=15, algo=1)


def test_deprecated_aliases_forward_like_the_reference(monkeypatch):
    """ddeoldify_main / ddeoldify / ddeoldify_stabilizer (vsdeoldify/__init__.py:3631-3664): same names, parameter order and
    forwarding as the reference's deprecated wrappers."""
    import vsdeoldify_b200 as pkg
    from vsdeoldify_b200 import havc
    from vsdeoldify_b200.constants import DEF_CRT_p
    assert list(inspect.signature(havc.ddeoldify_main).parameters) == ["clip", "Preset", "VideoTune", "ColorFix", "ColorTune", "ColorMap",
                                                                       "degrain_strength", "enable_fp16"]
    assert list(inspect.signature(havc.ddeoldify).parameters)[:9] == ["clip", "method", "mweight", "deoldify_p", "ddcolor_p", "dotweak",
                                                                      "dotweak_p", "ddtweak", "ddtweak_p"]
    assert inspect.signature(havc.ddeoldify_main).parameters["Preset"].default == "Fast"
    calls = {}
    monkeypatch.setattr(havc, "HAVC_main", lambda **kw: calls.setdefault("main", kw))
    monkeypatch.setattr(havc, "HAVC_colorizer", lambda *a, **kw: calls.setdefault("col", (a, kw)))
    monkeypatch.setattr(havc, "HAVC_stabilizer", lambda *a: calls.setdefault("stab", a))
    clip = _clip()
    pkg.ddeoldify_main(clip, Preset="Slow", ColorMap="red->brown")
    assert calls["main"]["Preset"] == "Slow" and calls["main"]["ColorMap"] == "red->brown" and calls["main"]["clip"] is clip
    havc.ddeoldify(clip, 3, 0.5, ddtweak=True, cmc_tresh=0.3)
    a, kw = calls["col"]
    assert a[1:3] == (3, 0.5) and a[5] == [True, False, False] and a[7] == [0.3] and a[10] == DEF_CRT_p
    assert kw["sc_threshold"] == 0 and kw["sc_min_freq"] == 0
    havc.ddeoldify_stabilizer(clip, True, (0.3, 0.7), colormap="blue->brown", render_factor=20)
    assert calls["stab"][1] is True and calls["stab"][7] == "blue->brown" and calls["stab"][8] == 20


def test_convert_and_restore_format_call_vapoursynth_like_the_reference(monkeypatch):
    """Under real VapourSynth convert_format_RGB24 / restore_format (havc_utils.py:57-237) are VapourSynth's own resize.Bicubic
    with the reference's arguments (zimg itself does the work: any format).  VapourSynth is not installable here, so a recording
    stand-in of the few calls is swapped in for `havc.vs` and the argument lists are compared with the reference's source."""
    import types
    from vsdeoldify_b200 import havc
    calls = []

    class Fmt:
        def __init__(self, id, family, bits):
            self.id, self.color_family, self.bits_per_sample = id, family, bits

        def replace(self, bits_per_sample):
            return Fmt(self.id + 1000, self.color_family, bits_per_sample)

    class Clip:
        def __init__(self, fmt, props=None):
            self.format, self.width, self.height, self.num_frames = fmt, 64, 48, 5
            self._props = props or {}
            self.std = types.SimpleNamespace(SetFrameProps=lambda **kw: (calls.append(("SetFrameProps", kw)), Clip(self.format, {**self._props, **kw}))[1])

        def get_frame(self, n):
            return types.SimpleNamespace(props=self._props)

    def bicubic(clip, **kw):
        calls.append(("Bicubic", kw))
        f = kw["format"]
        return Clip(f if isinstance(f, Fmt) else Fmt(f, "RGB" if f == 1 else "YUV", 8), clip._props)
    fake = types.SimpleNamespace(RGB24=1, YUV420P8=3, YUV="YUV", GRAY="GRAY", RGB="RGB", MATRIX_BT709=1, Error=havc.vs.Error,
                                 core=types.SimpleNamespace(core_version=types.SimpleNamespace(release_major=70),
                                                            resize=types.SimpleNamespace(Bicubic=bicubic)))
    monkeypatch.setattr(havc, "vs", fake)
    # a 10-bit BT.601 limited-range YUV clip
    src = Clip(Fmt(77, "YUV", 10), {"_Matrix": 6, "_ColorRange": 1})
    rgb, restore = havc.convert_format_RGB24(src)
    assert [c[0] for c in calls] == ["Bicubic", "Bicubic", "SetFrameProps"]
    assert calls[0][1]["format"].bits_per_sample == 8                                                     # :126-127
    assert {k: v for k, v in calls[1][1].items()} == dict(format=1, matrix_in=6, range_in_s="limited", range_s="full",
                                                          dither_type="error_diffusion")                  # :133-143
    assert calls[2][1] == {"_ColorRange": 0}                                                               # :160-163
    calls.clear()
    restore(rgb)
    assert calls == [("Bicubic", dict(format=77, matrix_in=1, matrix=6, range_in_s="full", range_s="limited",
                                      dither_type="error_diffusion"))]                                    # :199-207
    # GRAY8 full range: no dither on the way in, YUV420P8 BT.709 full range on the way out (:145-151, :208-222)
    calls.clear()
    rgb, restore = havc.convert_format_RGB24(Clip(Fmt(9, "GRAY", 8), {"_ColorRange": 0}))
    assert calls[0] == ("Bicubic", dict(format=1, range_in_s="limited", range_s="full"))
    calls.clear()
    restore(rgb)
    assert calls == [("Bicubic", dict(format=3, matrix=1, range_in_s="full", range_s="full", dither_type="error_diffusion"))]
    # RGB24 passes through untouched
    calls.clear()
    c = Clip(Fmt(1, "RGB", 8))
    rgb, restore = havc.convert_format_RGB24(c)
    assert rgb is c and restore(c) is c and calls == []
