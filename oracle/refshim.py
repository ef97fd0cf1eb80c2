"""TEST INFRASTRUCTURE — imports the REAL reference (dan64/vs-deoldify, read-only at /root/reference)
in-process so that golden vectors can be generated from the reference's own code.

Only `tests/golden/make_golden.py` (and tests that are skipped when /root/reference is absent) use this.
Nothing here ships: /root/reference does not exist on the GPU box.

Recipe:
  1. register a *namespace* package `vsdeoldify` so vsdeoldify/__init__.py (which imports vapoursynth,
     ColorMNet, ...) is never executed;
  2. stub `fastprogress` and `matplotlib`, which the vendored fastai imports at module load
     (vsdeoldify/fastai/imports/core.py:2,18,24-25; vsdeoldify/fastai/basic_train.py:8);
  3. build the bare `DynamicUnetWide/Deep` on a torchvision resnet body without pretrained weights
     (vsdeoldify/fastai/vision/learner.py:54-63 would download them).
"""
from __future__ import annotations

import importlib.metadata
import os
import sys
import types

REF_ROOT = os.environ.get("HAVC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "vsdeoldify", "deoldify"))


class _AnyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_any(name)


def _make_any(name):
    return _AnyMeta(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = sys.modules.get(self.__name__ + "." + name)
        if sub is not None:
            return sub
        obj = _make_any(name)
        setattr(self, name, obj)
        return obj


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    pkg = types.ModuleType("vsdeoldify")
    pkg.__path__ = [os.path.join(REF_ROOT, "vsdeoldify")]
    sys.modules["vsdeoldify"] = pkg
    for name in ("fastprogress", "fastprogress.fastprogress", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.patches", "matplotlib.patheffects"):
        if name not in sys.modules:
            m = _Stub(name)
            m.__path__ = []
            sys.modules[name] = m
    fp = sys.modules["fastprogress.fastprogress"]
    for n in ("MasterBar", "ProgressBar", "master_bar", "progress_bar", "format_time", "force_console_behavior"):
        getattr(fp, n)
    fp.IN_NOTEBOOK = False
    sys.modules["fastprogress"].fastprogress = fp
    _orig_version = importlib.metadata.version

    def _version(name):
        if name == "fastprogress":
            return "1.0.3"
        return _orig_version(name)

    importlib.metadata.version = _version
    _installed = True


def build_unet(arch: str = "wide"):
    """The reference's own generator module (random init): 'wide' = video/stable (ResNet-101,
    nf_factor 2, vsdeoldify/deoldify/generators.py:12-37), 'deep' = artistic (ResNet-34, nf_factor 1.5,
    generators.py:85-110)."""
    install()
    import torch.nn as nn
    import torchvision
    from vsdeoldify.deoldify.unet import DynamicUnetDeep, DynamicUnetWide
    from vsdeoldify.fastai.layers import NormType
    if arch == "wide":
        body = nn.Sequential(*list(torchvision.models.resnet101(weights=None).children())[:-2])
        m = DynamicUnetWide(body, n_classes=3, blur=True, blur_final=True, self_attention=True,
                            y_range=(-3.0, 3.0), norm_type=NormType.Spectral, last_cross=True, bottle=False,
                            nf_factor=2)
    elif arch == "deep":
        body = nn.Sequential(*list(torchvision.models.resnet34(weights=None).children())[:-2])
        m = DynamicUnetDeep(body, n_classes=3, blur=True, blur_final=True, self_attention=True,
                            y_range=(-3.0, 3.0), norm_type=NormType.Spectral, last_cross=True, bottle=False,
                            nf_factor=1.5)
    else:
        raise ValueError(arch)
    return m.eval()


def build_zhang(name: str):
    """ECCVGenerator / SIGGRAPHGenerator (vsdeoldify/colorization/colorizers/*.py) without touching
    vsdeoldify/colorization/__init__.py (which calls torch.cuda.set_device(0), line 28)."""
    install()
    import importlib.util
    base = os.path.join(REF_ROOT, "vsdeoldify", "colorization", "colorizers")
    pkgname = "_havc_ref_colorizers"
    if pkgname not in sys.modules:
        p = types.ModuleType(pkgname)
        p.__path__ = [base]
        sys.modules[pkgname] = p
    mods = {}
    for mod in ("base_color", "eccv16", "siggraph17"):
        full = f"{pkgname}.{mod}"
        if full not in sys.modules:
            spec = importlib.util.spec_from_file_location(full, os.path.join(base, mod + ".py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[full] = m
            spec.loader.exec_module(m)
        mods[mod] = sys.modules[full]
    if name == "eccv16":
        return mods["eccv16"].ECCVGenerator().eval()
    if name == "siggraph17":
        return mods["siggraph17"].SIGGRAPHGenerator().eval()
    raise ValueError(name)
