"""TEST INFRASTRUCTURE (parity oracle) — Zhang et al. colorizers (eccv16 / siggraph17) and their LAB pre/post
processing, restated functionally (torch fp32 on CPU for the networks, numpy float64 for LAB).

Reference: vsdeoldify/colorization/colorizers/eccv16.py:9-98, siggraph17.py:7-161, base_color.py:5-23,
util.py:21-55, colorization/__init__.py:76-95 (ModelColorization.colorize_frame).

Pins: the network restatements are checked against the REAL reference modules (imported through oracle/refshim.py)
by tests/golden/make_golden.py -> tests/golden/zhang_*.npz; Pillow's BICUBIC resize is pinned bit-exact
(pixel_oracle.pil_resize).  scikit-image is NOT installed here, so `rgb2lab` / `lab2rgb` follow the published
definitions and that edge is PARITY UNPINNED.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from . import pixel_oracle as px

SD = Dict[str, torch.Tensor]

# ---- network definitions as data: (kind, cin, cout, ks, stride, dilation) per conv, BN at the end of a block -------
# eccv16.py:13-72
ECCV_BLOCKS = {
    "model1": [(1, 64, 1, 1), (64, 64, 2, 1)],
    "model2": [(64, 128, 1, 1), (128, 128, 2, 1)],
    "model3": [(128, 256, 1, 1), (256, 256, 1, 1), (256, 256, 2, 1)],
    "model4": [(256, 512, 1, 1), (512, 512, 1, 1), (512, 512, 1, 1)],
    "model5": [(512, 512, 1, 2)] * 3,
    "model6": [(512, 512, 1, 2)] * 3,
    "model7": [(512, 512, 1, 1)] * 3,
}
# siggraph17.py:11-60 (all stride 1: the down-sampling is the [::2, ::2] slicing in forward)
SIG_BLOCKS = {
    "model1": [(4, 64, 1, 1), (64, 64, 1, 1)],
    "model2": [(64, 128, 1, 1), (128, 128, 1, 1)],
    "model3": [(128, 256, 1, 1), (256, 256, 1, 1), (256, 256, 1, 1)],
    "model4": [(256, 512, 1, 1), (512, 512, 1, 1), (512, 512, 1, 1)],
    "model5": [(512, 512, 1, 2)] * 3,
    "model6": [(512, 512, 1, 2)] * 3,
    "model7": [(512, 512, 1, 1)] * 3,
}
BN_EPS = 1e-5


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def _block(sd: SD, name: str, spec, x, calibrate=False):
    """conv -> ReLU repeated, BatchNorm last (index 2*len(spec) of the nn.Sequential)."""
    for i, (cin, cout, stride, dil) in enumerate(spec):
        p = f"{name}.{2 * i}"
        x = F.relu(F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=dil, dilation=dil))
    p = f"{name}.{2 * len(spec)}"
    if calibrate:
        sd[p + ".running_mean"] = x.mean(dim=(0, 2, 3))
        sd[p + ".running_var"] = x.var(dim=(0, 2, 3), unbiased=False) + 1e-3
    return _bn(sd, p, x)


def eccv16_forward(sd: SD, l: torch.Tensor, calibrate: bool = False, taps=None) -> torch.Tensor:
    """ECCVGenerator.forward (eccv16.py:87-98): L [B,1,256,256] -> ab [B,2,256,256]."""
    x = (l - 50.0) / 100.0
    for name, spec in ECCV_BLOCKS.items():
        x = _block(sd, name, spec, x, calibrate)
        if taps is not None:
            taps[name] = x
    x = F.relu(F.conv_transpose2d(x, sd["model8.0.weight"], sd["model8.0.bias"], stride=2, padding=1))
    x = F.relu(F.conv2d(x, sd["model8.2.weight"], sd["model8.2.bias"], padding=1))
    x = F.relu(F.conv2d(x, sd["model8.4.weight"], sd["model8.4.bias"], padding=1))
    x = F.conv2d(x, sd["model8.6.weight"], sd["model8.6.bias"])
    if taps is not None:
        taps["logits"] = x
    out_reg = F.conv2d(F.softmax(x, dim=1), sd["model_out.weight"])
    if taps is not None:
        taps["out_reg"] = out_reg
    return F.interpolate(out_reg, scale_factor=4, mode="bilinear") * 110.0


def siggraph17_forward(sd: SD, l: torch.Tensor, calibrate: bool = False, taps=None) -> torch.Tensor:
    """SIGGRAPHGenerator.forward with input_B = mask_B = 0 (siggraph17.py:133-161)."""
    z = l * 0
    x = torch.cat(((l - 50.0) / 100.0, z / 110.0, z / 110.0, z), dim=1)
    c1 = _block(sd, "model1", SIG_BLOCKS["model1"], x, calibrate)
    c2 = _block(sd, "model2", SIG_BLOCKS["model2"], c1[:, :, ::2, ::2], calibrate)
    c3 = _block(sd, "model3", SIG_BLOCKS["model3"], c2[:, :, ::2, ::2], calibrate)
    x = _block(sd, "model4", SIG_BLOCKS["model4"], c3[:, :, ::2, ::2], calibrate)
    for name in ("model5", "model6", "model7"):
        x = _block(sd, name, SIG_BLOCKS[name], x, calibrate)
    if taps is not None:
        taps.update(model1=c1, model2=c2, model3=c3, model7=x)

    def bn_tail(p, y):
        if calibrate:
            sd[p + ".running_mean"] = y.mean(dim=(0, 2, 3))
            sd[p + ".running_var"] = y.var(dim=(0, 2, 3), unbiased=False) + 1e-3
        return _bn(sd, p, y)
    up = F.conv_transpose2d(x, sd["model8up.0.weight"], sd["model8up.0.bias"], stride=2, padding=1) + \
        F.conv2d(c3, sd["model3short8.0.weight"], sd["model3short8.0.bias"], padding=1)
    y = F.relu(up)
    y = F.relu(F.conv2d(y, sd["model8.1.weight"], sd["model8.1.bias"], padding=1))
    y = F.relu(F.conv2d(y, sd["model8.3.weight"], sd["model8.3.bias"], padding=1))
    c8 = bn_tail("model8.5", y)
    up = F.conv_transpose2d(c8, sd["model9up.0.weight"], sd["model9up.0.bias"], stride=2, padding=1) + \
        F.conv2d(c2, sd["model2short9.0.weight"], sd["model2short9.0.bias"], padding=1)
    y = F.relu(F.conv2d(F.relu(up), sd["model9.1.weight"], sd["model9.1.bias"], padding=1))
    c9 = bn_tail("model9.3", y)
    up = F.conv_transpose2d(c9, sd["model10up.0.weight"], sd["model10up.0.bias"], stride=2, padding=1) + \
        F.conv2d(c1, sd["model1short10.0.weight"], sd["model1short10.0.bias"], padding=1)
    y = F.leaky_relu(F.conv2d(F.relu(up), sd["model10.1.weight"], sd["model10.1.bias"], padding=1), 0.2)
    if taps is not None:
        taps.update(model8=c8, model9=c9, model10=y)
    out = torch.tanh(F.conv2d(y, sd["model_out.0.weight"], sd["model_out.0.bias"]))
    return out * 110.0


# ---- synthetic weights in the reference's state-dict schema ----------------------------------------------------------
def _kaiming(g, shape, fan_in):
    return torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5


def make_zhang_state_dict(name: str, seed: int = 1234, calibrate: bool = True) -> SD:
    """Seeded weights for 'eccv16' / 'siggraph17' with the key names and shapes of the reference modules
    (make_golden.py loads them with strict=True).  BN statistics are calibrated on a seeded L batch and the output
    layer is widened so that the ab output has a realistic spread (std ~ 20) without saturating."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = OrderedDict()
    blocks = ECCV_BLOCKS if name == "eccv16" else SIG_BLOCKS

    def conv(p, cin, cout, ks=3):
        sd[p + ".weight"] = _kaiming(g, (cout, cin, ks, ks), cin * ks * ks)
        sd[p + ".bias"] = torch.randn(cout, generator=g) * 0.05

    def convt(p, cin, cout):
        sd[p + ".weight"] = _kaiming(g, (cin, cout, 4, 4), cin * 4)      # each output sees 2x2 taps
        sd[p + ".bias"] = torch.randn(cout, generator=g) * 0.05

    def bn(p, c):
        sd[p + ".weight"] = torch.ones(c) + 0.1 * torch.randn(c, generator=g)
        sd[p + ".bias"] = 0.1 * torch.randn(c, generator=g) + 0.2
        sd[p + ".running_mean"] = torch.zeros(c)
        sd[p + ".running_var"] = torch.ones(c)
        sd[p + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    for bname, spec in blocks.items():
        for i, (cin, cout, stride, dil) in enumerate(spec):
            conv(f"{bname}.{2 * i}", cin, cout)
        bn(f"{bname}.{2 * len(spec)}", spec[-1][1])
    if name == "eccv16":
        convt("model8.0", 512, 256)
        conv("model8.2", 256, 256)
        conv("model8.4", 256, 256)
        conv("model8.6", 256, 313, ks=1)
        sd["model_out.weight"] = (torch.rand(2, 313, 1, 1, generator=g) * 2 - 1)
    else:
        convt("model8up.0", 512, 256)
        conv("model3short8.0", 256, 256)
        conv("model8.1", 256, 256)
        conv("model8.3", 256, 256)
        bn("model8.5", 256)
        convt("model9up.0", 256, 128)
        conv("model2short9.0", 128, 128)
        conv("model9.1", 128, 128)
        bn("model9.3", 128)
        convt("model10up.0", 128, 128)
        conv("model1short10.0", 64, 128)
        conv("model10.1", 128, 128)
        conv("model_class.0", 256, 529, ks=1)
        conv("model_out.0", 128, 2, ks=1)
    if calibrate:
        gl = torch.Generator().manual_seed(seed + 5)
        low = torch.rand(2, 1, 16, 16, generator=gl)
        l = F.interpolate(low, size=(128, 128), mode="bilinear", align_corners=False) * 100.0
        l = (l + 3.0 * torch.randn(2, 1, 128, 128, generator=gl)).clamp(0, 100)
        taps = {}
        if name == "eccv16":
            eccv16_forward(sd, l, calibrate=True, taps=taps)
            lg = taps["logits"]
            k = 4.0 / float(lg.std().clamp_min(1e-6))          # peaky soft-max -> ab spread
            sd["model8.6.weight"] = sd["model8.6.weight"] * k
            sd["model8.6.bias"] = (sd["model8.6.bias"] - lg.mean(dim=(0, 2, 3))) * k + 0.0
        else:
            siggraph17_forward(sd, l, calibrate=True, taps=taps)
            pre = F.conv2d(taps["model10"], sd["model_out.0.weight"], sd["model_out.0.bias"])
            k = 0.12 / float(pre.std().clamp_min(1e-6))
            sd["model_out.0.weight"] = sd["model_out.0.weight"] * k
            sd["model_out.0.bias"] = (sd["model_out.0.bias"] - pre.mean(dim=(0, 2, 3))) * k
    return sd


# ---- scikit-image LAB (restated; parity unpinned) ----------------------------------------------------------------------
_M_RGB2XYZ = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169], [0.019334, 0.119193, 0.950227]])
_WHITE = np.array([0.95047, 1.0, 1.08883])


def rgb2lab(rgb_u8: np.ndarray) -> np.ndarray:
    """skimage.color.rgb2lab on a uint8 image: float64 [H,W,3]."""
    c = rgb_u8.astype(np.float64) / 255.0
    lin = np.where(c > 0.04045, np.power((c + 0.055) / 1.055, 2.4), c / 12.92)
    xyz = lin @ _M_RGB2XYZ.T
    t = xyz / _WHITE
    f = np.where(t > 0.008856, np.cbrt(t), 7.787 * t + 16.0 / 116.0)
    L = 116.0 * f[..., 1] - 16.0
    a = 500.0 * (f[..., 0] - f[..., 1])
    b = 200.0 * (f[..., 1] - f[..., 2])
    return np.stack([L, a, b], -1)


_M_XYZ2RGB = np.linalg.inv(_M_RGB2XYZ)


def lab2rgb(lab: np.ndarray) -> np.ndarray:
    """skimage.color.lab2rgb: float64 [H,W,3] in [0,1] (negative Z clamped to 0, output clipped)."""
    lab = lab.astype(np.float64)
    fy = (lab[..., 0] + 16.0) / 116.0
    fx = lab[..., 1] / 500.0 + fy
    fz = fy - lab[..., 2] / 200.0
    fz = np.where(fz < 0, 0.0, fz)
    f = np.stack([fx, fy, fz], -1)
    xyz = np.where(f > 0.2068966, np.power(f, 3.0), (f - 16.0 / 116.0) / 7.787) * _WHITE
    rgb = xyz @ _M_XYZ2RGB.T
    out = np.where(rgb > 0.0031308, 1.055 * np.power(np.clip(rgb, 0, None), 1 / 2.4) - 0.055, 12.92 * rgb)
    return np.clip(out, 0, 1)


def colorize_frame(sd: SD, name: str, frame_rgb: np.ndarray) -> np.ndarray:
    """ModelColorization.colorize_frame (colorization/__init__.py:76-95): uint8 [H,W,3] -> uint8 [H,W,3].

    Pillow BICUBIC to 256 x 256, L of the original and of the resized image, network on the resized L, ab resized
    bilinearly to the frame size, lab2rgb, np.uint8(np.clip(x * 255, 0, 255))."""
    rs = px.pil_resize(frame_rgb, 256, 256, "bicubic")
    l_orig = torch.from_numpy(rgb2lab(frame_rgb)[..., 0]).float()[None, None]
    l_rs = torch.from_numpy(rgb2lab(rs)[..., 0]).float()[None, None]
    with torch.no_grad():
        ab = (eccv16_forward if name == "eccv16" else siggraph17_forward)(sd, l_rs)
        H, W = frame_rgb.shape[:2]
        if (H, W) != (256, 256):
            ab = F.interpolate(ab, size=(H, W), mode="bilinear")
        lab = torch.cat([l_orig, ab], 1)[0].permute(1, 2, 0).numpy()
    return np.uint8(np.clip(lab2rgb(lab) * 255, 0, 255))
