"""TEST INFRASTRUCTURE — seeded synthetic weights in the REFERENCE's state-dict schema.

The trained weights (ColorizeVideo_gen.pth, ...) are user-downloaded files that are not in the reference
tree (README.md:44-77), so parity is judged on synthetic weights that both sides load:

  * the schema (key names, shapes, dtypes) is the reference's own (`Learner.load`,
    vsdeoldify/fastai/basic_train.py:264-286) — tests/golden/make_golden.py loads
    these dicts into the real reference modules with strict=True;
  * spectral-norm `weight_u/weight_v` are power-iterated so sigma is the true spectral norm, and every
    BatchNorm's running statistics are calibrated on a seeded batch (a raw random init saturates every
    output at +-3), so all layers carry signal;
  * real checkpoints drop in unchanged: nothing downstream knows the weights are synthetic.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import torch

from . import unet_oracle

SD = Dict[str, torch.Tensor]

# A random-init BatchNorm ResNet is chaotic (perturbations grow ~1.3x per block: measured 1e-3 -> 4e-1 over
# ResNet-101), which no trained network is; the last BN of every residual branch therefore gets a small gain,
# as trained ResNets have, so that parity measures arithmetic precision rather than amplified noise.
RESIDUAL_GAMMA = 0.1
HEAD_LOGIT_STD = 0.5

RESNET = {
    "wide": dict(block="bottleneck", layers=[3, 4, 23, 3]),   # resnet101
    "deep": dict(block="basic", layers=[3, 4, 6, 3]),         # resnet34
}


def _bn(sd, p, c, bias=0.0, gamma=1.0):
    sd[p + ".weight"] = torch.full((c,), gamma)
    sd[p + ".bias"] = torch.full((c,), bias)
    sd[p + ".running_mean"] = torch.zeros(c)
    sd[p + ".running_var"] = torch.ones(c)
    sd[p + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _kaiming(g, shape):
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5


def _sn_conv(sd, g, p, cout, cin, ks, bias=False, conv1d=False, iters=12):
    shape = (cout, cin, ks) if conv1d else (cout, cin, ks, ks)
    w = _kaiming(g, shape)
    wm = w.flatten(1)
    u = torch.nn.functional.normalize(torch.randn(cout, generator=g), dim=0)
    v = None
    for _ in range(iters):
        v = torch.nn.functional.normalize(torch.mv(wm.t(), u), dim=0)
        u = torch.nn.functional.normalize(torch.mv(wm, v), dim=0)
    if bias:
        sd[p + ".bias"] = torch.randn(cout, generator=g) * 0.05
    sd[p + ".weight_orig"] = w
    sd[p + ".weight_u"] = u
    sd[p + ".weight_v"] = v


def _encoder(sd, g, arch):
    cfg = RESNET[arch]
    sd["layers.0.0.weight"] = _kaiming(g, (64, 3, 7, 7))
    _bn(sd, "layers.0.1", 64)
    inplanes = 64
    exp = 4 if cfg["block"] == "bottleneck" else 1
    for li, (planes, n) in enumerate(zip([64, 128, 256, 512], cfg["layers"])):
        for bi in range(n):
            p = f"layers.0.{4 + li}.{bi}"
            stride = 2 if (bi == 0 and li > 0) else 1
            if cfg["block"] == "bottleneck":
                sd[p + ".conv1.weight"] = _kaiming(g, (planes, inplanes, 1, 1))
                _bn(sd, p + ".bn1", planes)
                sd[p + ".conv2.weight"] = _kaiming(g, (planes, planes, 3, 3))
                _bn(sd, p + ".bn2", planes)
                sd[p + ".conv3.weight"] = _kaiming(g, (planes * 4, planes, 1, 1))
                _bn(sd, p + ".bn3", planes * 4, gamma=RESIDUAL_GAMMA)
            else:
                sd[p + ".conv1.weight"] = _kaiming(g, (planes, inplanes, 3, 3))
                _bn(sd, p + ".bn1", planes)
                sd[p + ".conv2.weight"] = _kaiming(g, (planes, planes, 3, 3))
                _bn(sd, p + ".bn2", planes, gamma=RESIDUAL_GAMMA)
            if bi == 0 and (stride != 1 or inplanes != planes * exp):
                sd[p + ".downsample.0.weight"] = _kaiming(g, (planes * exp, inplanes, 1, 1))
                _bn(sd, p + ".downsample.1", planes * exp)
            inplanes = planes * exp
    # channels of encoder[6,5,4,2] and of the encoder output
    return [256 * exp, 128 * exp, 64 * exp, 64], 512 * exp


def make_unet_state_dict(arch: str = "wide", seed: int = 1234, calibrate: bool = True, calib_size: int = 192) -> SD:
    """arch 'wide' (video/stable) or 'deep' (artistic).  Deterministic for a given (arch, seed)."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = OrderedDict()
    skip_c, ni = _encoder(sd, g, arch)
    _bn(sd, "layers.1", ni, bias=1e-3)
    # middle_conv (unet.py:235-242)
    _sn_conv(sd, g, "layers.3.0.0", ni * 2, ni, 3)
    _bn(sd, "layers.3.0.2", ni * 2, 1e-3)
    _sn_conv(sd, g, "layers.3.1.0", ni, ni * 2, 3)
    _bn(sd, "layers.3.1.2", ni, 1e-3)
    x_c = ni
    for i, sc in enumerate(skip_c):
        p = f"layers.{4 + i}"
        not_final = i != len(skip_c) - 1
        if arch == "wide":
            nf = 512 * 2
            n_out = nf if not_final else nf // 2
            up_out = n_out // 2
            _sn_conv(sd, g, p + ".shuf.conv.0", up_out * 4, x_c, 1)
            _bn(sd, p + ".shuf.conv.1", up_out * 4, 1e-3)
            _bn(sd, p + ".bn", sc, 1e-3)
            cin = up_out + sc
            _sn_conv(sd, g, p + ".conv.0", n_out // 2, cin, 3)
            _bn(sd, p + ".conv.2", n_out // 2, 1e-3)
            att, x_c = p + ".conv.3", n_out // 2
        else:
            up_out = x_c // 2
            _sn_conv(sd, g, p + ".shuf.conv.0", up_out * 4, x_c, 1)
            _bn(sd, p + ".shuf.conv.1", up_out * 4, 1e-3)
            _bn(sd, p + ".bn", sc, 1e-3)
            cin = up_out + sc
            nf = int((cin if not_final else cin // 2) * 1.5)
            _sn_conv(sd, g, p + ".conv1.0", nf, cin, 3)
            _bn(sd, p + ".conv1.2", nf, 1e-3)
            _sn_conv(sd, g, p + ".conv2.0", nf, nf, 3)
            _bn(sd, p + ".conv2.2", nf, 1e-3)
            att, x_c = p + ".conv2.3", nf
        if i == len(skip_c) - 3:  # self attention (unet.py:250 / 136)
            sd[att + ".gamma"] = torch.tensor([0.6])
            _sn_conv(sd, g, att + ".query", x_c // 8, x_c, 1, conv1d=True)
            _sn_conv(sd, g, att + ".key", x_c // 8, x_c, 1, conv1d=True)
            _sn_conv(sd, g, att + ".value", x_c, x_c, 1, conv1d=True)
    # layers.8 PixelShuffle_ICNR: weight-norm conv with bias
    v = _kaiming(g, (x_c * 4, x_c, 1, 1))
    sd["layers.8.conv.0.bias"] = torch.randn(x_c * 4, generator=g) * 0.05
    sd["layers.8.conv.0.weight_g"] = v.flatten(1).norm(dim=1).view(-1, 1, 1, 1) * 1.5
    sd["layers.8.conv.0.weight_v"] = v
    nr = x_c + 3
    _sn_conv(sd, g, "layers.10.layers.0.0", nr, nr, 3, bias=True)
    _sn_conv(sd, g, "layers.10.layers.1.0", nr, nr, 3, bias=True)
    _sn_conv(sd, g, "layers.11.0", 3, nr, 1, bias=True)
    if calibrate:
        calibrate_bn(sd, seed, calib_size)
    return sd


def calibration_batch(seed: int, size: int, n: int = 2) -> torch.Tensor:
    """Seeded smooth+noise gray images, ImageNet-normalised like BaseFilter._model_process (filters.py:50-53)."""
    g = torch.Generator().manual_seed(seed + 77)
    low = torch.rand(n, 1, size // 8, size // 8, generator=g)
    img = torch.nn.functional.interpolate(low, size=(size, size), mode="bilinear", align_corners=False)
    img = (img + 0.08 * torch.randn(n, 1, size, size, generator=g)).clamp(0, 1)
    img = (img * 255).round() / 255
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    return (img.expand(n, 3, size, size) - mean) / std


def calibrate_bn(sd: SD, seed: int, size: int):
    """One train-mode-BN pass: every BN's running stats <- statistics of its own input, then a widening of
    the output head so chroma is spread over the (-3,3) range without saturating."""
    x = calibration_batch(seed, size)
    taps = {}
    unet_oracle.unet_forward(sd, x, calibrate=True, taps=taps)
    logits = taps["logits"]
    # W = weight_orig / sigma with sigma = u^T W v: dividing u by k multiplies W by k
    k = float(HEAD_LOGIT_STD / logits.std().clamp_min(1e-6))
    sd["layers.11.0.weight_u"] = sd["layers.11.0.weight_u"] / k
    b_old = sd["layers.11.0.bias"]
    sd["layers.11.0.bias"] = b_old - k * (logits.mean(dim=(0, 2, 3)) - b_old)   # new logit mean == b_old


def make_test_frame(seed: int, h: int, w: int) -> torch.Tensor:
    """Seeded uint8 gray test image [h, w] with structure (blobs + edges + noise)."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(1, 1, max(2, h // 16), max(2, w // 16), generator=g)
    img = torch.nn.functional.interpolate(low, size=(h, w), mode="bicubic", align_corners=False)[0, 0]
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    img = img + 0.25 * ((xx // max(1, w // 6) + yy // max(1, h // 5)) % 2).float() - 0.1
    img = img + 0.03 * torch.randn(h, w, generator=g)
    return (img.clamp(0, 1) * 255).round().to(torch.uint8)
