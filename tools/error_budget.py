#!/usr/bin/env python
"""Per-layer error budget of 16-bit operand storage for DynamicUnetWide @ S x S (CPU emulation on the fp32 oracle).

For each layer group the conv inputs AND weights of that group alone are rounded to fp16 (fp32 accumulate, exactly what
the tensor-core path does: one rounding of the stored activation, one of the folded weight), everything else stays fp32;
the logit RMS error against the all-fp32 oracle is that group's contribution.  `all` rounds every group (the product's
fp16 path); `split` emulates the hi+lo operand split (value = hi + lo, both fp16, products hi*hi + lo*hi + hi*lo).

Usage: python tools/error_budget.py [--size 384] [--seed 1234] [--dtype fp16|bf16]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth_weights, unet_oracle, pixel_oracle as px  # noqa: E402

GROUPS = ["enc.stem", "enc.layer1", "enc.layer2", "enc.layer3", "enc.layer4", "middle", "block0", "block1", "block2", "block3",
          "shuf8", "res0", "res1", "head"]


def group_of(p: str) -> str:
    if p.startswith("layers.0.0"):
        return "enc.stem"
    for li in (4, 5, 6, 7):
        if p.startswith(f"layers.0.{li}."):
            return f"enc.layer{li - 3}"
    if p.startswith("layers.3."):
        return "middle"
    for i in range(4):
        if p.startswith(f"layers.{4 + i}."):
            return f"block{i}"
    if p.startswith("layers.8."):
        return "shuf8"
    if p.startswith("layers.10.layers.0"):
        return "res0"
    if p.startswith("layers.10.layers.1"):
        return "res1"
    if p.startswith("layers.11"):
        return "head"
    return "other"


def run(sd, x, active, dt, split=False):
    """oracle forward with fp16/bf16 rounding of conv operands in the `active` groups."""
    orig_conv = unet_oracle._Ctx.conv

    def rnd(t):
        if not split:
            return t.to(dt).float()
        hi = t.to(dt).float()
        return hi + (t - hi).to(dt).float()

    def conv(self, xin, p, stride=1, padding=0):
        if group_of(p) in active:
            w = unet_oracle.conv_weight(self.sd, p)
            if p + ".running_mean" in self.sd:
                pass
            return F.conv2d(rnd(xin), rnd(w), self.sd.get(p + ".bias"), stride=stride, padding=padding)
        return orig_conv(self, xin, p, stride, padding)

    unet_oracle._Ctx.conv = conv
    try:
        taps = {}
        unet_oracle.unet_forward(sd, x, taps=taps)
    finally:
        unet_oracle._Ctx.conv = orig_conv
    return taps["logits"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=384)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--dtype", default="fp16")
    a = ap.parse_args()
    dt = torch.float16 if a.dtype == "fp16" else torch.bfloat16
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth_weights.make_unet_state_dict("wide", a.seed)
    g = synth_weights.make_test_frame(7, a.size, a.size).numpy()
    x = torch.from_numpy(px.normalize_gray(g))[None]
    ref = run(sd, x, set(), dt)
    print(f"wide @{a.size}, seed {a.seed}, {a.dtype}: logit std {float(ref.std()):.3f}")
    tot2 = 0.0
    for grp in GROUPS:
        e = run(sd, x, {grp}, dt) - ref
        r = float((e ** 2).mean().sqrt())
        tot2 += r * r
        print(f"  {grp:12s} logit rms err {r:.3e}")
    print(f"  root-sum-square of the groups {tot2 ** 0.5:.3e}")
    e = run(sd, x, set(GROUPS), dt) - ref
    print(f"  all groups at once            {float((e ** 2).mean().sqrt()):.3e}")
    e = run(sd, x, set(GROUPS), dt, split=True) - ref
    print(f"  all groups, hi+lo split        {float((e ** 2).mean().sqrt()):.3e}")


if __name__ == "__main__":
    main()
