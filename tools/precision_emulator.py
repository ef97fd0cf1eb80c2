#!/usr/bin/env python
"""CPU emulation of the GPU launch program's STORAGE / OPERAND precision (DynamicUnetWide / Deep), layer group by layer group.

Mirrors vsdeoldify_b200/unet.py op by op (BatchNorm folded into the encoder convs, fused epilogues: one rounding per STORED
tensor, one per packed weight), computing every contraction in fp32 on the CPU.  A policy maps a layer group to
  'f32' : nothing rounded (the oracle)
  'x1'  : operands and stored outputs rounded once to the 16-bit type (one MMA per K step: the round-1 path)
  'x3'  : operands and stored outputs kept as hi + lo pairs of the 16-bit type (three MMAs per K step: hi*hi + lo*hi + hi*lo)
so that mixed-precision designs can be compared (logit RMS error vs the all-f32 run, final-frame parity) without GPU time.

Usage: python tools/precision_emulator.py [--arch wide] [--size 384] [--seed 1234] [--dtype fp16]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth_weights, unet_oracle, pixel_oracle as px  # noqa: E402
from oracle.unet_oracle import conv_weight  # noqa: E402

GROUPS = ["input", "enc.stem", "enc.layer1", "enc.layer2", "enc.layer3", "enc.layer4", "middle", "block0", "block1", "block2",
          "block3", "shuf8", "res"]
BN_EPS = 1e-5


class Emu:
    def __init__(self, sd, policy, dt=torch.float16):
        self.sd, self.policy, self.dt = sd, policy, dt

    # ---- precision primitives -----------------------------------------------------------------
    def q(self, t, mode):
        if mode == "f32":
            return t
        hi = t.to(self.dt).float()
        if mode == "x1":
            return hi
        return hi + (t - hi).to(self.dt).float()

    def bn(self, p):
        sd = self.sd
        sc = sd[p + ".weight"].double() / torch.sqrt(sd[p + ".running_var"].double() + BN_EPS)
        sh = sd[p + ".bias"].double() - sd[p + ".running_mean"].double() * sc
        return sc.float().view(1, -1, 1, 1), sh.float().view(1, -1, 1, 1)

    def conv(self, x, w, g, stride=1, padding=0):
        """one fused launch of group g: operands at g's precision, fp32 accumulate; the caller applies the epilogue + store"""
        m = self.policy[g]
        return F.conv2d(self.q(x, m), self.q(w, m), None, stride=stride, padding=padding)

    def fold(self, pc, pb):
        sc, sh = self.bn(pb)
        return self.sd[pc + ".weight"].float() * sc.view(-1, 1, 1, 1), sh

    # ---- network ------------------------------------------------------------------------------
    def bottleneck(self, x, p, stride, g):
        m = self.policy[g]
        w1, b1 = self.fold(p + ".conv1", p + ".bn1")
        w2, b2 = self.fold(p + ".conv2", p + ".bn2")
        w3, b3 = self.fold(p + ".conv3", p + ".bn3")
        c1 = self.q(F.relu(self.conv(x, w1, g) + b1), m)
        c2 = self.q(F.relu(self.conv(c1, w2, g, stride=stride, padding=1) + b2), m)
        idt = x
        if p + ".downsample.0.weight" in self.sd:
            wd, bd = self.fold(p + ".downsample.0", p + ".downsample.1")
            idt = self.q(self.conv(x, wd, g, stride=stride) + bd, m)
        return self.q(F.relu(self.conv(c2, w3, g) + b3 + idt), m)

    def basic(self, x, p, stride, g):
        m = self.policy[g]
        w1, b1 = self.fold(p + ".conv1", p + ".bn1")
        w2, b2 = self.fold(p + ".conv2", p + ".bn2")
        c1 = self.q(F.relu(self.conv(x, w1, g, stride=stride, padding=1) + b1), m)
        idt = x
        if p + ".downsample.0.weight" in self.sd:
            wd, bd = self.fold(p + ".downsample.0", p + ".downsample.1")
            idt = self.q(self.conv(x, wd, g, stride=stride) + bd, m)
        return self.q(F.relu(self.conv(c1, w2, g, padding=1) + b2 + idt), m)

    def attention(self, x, p, g):
        m = self.policy[g]
        sd = self.sd
        size = x.size()
        xf = x.view(size[0], size[1], -1)
        wq, wk, wv = (conv_weight(sd, f"{p}.{n}") for n in ("query", "key", "value"))
        q = lambda t: self.q(t, m)
        f = q(F.conv1d(q(xf), q(wq)))
        gk = q(F.conv1d(q(xf), q(wk)))
        h = q(F.conv1d(q(xf), q(wv)))
        beta = q(F.softmax(torch.bmm(f.permute(0, 2, 1).contiguous(), gk), dim=1))       # logits stay fp32, P is stored
        o = sd[p + ".gamma"] * torch.bmm(h, beta) + xf
        return q(o.view(*size).contiguous())

    def unet_block(self, up_in, skip, p, g):
        m = self.policy[g]
        sd = self.sd
        ws = conv_weight(sd, p + ".shuf.conv.0")
        sc, sh = self.bn(p + ".shuf.conv.1")
        t = self.q(F.relu(self.conv(up_in, ws * sc.view(-1, 1, 1, 1), g) + sh), m)
        u = self.q(unet_oracle.blur(F.pixel_shuffle(t, 2)), m)
        sc, sh = self.bn(p + ".bn")
        sb = self.q(F.relu(skip * sc + sh), m)
        y = torch.cat([u, sb], 1)
        convs = ["conv"] if p + ".conv.0.weight_orig" in sd else ["conv1", "conv2"]
        for cv in convs:
            bsc, bsh = self.bn(f"{p}.{cv}.2")
            y = self.q(F.relu(self.conv(y, conv_weight(sd, f"{p}.{cv}.0"), g, padding=1)) * bsc + bsh, m)
            if f"{p}.{cv}.3.gamma" in sd:
                y = self.attention(y, f"{p}.{cv}.3", g)
        return y

    def forward(self, x):
        sd, P = self.sd, self.policy
        with torch.no_grad():
            x_in = self.q(x, P["input"])                         # the normalised image the pre kernel stores
            w, b = self.fold("layers.0.0", "layers.0.1")
            g = "enc.stem"
            y = self.q(F.relu(self.conv(x_in, w, g, stride=2, padding=3) + b), P[g])
            stem = y
            y = F.max_pool2d(y, 3, 2, 1)
            bott = "layers.0.4.0.conv3.weight" in sd
            skips = []
            for li in (4, 5, 6, 7):
                bi = 0
                g = f"enc.layer{li - 3}"
                while f"layers.0.{li}.{bi}.conv1.weight" in sd:
                    stride = 2 if (bi == 0 and li > 4) else 1
                    y = (self.bottleneck if bott else self.basic)(y, f"layers.0.{li}.{bi}", stride, g)
                    bi += 1
                skips.append(y)
            skips = [skips[2], skips[1], skips[0], stem]
            g = "middle"
            sc, sh = self.bn("layers.1")
            y = self.q(F.relu(y * sc + sh), P[g])
            for j in (0, 1):
                sc, sh = self.bn(f"layers.3.{j}.2")
                y = self.q(F.relu(self.conv(y, conv_weight(sd, f"layers.3.{j}.0"), g, padding=1)) * sc + sh, P[g])
            for i, s in enumerate(skips):
                y = self.unet_block(y, s, f"layers.{4 + i}", f"block{i}")
            g = "shuf8"
            t = self.q(F.relu(self.conv(y, conv_weight(sd, "layers.8.conv.0"), g) + sd["layers.8.conv.0.bias"].view(1, -1, 1, 1)), P[g])
            u = self.q(unet_oracle.blur(F.pixel_shuffle(t, 2)), P[g])
            g = "res"
            cat = torch.cat([u, x_in], 1)
            r = self.q(F.relu(self.conv(cat, conv_weight(sd, "layers.10.layers.0.0"), g, padding=1)
                              + sd["layers.10.layers.0.0.bias"].view(1, -1, 1, 1)), P[g])
            r2 = F.relu(self.conv(r, conv_weight(sd, "layers.10.layers.1.0"), g, padding=1)
                        + sd["layers.10.layers.1.0.bias"].view(1, -1, 1, 1)) + cat
            # fused fp32 head on the fp32 epilogue values
            return F.conv2d(r2, conv_weight(sd, "layers.11.0"), sd["layers.11.0.bias"])


def logits(sd, x, policy, dt=torch.float16):
    return Emu(sd, policy, dt).forward(x)


def uniform(mode):
    return {g: mode for g in GROUPS}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="wide")
    ap.add_argument("--size", type=int, default=384)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--dtype", default="fp16")
    a = ap.parse_args()
    dt = torch.float16 if a.dtype == "fp16" else torch.bfloat16
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth_weights.make_unet_state_dict(a.arch, a.seed)
    gimg = synth_weights.make_test_frame(7, a.size, a.size).numpy()
    x = torch.from_numpy(px.normalize_gray(gimg))[None]
    ref = logits(sd, x, uniform("f32"), dt)
    taps = {}
    unet_oracle.unet_forward(sd, x, taps=taps)
    print(f"{a.arch} @{a.size} seed {a.seed} {a.dtype}: logit std {float(ref.std()):.3f}; emulator(f32) vs oracle max |d| "
          f"{float((ref - taps['logits']).abs().max()):.2e}")
    rms = lambda e: float((e ** 2).mean().sqrt())
    print("one group at x1, the rest f32 (that group's contribution):")
    for g in GROUPS:
        pol = uniform("f32")
        pol[g] = "x1"
        print(f"  {g:11s} {rms(logits(sd, x, pol, dt) - ref):.3e}")
    print(f"all x1: {rms(logits(sd, x, uniform('x1'), dt) - ref):.3e}")
    print(f"all x3: {rms(logits(sd, x, uniform('x3'), dt) - ref):.3e}")
    for name, x3g in (("encoder+input x3", ["input", "enc.stem", "enc.layer1", "enc.layer2", "enc.layer3", "enc.layer4"]),
                      ("encoder+input+blocks x3", ["input", "enc.stem", "enc.layer1", "enc.layer2", "enc.layer3", "enc.layer4", "middle",
                                                   "block0", "block1", "block2", "block3"])):
        pol = uniform("x1")
        for g in x3g:
            pol[g] = "x3"
        print(f"{name}, rest x1: {rms(logits(sd, x, pol, dt) - ref):.3e}")


if __name__ == "__main__":
    main()
