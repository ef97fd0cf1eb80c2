#!/usr/bin/env python
"""Micro-benchmark of havc_conv_gemm on synthetic shapes (CUDA events, L2 flushed by rotating buffers)."""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vsdeoldify_b200 import ops

CASES = {
    # name: (B, H, W, cins, Cout, ks, bn, extra)
    "res0_272":   (8, 384, 384, [264], 259, 3, None, dict(bias=True, relu1=True)),
    "res1_272":   (8, 384, 384, [264], 259, 3, None, dict(bias=True, relu1=True, residual=True)),
    "c256_k256":  (8, 384, 384, [256], 256, 3, 256, dict(bias=True, relu1=True)),
    "c256_k264":  (8, 384, 384, [264], 256, 3, 256, dict(bias=True, relu1=True)),
    "c256_k320":  (8, 384, 384, [320], 256, 3, 256, dict(bias=True, relu1=True)),
    "c128_k264":  (8, 384, 384, [264], 128, 3, 128, dict(bias=True, relu1=True)),
    "shuf8":      (8, 192, 192, [256], 1024, 1, 256, dict(bias=True, relu1=True, shuffle=True)),
    "l6conv":     (8, 96, 96, [512, 256], 512, 3, 256, dict(relu1=True, affine=True)),
    "enc_l3_c2":  (8, 24, 24, [256], 256, 3, None, dict(bias=True, relu1=True)),
    "enc_l3_c2_b32": (32, 24, 24, [256], 256, 3, None, dict(bias=True, relu1=True)),
    "enc_l3_c1_b32": (32, 24, 24, [1024], 256, 1, None, dict(bias=True, relu1=True)),
    "enc_l3_c3_b32": (32, 24, 24, [256], 1024, 1, None, dict(bias=True, residual=True, relu2=True)),
    "enc_l1_c3_b16": (16, 96, 96, [64], 256, 1, None, dict(bias=True, residual=True, relu2=True)),
    "enc_l1_c1_b16": (16, 96, 96, [256], 64, 1, None, dict(bias=True, relu1=True)),
    "enc_l2_c3_b16": (16, 48, 48, [128], 512, 1, None, dict(bias=True, residual=True, relu2=True)),
    "middle0_b8": (8, 12, 12, [2048], 4096, 3, None, dict(relu1=True, affine=True)),
    "middle0_b32": (32, 12, 12, [2048], 4096, 3, None, dict(relu1=True, affine=True)),
}

def build(name, dtype=torch.float16, box=None, bn_override=None):
    B, H, W, cins, Cout, ks, bn, ex = CASES[name]
    dev = "cuda"
    srcs = [torch.randn(B, H, W, ops.pad_to(c, 8), device=dev).to(dtype) for c in cins]
    w = torch.randn(Cout, sum(cins), ks, ks) / (sum(cins) * ks * ks) ** 0.5
    shuffle = ex.get("shuffle", False)
    wp, meta = ops.pack_conv_weight(w, cins, dtype=dtype, shuffle=shuffle)
    wp = wp.to(dev)
    n_total = meta["rows"]
    vec = lambda: torch.randn(n_total, device=dev)
    if shuffle:
        out = torch.empty(B, 2 * H, 2 * W, ops.pad_to(Cout // 4, 8), device=dev, dtype=dtype)
    else:
        out = torch.empty(B, H, W, ops.pad_to(Cout, 8), device=dev, dtype=dtype)
    res = torch.randn(B, H, W, ops.pad_to(Cout, 8), device=dev).to(dtype) if ex.get("residual") else None
    op = ops.make_conv(srcs[0], wp, out, ops.taps_for(ks), src1=srcs[1] if len(srcs) > 1 else None, w_c1_off=meta["c1_off"],
                       n_total=n_total, bn=bn_override or bn, box=box, bias=vec() if ex.get("bias") else None,
                       scale=vec() if ex.get("affine") else None, shift=vec() if ex.get("affine") else None,
                       relu1=ex.get("relu1", False), relu2=ex.get("relu2", False), residual=res, shuffle=shuffle,
                       group_n=meta.get("group_n", 0))
    flops = 2.0 * B * H * W * Cout * sum(cins) * ks * ks
    return op, flops

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default=",".join(CASES))
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--box", default="")
    a = ap.parse_args()
    box = tuple(int(v) for v in a.box.split("x")) if a.box else None
    for name in a.cases.split(","):
        op, flops = build(name, box=box)
        d = op.desc
        for _ in range(2):
            op.launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            op.launch()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        print(f"{name:16s} {ms:8.4f} ms  {flops / ms / 1e9:7.1f} TF  BN={d.BN} N={d.N_total} box=({d.box_w},{d.box_h},{d.box_b})", flush=True)
