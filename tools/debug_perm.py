#!/usr/bin/env python
"""Diagnostic: is a frame's output independent of its batch slot?  (run under different HAVC_B200_* switches)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from oracle import synth_weights
from vsdeoldify_b200.engine import DeoldifyEngine

W, H, rf, B = int(os.environ.get("DBG_W", 1920)), int(os.environ.get("DBG_H", 1080)), int(os.environ.get("DBG_RF", 24)), int(os.environ.get("DBG_B", 2))
sd = synth_weights.make_unet_state_dict("wide", 1234)
eng = DeoldifyEngine(sd, W, H, render_factor=rf, batch=B, dtype=torch.float16, debug_net_out=True, use_graph=os.environ.get("DBG_GRAPH", "1") == "1")
clip = bench.synth_clip(B, H, W, seed=7)
a = eng.colorize_batch(clip); na = eng.net_out.cpu().clone(); sa = eng.rgb_small.cpu().clone()
a2 = eng.colorize_batch(clip); na2 = eng.net_out.cpu().clone()
perm = np.arange(B)[::-1].copy()
b = eng.colorize_batch(np.ascontiguousarray(clip[perm])); nb = eng.net_out.cpu().clone(); sb = eng.rgb_small.cpu().clone()
def stat(x, y, what):
    d = np.abs(x.astype(np.float64) - y.astype(np.float64))
    print(f"{what}: differing {int((d > 0).sum())} of {d.size}, max {d.max():.4g}")
stat(a, a2, "repeat        out")
stat(na.numpy(), na2.numpy(), "repeat        net_out")
stat(b, a[perm], "permuted      out")
stat(nb.numpy(), na.numpy()[perm], "permuted      net_out")
stat(sb.numpy(), sa.numpy()[perm], "permuted      rgb_small")
for i in range(B):
    stat(nb.numpy()[i], na.numpy()[perm][i], f"  frame slot {i} net_out")
