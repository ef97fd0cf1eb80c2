#!/usr/bin/env python
"""Prints an md5 of the colourised bytes of a fixed clip (to compare kernel variants selected by HAVC_B200_* switches)."""
import hashlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from oracle import synth_weights
from vsdeoldify_b200.engine import DeoldifyEngine
W, H, rf, B = int(os.environ.get("DBG_W", 1920)), int(os.environ.get("DBG_H", 1080)), int(os.environ.get("DBG_RF", 24)), 2
sd = synth_weights.make_unet_state_dict("wide", 1234)
eng = DeoldifyEngine(sd, W, H, render_factor=rf, batch=B, dtype=torch.float16)
out = eng.colorize_batch(bench.synth_clip(B, H, W, seed=7))
print("md5", hashlib.md5(out.tobytes()).hexdigest(), out.shape)
