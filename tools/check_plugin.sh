#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_surface.py -m gpu -q -x 2>&1 | tail -3
timeout 200 python - <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
from oracle import synth_weights
sd = synth_weights.make_unet_state_dict("wide", 1234)
clip = bench.synth_clip(128, 1080, 1920, seed=100)
print("plugin_surface", bench.plugin_surface_fps(sd, clip, 128, 32))
PY
