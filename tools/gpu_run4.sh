#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "surface or pipeline or unet or filters" 2>&1 | tail -5
python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_px.json 2>&1 | tail -7
python bench.py --batch 32 --steps 8 --cpu-frames 0 > gpurun_out/bench_b32_px.json 2> gpurun_out/bench_b32_px.err
cut -c1-200 gpurun_out/bench_b32_px.json
