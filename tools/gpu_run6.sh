#!/bin/bash
# Pixel-pass / im2col rewrites + pair policy + stabilizer stages: tests, per-op timing, bench A/B (legacy kernels via env switches).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r6_tests.txt; cat gpurun_out/r6_tests.txt
timeout 300 python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_r6.json > gpurun_out/ops_b32_r6.txt 2>&1; head -24 gpurun_out/ops_b32_r6.txt; tail -8 gpurun_out/ops_b32_r6.txt
timeout 300 python bench.py --batch 32 --steps 8 --cpu-frames 0 > gpurun_out/bench_b32_r6.json 2> gpurun_out/bench_b32_r6.err; cut -c1-200 gpurun_out/bench_b32_r6.json
HAVC_B200_LEGACY_PIXEL=1 HAVC_B200_LEGACY_IM2COL=1 HAVC_B200_PAIR=0 timeout 300 python bench.py --batch 32 --steps 8 --cpu-frames 0 > gpurun_out/bench_b32_r6_legacy.json 2> gpurun_out/bench_b32_r6_legacy.err; cut -c1-200 gpurun_out/bench_b32_r6_legacy.json
timeout 300 python bench.py --batch 32 --steps 8 --cpu-frames 0 > gpurun_out/bench_b32_r6b.json 2> gpurun_out/bench_b32_r6b.err; cut -c1-200 gpurun_out/bench_b32_r6b.json
