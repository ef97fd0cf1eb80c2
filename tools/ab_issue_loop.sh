#!/bin/bash
# Uniform-issue rewrite of the producer / MMA warps + TMA-loaded residual: parity tests, per-op timing, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r8_convtests.txt; cat gpurun_out/r8_convtests.txt
if grep -q "failed\|rror" gpurun_out/r8_convtests.txt; then exit 1; fi
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_conv_gemm.py 2>&1 | tail -8 > gpurun_out/r8_tests.txt; cat gpurun_out/r8_tests.txt
timeout 300 python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_r8.json > gpurun_out/ops_b32_r8.txt 2>&1; head -30 gpurun_out/ops_b32_r8.txt; tail -7 gpurun_out/ops_b32_r8.txt
HAVC_B200_NO_RES_TMA=1 timeout 300 python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_r8_norestma.json > gpurun_out/ops_b32_r8_norestma.txt 2>&1; head -3 gpurun_out/ops_b32_r8_norestma.txt
HAVC_B200_PAIR=0 timeout 300 python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_r8_nopair.json > gpurun_out/ops_b32_r8_nopair.txt 2>&1; head -3 gpurun_out/ops_b32_r8_nopair.txt
HAVC_B200_PAIR=1 timeout 300 python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_r8_allpair.json > gpurun_out/ops_b32_r8_allpair.txt 2>&1; head -3 gpurun_out/ops_b32_r8_allpair.txt
timeout 300 python bench.py --batch 32 --steps 8 --cpu-frames 0 > gpurun_out/bench_b32_r8.json 2> gpurun_out/bench_b32_r8.err; cut -c1-200 gpurun_out/bench_b32_r8.json
