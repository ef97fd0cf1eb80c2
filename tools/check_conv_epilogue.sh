#!/bin/bash
# Warp-private fast epilogue: parity tests, per-op timing, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r10_convtests.txt; cat gpurun_out/r10_convtests.txt
if grep -q "failed\|rror" gpurun_out/r10_convtests.txt; then exit 1; fi
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_conv_gemm.py 2>&1 | tail -8 > gpurun_out/r10_tests.txt; cat gpurun_out/r10_tests.txt
timeout 300 python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_r10.json > gpurun_out/ops_b32_r10.txt 2>&1; head -34 gpurun_out/ops_b32_r10.txt; tail -7 gpurun_out/ops_b32_r10.txt
timeout 300 python bench.py --batch 32 --steps 8 --cpu-frames 0 > gpurun_out/bench_b32_r10.json 2> gpurun_out/bench_b32_r10.err; cut -c1-200 gpurun_out/bench_b32_r10.json
