#!/bin/bash
# CTA-pair (cta_group::2) bring-up: parity tests, then per-launch timing and the bench with pairs on and off.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_gemm.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pair_convtests.txt; cat gpurun_out/pair_convtests.txt
if grep -q "failed\|error\|Error" gpurun_out/pair_convtests.txt; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_conv_gemm.py 2>&1 | tail -6 > gpurun_out/pair_alltests.txt; cat gpurun_out/pair_alltests.txt
timeout 300 python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_pair.json > gpurun_out/ops_b32_pair.txt 2>&1; head -30 gpurun_out/ops_b32_pair.txt
HAVC_B200_PAIR=0 timeout 300 python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_nopair.json > gpurun_out/ops_b32_nopair.txt 2>&1; head -12 gpurun_out/ops_b32_nopair.txt
timeout 300 python bench.py --batch 32 --steps 8 --cpu-frames 0 > gpurun_out/bench_b32_pair.json 2> gpurun_out/bench_b32_pair.err; cut -c1-220 gpurun_out/bench_b32_pair.json
HAVC_B200_PAIR=0 timeout 300 python bench.py --batch 32 --steps 8 --cpu-frames 0 > gpurun_out/bench_b32_nopair.json 2> gpurun_out/bench_b32_nopair.err; cut -c1-220 gpurun_out/bench_b32_nopair.json
