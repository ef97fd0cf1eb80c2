#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r17_tests.txt; cat gpurun_out/r17_tests.txt
timeout 300 python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_r17.json > gpurun_out/ops_b32_r17.txt 2>&1; head -3 gpurun_out/ops_b32_r17.txt; grep -E "blur|skip_bn|full batch" gpurun_out/ops_b32_r17.txt
timeout 300 python bench.py --batch 32 --steps 10 --cpu-frames 0 --plugin-frames 0 > gpurun_out/bench_b32_r17.json 2> gpurun_out/bench_b32_r17.err; cut -c1-170 gpurun_out/bench_b32_r17.json
