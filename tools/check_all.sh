#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r15_tests.txt; cat gpurun_out/r15_tests.txt
timeout 300 python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_r15.json > gpurun_out/ops_b32_r15.txt 2>&1; head -12 gpurun_out/ops_b32_r15.txt; tail -7 gpurun_out/ops_b32_r15.txt
timeout 300 python bench.py --batch 32 --steps 10 --cpu-frames 0 > gpurun_out/bench_b32_r15.json 2> gpurun_out/bench_b32_r15.err; cut -c1-170 gpurun_out/bench_b32_r15.json
