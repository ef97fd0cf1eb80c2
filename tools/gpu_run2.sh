#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_isolated.sh tests > gpurun_out/isolated_summary.txt 2>&1
tail -4 gpurun_out/isolated_summary.txt; grep FAIL gpurun_out/isolated_summary.txt
python bench.py --batch 32 --steps 8 --cpu-frames 0 > gpurun_out/bench_b32_r2.json 2> gpurun_out/bench_b32_r2.err
cut -c1-300 gpurun_out/bench_b32_r2.json
python tools/bench_conv.py --cases enc_l1_c3_b16,enc_l1_c1_b16,enc_l2_c3_b16,enc_l3_c3_b32,enc_l3_c1_b32,enc_l3_c2_b32 2>&1 | tee gpurun_out/conv_r2.txt
