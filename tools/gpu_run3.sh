#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_isolated.sh tests > gpurun_out/isolated_summary.txt 2>&1
tail -3 gpurun_out/isolated_summary.txt; grep FAIL gpurun_out/isolated_summary.txt
CASES=enc_l1_c3_b16,enc_l1_c1_b16,enc_l2_c3_b16,enc_l3_c3_b32,enc_l3_c1_b32,enc_l3_c2_b32,shuf8,c256_k256,l6conv
python tools/bench_conv.py --cases $CASES 2>&1 | tee gpurun_out/conv_fast.txt
HAVC_B200_NO_FAST_EPILOGUE=1 python tools/bench_conv.py --cases $CASES 2>&1 | tee gpurun_out/conv_nofast.txt
python bench.py --batch 32 --steps 8 --cpu-frames 0 > gpurun_out/bench_b32_fast.json 2> gpurun_out/bench_b32_fast.err
cut -c1-300 gpurun_out/bench_b32_fast.json
python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_fast.json > gpurun_out/ops_b32_fast.txt 2>&1; head -3 gpurun_out/ops_b32_fast.txt
