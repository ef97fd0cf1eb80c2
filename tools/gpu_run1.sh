#!/bin/bash
# round-1 session-2 GPU check: tests, balanced-split A/B, epilogue-bound conv capture
mkdir -p gpurun_out
bash tools/gpu_isolated.sh tests > gpurun_out/isolated_summary.txt 2>&1
tail -3 gpurun_out/isolated_summary.txt
python tools/bench_conv.py --cases res0_272,res1_272,enc_l1_c3_b16,enc_l1_c1_b16,enc_l2_c3_b16,enc_l3_c3_b32,enc_l3_c1_b32,enc_l3_c2_b32 > gpurun_out/conv_new.txt 2>&1
HAVC_B200_SPLIT256=1 python tools/bench_conv.py --cases res0_272,res1_272 > gpurun_out/conv_legacy.txt 2>&1
cat gpurun_out/conv_new.txt gpurun_out/conv_legacy.txt
python bench.py --batch 32 --steps 8 --cpu-frames 0 > gpurun_out/bench_b32_split.json 2> gpurun_out/bench_b32_split.err
HAVC_B200_SPLIT256=1 python bench.py --batch 32 --steps 8 --cpu-frames 0 > gpurun_out/bench_b32_legacy.json 2>> gpurun_out/bench_b32_split.err
cut -c1-400 gpurun_out/bench_b32_split.json gpurun_out/bench_b32_legacy.json
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm -s 2 -c 1 -f -o gpurun_out/l1c3 python tools/bench_conv.py --cases enc_l1_c3_b16 --reps 1 > gpurun_out/ncu_l1c3.log 2>&1
tail -2 gpurun_out/ncu_l1c3.log
