#!/bin/bash
# ncu source-level captures of epilogue-bound launches: shuf8.conv (PixelShuffle 1x1, K = 256) and three encoder launches.
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 119 --launch-count 1 -f -o gpurun_out/r9_shuf8 python bench.py --steps 1 --warmup 3 --cpu-frames 0 --no-graph > gpurun_out/r9_ncu_shuf8.log 2>&1; tail -2 gpurun_out/r9_ncu_shuf8.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 46 --launch-count 3 -f -o gpurun_out/r9_enc python bench.py --steps 1 --warmup 3 --cpu-frames 0 --no-graph > gpurun_out/r9_ncu_enc.log 2>&1; tail -2 gpurun_out/r9_ncu_enc.log | cut -c1-200
timeout 300 python bench.py --batch 32 --steps 8 --cpu-frames 0 > gpurun_out/bench_b32_r9.json 2> gpurun_out/bench_b32_r9.err; cut -c1-200 gpurun_out/bench_b32_r9.json
ls -la gpurun_out/r9_*.ncu-rep
