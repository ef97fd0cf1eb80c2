#!/bin/bash
# im2col static-cin A/B, ncu captures of the pair res conv, the multi-row horizontal pass and post_horizontal.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "unet or surface or zhang" 2>&1 | tail -6 > gpurun_out/r7_tests.txt; cat gpurun_out/r7_tests.txt
timeout 300 python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_r7.json > gpurun_out/ops_b32_r7.txt 2>&1; grep -E "im2col|total|pre\.|post\.|full batch" gpurun_out/ops_b32_r7.txt
HAVC_B200_LEGACY_IM2COL=1 timeout 300 python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_r7_legacyim.json > gpurun_out/ops_b32_r7_legacyim.txt 2>&1; grep -E "im2col|total" gpurun_out/ops_b32_r7_legacyim.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 120 --launch-count 1 -f -o gpurun_out/r7_resconv0_pair python bench.py --steps 1 --warmup 3 --cpu-frames 0 --no-graph > gpurun_out/r7_ncu_res.log 2>&1; tail -2 gpurun_out/r7_ncu_res.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:resample_h_rows --launch-skip 1 --launch-count 1 -f -o gpurun_out/r7_preh python bench.py --steps 1 --warmup 3 --cpu-frames 0 --no-graph > gpurun_out/r7_ncu_preh.log 2>&1; tail -2 gpurun_out/r7_ncu_preh.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:im2col_tile --launch-skip 4 --launch-count 1 -f -o gpurun_out/r7_im2col python bench.py --steps 1 --warmup 3 --cpu-frames 0 --no-graph > gpurun_out/r7_ncu_im2col.log 2>&1; tail -2 gpurun_out/r7_ncu_im2col.log | cut -c1-200
ls -la gpurun_out/r7_*.ncu-rep
