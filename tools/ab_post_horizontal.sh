#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tools/debug_md5.py | tail -1; HAVC_B200_LEGACY_PIXEL=1 timeout 300 python tools/debug_md5.py | tail -1 ) 2>&1 | grep md5
timeout 300 python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_r19.json 2>&1 | tail -6
HAVC_B200_LEGACY_PIXEL=1 timeout 300 python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_r19l.json 2>&1 | tail -6
timeout 600 ncu --set full --import-source on --clock-control none -k regex:post_horizontal --launch-skip 1 --launch-count 1 -f -o gpurun_out/r19_posth python bench.py --steps 1 --warmup 3 --cpu-frames 0 --plugin-frames 0 --no-graph > gpurun_out/r19_ncu_posth.log 2>&1; tail -1 gpurun_out/r19_ncu_posth.log | cut -c1-150
