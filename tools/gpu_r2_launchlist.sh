#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_b32.csv python bench.py --steps 2 --warmup 3 --preheat 0 --cpu-frames 0 --plugin-frames 0 --extras "" > gpurun_out/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
python tools/launch_shares.py gpurun_out/r02_launches_b32.csv gpurun_out/r02_launch_shares.txt | head -30
