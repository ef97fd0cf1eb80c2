#!/bin/bash
# Round measurement run: -m gpu tests, default bench, reference arm, ncu launch list, ncu --set full of the dominant launch.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-250 gpurun_out/bench_default.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-250 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_b32.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 120 --launch-count 1 -f -o gpurun_out/resconv0_b32 python bench.py --steps 1 --warmup 3 --cpu-frames 0 --no-graph > gpurun_out/ncu_resconv0.log 2>&1
tail -2 gpurun_out/ncu_resconv0.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:post_horizontal --launch-skip 1 --launch-count 1 -f -o gpurun_out/posth_b32 python bench.py --steps 1 --warmup 3 --cpu-frames 0 --no-graph > gpurun_out/ncu_posth.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
