#!/bin/bash
# Last run of the round on the final code: -m gpu tests, default bench, ncu launch list (the ncu --set full captures of the
# conv kernel are those of tools/gpu_measure.sh: the kernel did not change afterwards).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-200 gpurun_out/bench_default.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_b32.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0 --plugin-frames 0 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-120
timeout 200 python tools/profile_ops.py --batch 32 --out gpurun_out/ops_b32_final.json > gpurun_out/ops_b32_final.txt 2>&1; tail -7 gpurun_out/ops_b32_final.txt
