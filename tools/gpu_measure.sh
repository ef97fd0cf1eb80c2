#!/bin/bash
# Round measurement run: -m gpu tests, smoke, default bench, reference arm, ncu launch list, ncu --set full of shuf8.conv + res.conv0.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-250 gpurun_out/bench_default.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-250 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_b32.csv python bench.py --steps 2 --warmup 3 --cpu-frames 0 --plugin-frames 0 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 119 --launch-count 2 -f -o gpurun_out/final_shuf8_resconv0 python bench.py --steps 1 --warmup 3 --cpu-frames 0 --plugin-frames 0 --no-graph > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep | tail -2
