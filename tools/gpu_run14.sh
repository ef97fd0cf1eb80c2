#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r14_tests.txt; cat gpurun_out/r14_tests.txt
timeout 300 python bench.py --batch 32 --steps 10 --cpu-frames 0 > gpurun_out/bench_b32_r14.json 2> gpurun_out/bench_b32_r14.err; cut -c1-170 gpurun_out/bench_b32_r14.json
HAVC_B200_NO_PDL=1 timeout 300 python bench.py --batch 32 --steps 10 --cpu-frames 0 > gpurun_out/bench_b32_r14_nopdl.json 2> gpurun_out/bench_b32_r14_nopdl.err; cut -c1-170 gpurun_out/bench_b32_r14_nopdl.json
timeout 300 python bench.py --batch 32 --steps 10 --cpu-frames 0 > gpurun_out/bench_b32_r14b.json 2> gpurun_out/bench_b32_r14b.err; cut -c1-170 gpurun_out/bench_b32_r14b.json
python - <<'PY'
import json
for f in ("bench_b32_r14","bench_b32_r14_nopdl","bench_b32_r14b"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], "gemm frac", round(d["roofline"]["frac"],3))
PY
