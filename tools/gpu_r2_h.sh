#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_fullsize.py -x -q -k "not cfg4 and not cfg1" > gpurun_out/r2h_tests.txt 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2h_tests.txt
for b in 32 48 64; do
  timeout 600 python bench.py --batch $b --steps 10 --cpu-frames 0 --plugin-frames 0 --extras "" > gpurun_out/r2h_bench_b$b.json 2> gpurun_out/r2h_bench_b$b.err; echo "bench b=$b rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2h_bench_b$b.json").read().strip().splitlines()[-1])
    print("B=$b", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), "gemm", round(d["breakdown"]["gemm_ms_per_step"],2), "aux", round(d["breakdown"]["aux_ms_per_step"],2), "px", round(d["roofline_pixel"]["ms"],3), d["clocks"]["sm_mhz"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2h_bench_b$b.err").read()[-1500:])
PY
done
HAVC_B200_PRECISION=fast timeout 600 python tools/profile_ops.py --batch 32 --out gpurun_out/r2h_ops_fast.json > gpurun_out/r2h_ops_fast.txt 2>&1; head -1 gpurun_out/r2h_ops_fast.txt; grep -E "conv.3|shuf" gpurun_out/r2h_ops_fast.txt | head -12
