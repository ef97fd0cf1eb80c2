#!/bin/bash
# round 2: ICNR blur in packed 16-bit arithmetic (fused epilogue + havc_blur2x2): tests, per-launch table, bench with parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_unet.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2b16_tests.txt; echo "tests rc=$?"; cat gpurun_out/r2b16_tests.txt
timeout 600 python tools/profile_ops.py --batch 32 --out gpurun_out/r2b16_ops.json > gpurun_out/r2b16_ops.txt 2>&1; echo "ops rc=$?"
head -12 gpurun_out/r2b16_ops.txt | cut -c1-150; grep -E "shuf|full batch" gpurun_out/r2b16_ops.txt | cut -c1-150
timeout 600 python bench.py --steps 20 --warmup 3 --extras "" --cpu-frames 0 --plugin-frames 0 > gpurun_out/r2b16_bench.json 2> gpurun_out/r2b16_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b16_bench.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","clocks","parity"): print(k, d.get(k))
print("roofline", {k:d["roofline"][k] for k in ("achieved","frac")})
PY
