#!/bin/bash
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests/test_gpu_fullsize.py -q -s ) > gpurun_out/r2d_fullsize.txt 2>&1; echo "fullsize rc=$?"
grep -E "passed|failed|xfail|Error|parity|cfg1:|cfg4 method|real" gpurun_out/r2d_fullsize.txt | cut -c1-260 | tail -24
( time timeout 1500 python bench.py ) > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2d_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2d_bench.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","e2e","clocks","parity","roofline_pixel","cpu_baseline","plugin_surface","tensor_frac_whole_step"):
        print(k, d.get(k))
    print("roofline", {k:d["roofline"][k] for k in ("achieved","frac")})
    for k,v in (d.get("arms") or {}).items(): print(k, v)
except Exception as e:
    print("bench parse failed", e)
PY
