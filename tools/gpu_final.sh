#!/bin/bash
# round-end evidence: full -m gpu suite, smoke, default bench, reference arm
mkdir -p gpurun_out
( time timeout 3000 python -m pytest tests/ -q -m gpu ) > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest_gpu.txt
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02_smoke.txt 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r02_smoke.txt
( time python bench.py ) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_default.err
( time python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_default.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","clocks","parity","tensor_frac_whole_step","gpu_launches"): print(k, d.get(k))
print("roofline", {k:d["roofline"][k] for k in ("achieved","frac","traffic")})
print("pixel", {k:d["roofline_pixel"][k] for k in ("ms","frac_of_hbm","share_of_step")})
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["net_only"], d["cpu_baseline"]["cores"]); print("plugin", d["plugin_surface"])
for k,v in (d.get("arms") or {}).items(): print(k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("value","ms_per_step","tensor_frac_whole_step","error")}, v.get("parity",{}).get("mean_de00") if isinstance(v.get("parity"),dict) else None)
r=json.loads(open("gpurun_out/r02_bench_reference.json").read().strip().splitlines()[-1]); print("reference", r["value"], r["cpu_baseline"]["cores"])
PY
