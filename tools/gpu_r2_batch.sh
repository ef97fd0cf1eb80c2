#!/bin/bash
mkdir -p gpurun_out
for b in 32 64 48 32; do
  timeout 600 python bench.py --batch $b --steps 16 --warmup 3 --extras "" --cpu-frames 0 --plugin-frames 0 > gpurun_out/r2batch_$b.json 2> gpurun_out/r2batch_$b.err; echo "b=$b rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2batch_$b.json").read().strip().splitlines()[-1])
print($b, round(d["value"],1), round(d["ms_per_step"],3), round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],4))
PY
done
