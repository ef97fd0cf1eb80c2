#!/bin/bash
# round 2, step A: split-precision conv tests, U-Net parity per layer, fast vs balanced bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py -x -q > gpurun_out/r2a_conv.txt 2>&1; echo "conv rc=$?" 
tail -5 gpurun_out/r2a_conv.txt
timeout 900 python -m pytest tests/test_gpu_unet.py -x -q -s > gpurun_out/r2a_unet.txt 2>&1; echo "unet rc=$?"
tail -30 gpurun_out/r2a_unet.txt
for prec in fast balanced; do
  HAVC_B200_PRECISION=$prec timeout 600 python bench.py --steps 10 --cpu-frames 0 --plugin-frames 0 > gpurun_out/r2a_bench_$prec.json 2> gpurun_out/r2a_bench_$prec.err; echo "bench $prec rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2a_bench_$prec.json"))
    print("$prec", round(d["value"],1), "fps", round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["frac"],3), d["breakdown"]["gemm_ms_per_step"], d["breakdown"]["aux_ms_per_step"])
except Exception as e:
    print("bench $prec failed", e)
    print(open("gpurun_out/r2a_bench_$prec.err").read()[-2000:])
PY
done
