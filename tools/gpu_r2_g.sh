#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py -x -q -k "blur or shuffle" > gpurun_out/r2g_conv.txt 2>&1; echo "conv rc=$?"; tail -6 gpurun_out/r2g_conv.txt
timeout 900 python -m pytest tests/test_gpu_unet.py -x -q > gpurun_out/r2g_unet.txt 2>&1; echo "unet rc=$?"; tail -4 gpurun_out/r2g_unet.txt
for fb in 0 1; do
  HAVC_B200_FUSE_BLUR=$fb timeout 600 python bench.py --steps 10 --cpu-frames 0 --plugin-frames 0 --extras "" > gpurun_out/r2g_bench_fb$fb.json 2> gpurun_out/r2g_bench_fb$fb.err; echo "bench fb=$fb rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2g_bench_fb$fb.json").read().strip().splitlines()[-1])
    print("fuse_blur=$fb", round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms/step", d["breakdown"]["gemm_ms_per_step"], d["breakdown"]["aux_ms_per_step"], [ (t["op"], t["ms"]) for t in d["breakdown"]["top"][:8]])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2g_bench_fb$fb.err").read()[-1500:])
PY
done
