#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 1500 python -m pytest tests/test_gpu_unet.py tests/test_gpu_surface.py tests/test_gpu_filters.py -x -q > gpurun_out/r2i_tests.txt 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2i_tests.txt
N=$(nvidia-smi -L | wc -l)
if [ "$N" -ge 2 ]; then
  timeout 900 python tools/bench_sharded.py --gpus $N --frames $((512*N)) --out gpurun_out/r2i_sharded_${N}gpu.json 2> gpurun_out/r2i_sharded.err | tail -1; echo "sharded rc=$?"; tail -2 gpurun_out/r2i_sharded.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2i_bench_${N}gpu.json 2> gpurun_out/r2i_bench_${N}gpu.err; echo "bench$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2i_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","e2e","shard_check","tensor_frac_whole_step")})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r2i_bench_${N}gpu.err").read()[-1500:])
PY
fi
