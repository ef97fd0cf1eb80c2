#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"; nproc
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","e2e","shard_check","tensor_frac_whole_step")})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r02_bench_${N}gpu.err").read()[-1500:])
PY
timeout 900 python tools/bench_sharded.py --gpus $N --frames $((384*N)) --out gpurun_out/r02_sharded_inprocess_${N}gpu.json 2> gpurun_out/r02_sharded_${N}.err | tail -1 | cut -c1-600; echo "sharded rc=$?"
