#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python bench.py --steps 20 --warmup 3 --extras "" --cpu-frames 0 --plugin-frames 0 > gpurun_out/r2bc_1.json 2> gpurun_out/r2bc_1.err; echo "bench1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --extras "" --cpu-frames 0 --plugin-frames 0 > gpurun_out/r2bc_$N.json 2> gpurun_out/r2bc_$N.err; echo "bench$N rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2bc_ref_$N.json 2> gpurun_out/r2bc_ref_$N.err; echo "ref$N rc=$?"
python - <<PY
import json
for f in ("gpurun_out/r2bc_1.json","gpurun_out/r2bc_$N.json","gpurun_out/r2bc_ref_$N.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("impl","value","ms_per_step","n_gpus","e2e","shard_check")})
    except Exception as e:
        print(f, "parse failed", e); print(open(f.replace(".json",".err")).read()[-1200:])
PY
