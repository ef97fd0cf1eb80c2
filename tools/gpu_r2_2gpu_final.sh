#!/bin/bash
# round 2, final code on N GPUs of one box: the multi-GPU product-path test and the driver's torchrun bench command
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 900 python -m pytest tests/test_gpu_surface.py -x -q -m gpu -k "multi_gpu" 2>&1 | tail -4 > gpurun_out/r02_multigpu_test_${N}.txt; echo "test rc=$?"; cat gpurun_out/r02_multigpu_test_${N}.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 --extras "" > gpurun_out/r02_bench_${N}gpu_final.json 2> gpurun_out/r02_bench_${N}gpu_final.err; echo "bench$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_${N}gpu_final.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","e2e","shard_check","clocks")})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r02_bench_${N}gpu_final.err").read()[-1500:])
PY
