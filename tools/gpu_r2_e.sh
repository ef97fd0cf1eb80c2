#!/bin/bash
# 2-GPU box: multi-GPU product path + bench under torchrun (sharded single clip) + new surface tests
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_surface.py tests/test_gpu_filters.py -q -x -k "multi_gpu or clip_luma or chroma_resize or merge" > gpurun_out/r2e_tests.txt 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r2e_tests.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2e_bench_2gpu.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","e2e","shard_check","tensor_frac_whole_step")})
except Exception as e:
    print("bench2 parse failed", e); print(open("gpurun_out/r2e_bench_2gpu.err").read()[-1500:])
PY
timeout 900 python tools/bench_sharded.py --gpus 2 --frames 768 --out gpurun_out/r2e_sharded_2gpu.json 2> gpurun_out/r2e_sharded.err | tail -1; echo "sharded rc=$?"
tail -3 gpurun_out/r2e_sharded.err
