#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 900 python tools/bench_sharded.py --gpus $N --frames $((640*N)) --out gpurun_out/r02_sharded_inprocess_${N}gpu.json 2> gpurun_out/r02_sharded_${N}.err | tail -1 | cut -c1-900; echo "sharded rc=$?"
tail -3 gpurun_out/r02_sharded_${N}.err
