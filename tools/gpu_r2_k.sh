#!/bin/bash
python tools/bench_sharded.py --gpus 2 --frames 1536 --partition block 2>/dev/null | tail -1
python tools/bench_sharded.py --gpus 2 --frames 1536 --partition interleaved --out gpurun_out/r2k_sharded_2gpu.json 2>/dev/null | tail -1
