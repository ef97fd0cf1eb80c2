#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_surface.py tests/test_gpu_unet.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r16_tests.txt; cat gpurun_out/r16_tests.txt
timeout 400 python bench.py --steps 10 --cpu-frames 0 > gpurun_out/bench_b32_r16.json 2> gpurun_out/bench_b32_r16.err; cut -c1-170 gpurun_out/bench_b32_r16.json; tail -3 gpurun_out/bench_b32_r16.err
python -c "
import json; d=json.load(open('gpurun_out/bench_b32_r16.json')); print('e2e', d['e2e']['value'], 'plugin', d['plugin_surface'])"
