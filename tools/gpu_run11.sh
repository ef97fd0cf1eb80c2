#!/bin/bash
# Batch-size sweep of the bench (frames per CUDA-graph step).
mkdir -p gpurun_out
for b in 32 48 64; do
  timeout 400 python bench.py --batch $b --steps 8 --cpu-frames 0 > gpurun_out/bench_b${b}_r11.json 2> gpurun_out/bench_b${b}_r11.err; cut -c1-160 gpurun_out/bench_b${b}_r11.json; python -c "
import json; d=json.load(open('gpurun_out/bench_b${b}_r11.json')); print('  e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'whole', d.get('tensor_frac_whole_step'), d['clocks'])"
done
