#!/bin/bash
# A/B: suspend-time hint on the long mbarrier waits (power / issue-slot effect of spinning warps); each arm twice, interleaved.
mkdir -p gpurun_out
for arm in 0 2000 0 2000 20000; do
  HAVC_B200_WAIT_HINT=$arm timeout 300 python bench.py --batch 32 --steps 10 --cpu-frames 0 --plugin-frames 0 > gpurun_out/bench_hint_$arm.json 2> gpurun_out/bench_hint_$arm.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_hint_$arm.json')); print('hint $arm', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], 'gemm frac', round(d['roofline']['frac'],3))"
done
