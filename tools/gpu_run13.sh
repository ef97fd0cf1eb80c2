#!/bin/bash
mkdir -p gpurun_out
( echo "== default"; timeout 300 python tools/debug_perm.py
  echo "== B=4"; DBG_B=4 timeout 300 python tools/debug_perm.py
) > gpurun_out/r13_perm.txt 2>&1
cat gpurun_out/r13_perm.txt | grep -v Warning | tail -30
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -s 2>&1 | tail -15 > gpurun_out/r12_fullsize.txt; cat gpurun_out/r12_fullsize.txt
