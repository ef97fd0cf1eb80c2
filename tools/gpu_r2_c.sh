#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_zhang.py tests/test_gpu_unet.py tests/test_gpu_surface.py -q -s > gpurun_out/r2c_tests.txt 2>&1; echo "tests rc=$?"
grep -E "passed|failed|xfail|Error" gpurun_out/r2c_tests.txt | tail -8
grep -E "^(eccv16|siggraph17) (256|96) " gpurun_out/r2c_tests.txt | cut -c1-200
HAVC_B200_PAIR=1 HAVC_B200_PRECISION=balanced timeout 600 python tools/profile_ops.py --batch 32 --out gpurun_out/r2c_ops_balanced_pair.json > gpurun_out/r2c_ops_balanced_pair.txt 2>&1; echo "ops rc=$?"
head -1 gpurun_out/r2c_ops_balanced_pair.txt
