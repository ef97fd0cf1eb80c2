#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pixel.py tests/test_gpu_unet.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2stem_tests.txt; echo "tests rc=$?"; cat gpurun_out/r2stem_tests.txt
timeout 600 python tools/profile_ops.py --batch 32 --out gpurun_out/r2stem_ops.json > gpurun_out/r2stem_ops.txt 2>&1; echo "ops rc=$?"
head -1 gpurun_out/r2stem_ops.txt; grep -E "stem|full batch" gpurun_out/r2stem_ops.txt | cut -c1-150
