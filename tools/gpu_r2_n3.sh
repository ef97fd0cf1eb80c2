#!/bin/bash
# round 2: temporal stabiliser (N3) + fp16 saturation guard on the GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_surface.py tests/test_gpu_filters.py tests/test_gpu_conv_gemm.py -x -q -m gpu \
  -k "temporal or stabilizer or restore or crt or retention or saturate or combine" 2>&1 | tail -15 > gpurun_out/r2n3_tests.txt
echo "tests rc=$?"; cat gpurun_out/r2n3_tests.txt
