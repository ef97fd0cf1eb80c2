#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_filters.py tests/test_gpu_surface.py -x -q -k "vs_tweak or mask_saturation or sat_hue or merge_methods or combine" > gpurun_out/r2l_tests.txt 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2l_tests.txt | cut -c1-300
