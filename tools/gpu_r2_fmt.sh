#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_surface.py -x -q -m gpu -k "yuv or gray or temporal or stabilizer or merge" 2>&1 | tail -12 > gpurun_out/r2fmt_tests.txt; echo "tests rc=$?"; cat gpurun_out/r2fmt_tests.txt
