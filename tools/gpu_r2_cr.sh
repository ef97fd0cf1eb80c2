#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_surface.py -x -q -m gpu -k "main or resize_min" 2>&1 | tail -15 > gpurun_out/r2cr_tests.txt; echo "tests rc=$?"; cat gpurun_out/r2cr_tests.txt
