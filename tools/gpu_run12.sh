#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -s 2>&1 | tail -15 > gpurun_out/r12_fullsize.txt; cat gpurun_out/r12_fullsize.txt
