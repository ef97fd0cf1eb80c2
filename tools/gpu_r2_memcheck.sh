#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels' tests (periodic pixel passes, stem im2col, temporal stabiliser, zimg, resize)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_pixel.py -x -q -m gpu > gpurun_out/r02_memcheck_pixel.txt 2>&1; echo "memcheck pixel rc=$?"
tail -6 gpurun_out/r02_memcheck_pixel.txt
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_surface.py -x -q -m gpu -k "temporal_stabilizer_bit_exact or merge_and_temporal or zimg" > gpurun_out/r02_memcheck_temporal.txt 2>&1; echo "memcheck temporal rc=$?"
tail -6 gpurun_out/r02_memcheck_temporal.txt
