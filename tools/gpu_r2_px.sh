#!/bin/bash
# round 2: phase-periodic horizontal pixel passes: bit-exactness against the table kernels, the pipeline tests that run through them,
# per-launch timings, a short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pixel.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2px_tests.txt; echo "pixel tests rc=$?"; cat gpurun_out/r2px_tests.txt
timeout 900 python -m pytest tests/test_gpu_surface.py tests/test_gpu_fullsize.py tests/test_gpu_filters.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2px_tests2.txt; echo "pipeline tests rc=$?"; cat gpurun_out/r2px_tests2.txt
timeout 600 python tools/profile_ops.py --batch 32 --out gpurun_out/r2px_ops.json > gpurun_out/r2px_ops.txt 2>&1; echo "ops rc=$?"
head -1 gpurun_out/r2px_ops.txt; tail -9 gpurun_out/r2px_ops.txt
timeout 600 python bench.py --steps 20 --warmup 3 --extras "" --cpu-frames 0 --plugin-frames 0 > gpurun_out/r2px_bench.json 2> gpurun_out/r2px_bench.err; echo "bench rc=$?"
cut -c1-700 gpurun_out/r2px_bench.json
