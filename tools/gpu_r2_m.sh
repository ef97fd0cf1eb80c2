#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_surface.py tests/test_gpu_filters.py -x -q -k "yuv420p8 or zimg or vs_tweak or sat_hue or order_props" > gpurun_out/r2m_tests.txt 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2m_tests.txt | cut -c1-300
