#!/bin/bash
mkdir -p gpurun_out
( time timeout 3000 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.txt
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.txt
