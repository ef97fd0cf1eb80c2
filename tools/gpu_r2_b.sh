#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_gemm.py -x -q -k x3 > gpurun_out/r2b_conv.txt 2>&1; echo "conv rc=$?"; tail -3 gpurun_out/r2b_conv.txt
timeout 900 python -m pytest tests/test_gpu_unet.py -x -q > gpurun_out/r2b_unet.txt 2>&1; echo "unet rc=$?"; tail -5 gpurun_out/r2b_unet.txt
for prec in fast balanced; do
  HAVC_B200_PRECISION=$prec timeout 600 python tools/profile_ops.py --batch 32 --out gpurun_out/r2b_ops_$prec.json > gpurun_out/r2b_ops_$prec.txt 2>&1; echo "ops $prec rc=$?"
  head -1 gpurun_out/r2b_ops_$prec.txt; tail -8 gpurun_out/r2b_ops_$prec.txt
done
