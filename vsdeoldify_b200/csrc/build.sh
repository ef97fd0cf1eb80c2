#!/bin/bash
# Builds libhavc_b200.so in-tree for sm_100a (the only target).  Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
OUT=../libhavc_b200.so
SRCS=$(ls *.cu)
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC -shared -Xptxas -v "$@" -o $OUT $SRCS -lcudart 2>&1 | grep -E "error|warning|registers|spill|ptxas info    : Compiling" || true
ls -la $OUT
