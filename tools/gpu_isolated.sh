#!/bin/bash
# Runs every -m gpu test of the given files in its own process (a device-side trap poisons the CUDA
# context of the process that hit it), with a per-test timeout.  Log: gpurun_out/isolated.log
mkdir -p gpurun_out
LOG=gpurun_out/isolated.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $LOG 2>&1
ids=$(python -m pytest "$@" -m gpu --collect-only -q 2>/dev/null | grep "::")
pass=0; fail=0
for id in $ids; do
  echo "=== $id" >> $LOG
  timeout ${PER_TEST_TIMEOUT:-180} python -m pytest "$id" -x -q --no-header -p no:cacheprovider 2>&1 | tail -n 25 >> $LOG
  rc=${PIPESTATUS[0]}
  if [ $rc -eq 0 ]; then pass=$((pass+1)); echo "PASS $id"; else fail=$((fail+1)); echo "FAIL($rc) $id"; fi
done
echo "isolated: $pass passed, $fail failed" | tee -a $LOG
