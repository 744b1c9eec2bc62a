#!/bin/bash
# Runs the GPU parity tests file by file (each under its own timeout) and keeps the logs in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in ${TESTS:-test_gpu_kernels test_gpu_encoder test_gpu_scan}; do
  timeout ${TEST_TIMEOUT:-400} python -m pytest tests/$f.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/$f.log 2>&1
  r=$?; echo "$f exit=$r" >> gpurun_out/summary.txt; [ $r -ne 0 ] && rc=$r
  tail -25 gpurun_out/$f.log
done
exit $rc
