#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ffi.py -q -x > gpurun_out/r2c13_tests.log 2>&1; echo "ffi rc=$?" > gpurun_out/r2c13_summary.txt
tail -40 gpurun_out/r2c13_tests.log; cat gpurun_out/r2c13_summary.txt
