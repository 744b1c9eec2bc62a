#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_index_dir.py tests/test_gpu_ffi.py -q > gpurun_out/r2c8_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2c8_summary.txt
tail -40 gpurun_out/r2c8_tests.log; cat gpurun_out/r2c8_summary.txt
