#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_ffi.py -x -q > gpurun_out/r2c3_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2c3_summary.txt
timeout 900 python bench.py --no-index --no-cpu > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err; echo "bench rc=$?" >> gpurun_out/r2c3_summary.txt
KJC_BN_QKV=192 timeout 900 python bench.py --no-index --no-cpu > gpurun_out/r2c3_bench_qkv192.json 2> gpurun_out/r2c3_bench_qkv192.err
tail -5 gpurun_out/r2c3_tests.log; cat gpurun_out/r2c3_summary.txt
