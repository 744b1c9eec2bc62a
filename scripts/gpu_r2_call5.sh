#!/bin/bash
# H=768 fused GEMM+LN (CTA pair): kernel parity, encoder/ffi suites, config bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "fused_gemm_residual" > gpurun_out/r2c5_ln768.log 2>&1; echo "ln768 rc=$?" > gpurun_out/r2c5_summary.txt
timeout 1500 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_ffi.py -q > gpurun_out/r2c5_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c5_summary.txt
timeout 900 python bench.py --no-index --no-cpu > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err; echo "bench rc=$?" >> gpurun_out/r2c5_summary.txt
tail -5 gpurun_out/r2c5_ln768.log; tail -8 gpurun_out/r2c5_tests.log; cat gpurun_out/r2c5_summary.txt
