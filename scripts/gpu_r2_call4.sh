#!/bin/bash
# full GPU suite + default bench of the current build
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2c4_gpu.txt
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2c4_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2c4_summary.txt
timeout 900 python bench.py > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err; echo "bench rc=$?" >> gpurun_out/r2c4_summary.txt
tail -5 gpurun_out/r2c4_tests.log; cat gpurun_out/r2c4_summary.txt
