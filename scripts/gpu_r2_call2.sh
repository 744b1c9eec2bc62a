#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c2_tests.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/r2c2_summary.txt
timeout 900 python bench.py > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err; echo "bench rc=$?" >> gpurun_out/r2c2_summary.txt
tail -3 gpurun_out/r2c2_tests.log; cat gpurun_out/r2c2_summary.txt; tail -5 gpurun_out/r2c2_bench.err
