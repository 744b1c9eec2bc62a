#!/bin/bash
# round-2 call 1: new attention kernel parity + timing A/B, then the full GPU suite and the bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2c1_gpu.txt
timeout 300 python -m pytest tests/test_gpu_kernels.py -k attention -x -q > gpurun_out/r2c1_attn_tests.log 2>&1; echo "attn tests rc=$?" >> gpurun_out/r2c1_summary.txt
for v in tc ts; do
  for shape in "148 128 384 12" "148 128 768 12" "74 256 384 12" "37 512 768 12" "64 64 384 12"; do
    KJC_ATTN=$v timeout 120 python scripts/attn_trace.py $shape 2>&1 | grep "us/launch" | sed "s/^/$v /" >> gpurun_out/r2c1_attn_time.txt
  done
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_tests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/r2c1_summary.txt
timeout 600 python bench.py > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?" >> gpurun_out/r2c1_summary.txt
KJC_ATTN=tc timeout 600 python bench.py > gpurun_out/r2c1_bench_tc.json 2> gpurun_out/r2c1_bench_tc.err
tail -3 gpurun_out/r2c1_attn_tests.log; cat gpurun_out/r2c1_attn_time.txt; tail -3 gpurun_out/r2c1_tests.log; cat gpurun_out/r2c1_summary.txt
