#!/bin/bash
# 1 GPU: multi-GPU C-ABI tests with duplicated device ids, kernel + ffi tests of the rebuilt library, launch lists + ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2c7_multi.log 2>&1; echo "multi rc=$?" > gpurun_out/r2c7_summary.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_ffi.py tests/test_gpu_scan.py -q > gpurun_out/r2c7_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c7_summary.txt
# launch lists of C2 / C3 / C5 (one micro-batch set each): no HMMA kernel, no layernorm_kernel expected
for c in c2 c3 c5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c7_launches_$c.csv python scripts/prof_workload.py $c 1 > gpurun_out/r2c7_launches_$c.log 2>&1
done
# full captures: the H=768 GEMM+LN pair kernel (C2), the exact scan kernel, the 8-query filter scan
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_ln_kernel -s 2 -c 2 -o gpurun_out/r2c7_prof_ln768 -f python scripts/prof_workload.py c2 1 > gpurun_out/r2c7_prof_ln768.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk_kernel -s 1 -c 1 -o gpurun_out/r2c7_prof_scan_exact -f python scripts/prof_workload.py scan_exact > gpurun_out/r2c7_prof_scan_exact.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_gemm_kernel -s 2 -c 2 -o gpurun_out/r2c7_prof_scan_gemm8 -f python scripts/prof_workload.py scan_gemm8 > gpurun_out/r2c7_prof_scan_gemm8.log 2>&1
tail -15 gpurun_out/r2c7_multi.log; tail -5 gpurun_out/r2c7_tests.log; cat gpurun_out/r2c7_summary.txt
