#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_scan.py tests/test_gpu_multi.py tests/test_gpu_index_dir.py -q -x > gpurun_out/r2c14_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2c14_summary.txt
{ echo "== t8 exact scan, 6.25M x 384"; KJC_SCAN_GEMM_MIN_Q=100000 NQ=1,8,16 timeout 300 python scripts/scan_time.py 6250000
  echo "== old kernels"; KJC_SCAN_NO_T8=1 KJC_SCAN_GEMM_MIN_Q=100000 NQ=8 timeout 300 python scripts/scan_time.py 6250000
  echo "== filter path (rescore in the new order)"; NQ=8,4096 timeout 300 python scripts/scan_time.py 6250000
  echo "== dim 768"; DIM=768 K=10 NQ=8 timeout 300 python scripts/scan_time2.py 2000000
} > gpurun_out/r2c14_scan_times.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_t8_kernel -s 1 -c 1 -o gpurun_out/r2c14_prof_scan_t8 -f python scripts/prof_workload.py scan_exact > gpurun_out/r2c14_prof.log 2>&1
tail -30 gpurun_out/r2c14_tests.log; cat gpurun_out/r2c14_scan_times.txt; cat gpurun_out/r2c14_summary.txt
