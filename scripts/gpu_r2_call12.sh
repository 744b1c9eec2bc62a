#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_scan.py tests/test_gpu_multi.py tests/test_gpu_index_dir.py tests/test_gpu_ffi.py -q > gpurun_out/r2c12_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2c12_summary.txt
{ for cfg in "384 10" "384 50" "768 10" "768 50" "1024 10"; do set -- $cfg; echo "== dim $1 k $2"; DIM=$1 K=$2 NQ=8,256,4096 timeout 300 python scripts/scan_time2.py 2000000; done; } > gpurun_out/r2c12_scan_times.txt 2>&1
tail -30 gpurun_out/r2c12_tests.log; cat gpurun_out/r2c12_scan_times.txt; cat gpurun_out/r2c12_summary.txt
