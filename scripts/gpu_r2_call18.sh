#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_scan.py tests/test_gpu_multi.py tests/test_gpu_index_dir.py -q -x > gpurun_out/r2c18_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2c18_summary.txt
{ for cfg in "30000 1024 9 50" "300000 768 9 50" "300000 384 17 100" "30000 64 5 10" "30000 32 3 256" "100 384 8 10"; do echo "== $cfg"; timeout 120 python scripts/scan_repro.py $cfg 2>&1 | tail -1; done; } > gpurun_out/r2c18_repro.txt 2>&1
{ python scripts/scan_time3.py; DIM=768 K=10 NQ=8 timeout 300 python scripts/scan_time2.py 2000000; } > gpurun_out/r2c18_times.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_t8_kernel -s 1 -c 1 -o gpurun_out/r2c18_prof_scan_t8 -f python scripts/prof_workload.py scan_exact > gpurun_out/r2c18_prof.log 2>&1
tail -5 gpurun_out/r2c18_tests.log; cat gpurun_out/r2c18_repro.txt gpurun_out/r2c18_times.txt gpurun_out/r2c18_summary.txt
