#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_scan.py tests/test_gpu_multi.py tests/test_gpu_index_dir.py tests/test_gpu_encoder.py -q > gpurun_out/r2c9_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2c9_summary.txt
timeout 900 python bench.py > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err; echo "bench rc=$?" >> gpurun_out/r2c9_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_topk_kernel -s 1 -c 1 -o gpurun_out/r2c9_prof_scan_q8 -f python scripts/prof_workload.py scan_exact > gpurun_out/r2c9_prof_scan_q8.log 2>&1
tail -30 gpurun_out/r2c9_tests.log; cat gpurun_out/r2c9_summary.txt; tail -3 gpurun_out/r2c9_bench.err
