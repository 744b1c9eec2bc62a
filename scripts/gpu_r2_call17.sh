#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-extra --no-cpu > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err
{ KJC_SCAN_GEMM_MIN_Q=100000 NQ=8 timeout 300 python scripts/scan_time.py 6250000; nvidia-smi --query-gpu=clocks.sm,clocks.mem,temperature.gpu,temperature.memory,power.draw,clocks_event_reasons.active --format=csv; } > gpurun_out/r2c17_scan.txt 2>&1
python -c "
import json
d=json.load(open('gpurun_out/r2c17_bench.json')); it=d['index_topk']
print(it['exact_scan_8q']['ms_per_step'], it['exact_scan_8q']['roofline']['frac'], it['small_batch_8q']['ms_per_step'], it['ms_per_step'])"
cat gpurun_out/r2c17_scan.txt
