#!/bin/bash
mkdir -p gpurun_out
{ for cfg in "30000 1024 9 50" "30000 1024 8 10" "30000 512 9 50" "300000 768 9 50" "300000 384 17 100" "30000 64 5 10" "30000 32 3 256"; do echo "== $cfg"; timeout 120 python scripts/scan_repro.py $cfg 2>&1 | tail -1; done; } > gpurun_out/r2c15_repro.txt 2>&1
cat gpurun_out/r2c15_repro.txt
bash scripts/gpu_r2_call14.sh
