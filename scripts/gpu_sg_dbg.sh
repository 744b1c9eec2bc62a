#!/bin/bash
mkdir -p gpurun_out
TESTS=test_gpu_scan bash scripts/gpu_tests.sh
(echo seeded; python scripts/scan_time.py 6250000; echo unseeded; KJC_SG_NO_SEED=1 NQ=128,4096 python scripts/scan_time.py 6250000) 2>&1 | tee gpurun_out/sg_time.txt
