#!/bin/bash
mkdir -p gpurun_out
TESTS=test_gpu_scan bash scripts/gpu_tests.sh
(timeout 300 python scripts/scan_time.py 6250000; for d in 1 5; do echo dbg=$d; KJC_SG_DBG=$d NQ=4096 timeout 300 python scripts/scan_time.py 6250000; done) 2>&1 | tee gpurun_out/sg_time.txt
