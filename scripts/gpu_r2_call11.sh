#!/bin/bash
# (a) is the in-situ slowdown of the chained kernels a power/clock effect?  zeros vs random data, short vs long runs
# (b) exact-scan q8 variants
mkdir -p gpurun_out
{
echo "== zeros, 30 iters"; timeout 300 python scripts/chain_micro.py
echo "== random, 30 iters"; RANDOM_DATA=1 timeout 300 python scripts/chain_micro.py
echo "== zeros, 3000 iters"; ITERS=3000 timeout 300 python scripts/chain_micro.py
echo "== random, 3000 iters"; RANDOM_DATA=1 ITERS=3000 timeout 300 python scripts/chain_micro.py
} > gpurun_out/r2c11_chain_power.txt 2>&1
for v in 0 1 2; do
  echo "== q8 variant $v"; KJC_SCAN_Q8_VARIANT=$v KJC_SCAN_GEMM_MIN_Q=100000 NQ=8 timeout 300 python scripts/scan_time.py 6250000
done > gpurun_out/r2c11_scan_variants.txt 2>&1
KJC_SCAN_Q8_VARIANT=1 timeout 600 python -m pytest tests/test_gpu_scan.py -q > gpurun_out/r2c11_scan_v1_tests.log 2>&1; echo "scan v1 tests rc=$?" > gpurun_out/r2c11_summary.txt
timeout 600 python -m pytest tests/test_gpu_encoder.py -q -k "mpnet" > gpurun_out/r2c11_enc.log 2>&1; echo "mpnet rc=$?" >> gpurun_out/r2c11_summary.txt
cat gpurun_out/r2c11_chain_power.txt gpurun_out/r2c11_scan_variants.txt gpurun_out/r2c11_summary.txt
