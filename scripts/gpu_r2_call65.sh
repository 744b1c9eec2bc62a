#!/bin/bash
# batch-32 latency case under the launch-structure switches
mkdir -p gpurun_out
O=gpurun_out/r2c65_summary.txt
: > $O
for e in "" "KJC_NO_CHAIN=1" "KJC_CHAIN_PAIR=0" "KJC_NO_CHAIN=1 KJC_GEMM_PAIR=0" "KJC_NO_CHAIN=1 KJC_NO_FUSED_LN=1" ""; do
  env $e timeout 300 python scripts/c1_latency_ab.py 32 8 64 2>&1 | grep "^B=" >> $O
done
cat $O
