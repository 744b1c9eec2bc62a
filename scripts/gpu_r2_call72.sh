#!/bin/bash
# attention_ts loads-only probe (KJC_ATTN_DBG=1): 64-byte head rows (head_dim 32) against 128-byte head rows (head_dim 64), same number of rows
mkdir -p gpurun_out
O=gpurun_out/r2c72_summary.txt
: > $O
for dbg in 1 0; do
  for shape in "148 128 384 12" "148 128 768 12" "148 128 768 24" "148 128 384 6"; do
    echo "== KJC_ATTN_DBG=$dbg $shape" >> $O
    KJC_ATTN_DBG=$dbg timeout 120 python scripts/attn_trace.py $shape 2>&1 | grep "us/launch" >> $O
  done
done
cat $O
