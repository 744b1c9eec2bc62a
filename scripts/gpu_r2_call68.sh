#!/bin/bash
# attention_ts, S <= 128: de-phased slots (KJC_ATTN_STAGGER = cycles of start delay per slot).  Launch time via scripts/attn_trace.py.
mkdir -p gpurun_out
O=gpurun_out/r2c68_summary.txt
: > $O
for s in 0 500 1000 1500 2000 3000 0 1000; do
  echo "== KJC_ATTN_STAGGER=$s" >> $O
  KJC_ATTN_STAGGER=$s timeout 120 python scripts/attn_trace.py 148 128 384 12 2>&1 | grep -i "us\b\|launch" | tail -3 >> $O
  KJC_ATTN_STAGGER=$s timeout 120 python scripts/attn_trace.py 148 128 768 12 2>&1 | grep -i "us\b\|launch" | tail -1 >> $O
done
cat $O
