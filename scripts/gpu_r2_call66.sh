#!/bin/bash
# crossover of chained launches vs one kernel per op by tile count (KJC_CHAIN_MIN_TILES = 0: always chained, 10000: never), then the new tests + smoke
mkdir -p gpurun_out
O=gpurun_out/r2c66_summary.txt
: > $O
for e in "KJC_CHAIN_MIN_TILES=0" "KJC_CHAIN_MIN_TILES=10000" "KJC_CHAIN_MIN_TILES=0" "KJC_CHAIN_MIN_TILES=10000" ""; do
  env $e timeout 300 python scripts/c1_latency_ab.py 8 32 64 80 96 112 128 148 2>&1 | grep "^B=" >> $O
done
timeout 600 python -m pytest tests/test_gpu_encoder.py -x -q -k "small_batches or micro_batch_wider or chained_launch" 2>&1 | tail -4 >> $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> $O
cat $O
