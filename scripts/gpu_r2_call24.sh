#!/bin/bash
# chained kernels with x' as the phase-2 A operand in tensor memory (KJC_CHAIN_TS): parity, micro timing, whole-step A/B on one box
mkdir -p gpurun_out
O=gpurun_out/r2c24_summary.txt
: > $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "chained" -x -q 2>&1 | tail -5 >> $O
timeout 600 python -m pytest tests/test_gpu_encoder.py -k "chained_launch_variants" -x -q 2>&1 | tail -5 >> $O
timeout 300 python scripts/chain_micro.py >> $O 2>&1
RANDOM_DATA=1 ITERS=3000 timeout 300 python scripts/chain_micro.py >> $O 2>&1
for ts in 0 1 0 1; do
  KJC_CHAIN_TS=$ts timeout 600 python bench.py --no-index --no-cpu > gpurun_out/r2c24_bench_ts$ts.json 2> gpurun_out/r2c24_bench_ts$ts.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c24_bench_ts$ts.json'))
print('ts=$ts', d['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})" >> $O
done
cat $O
