#!/bin/bash
# CTA-pair chained kernels with deeper rings (4 x 40 KB phase-1 stages, 6 x 12 KB W2 stages): parity, timing, trace
mkdir -p gpurun_out
O=gpurun_out/r2c30_summary.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -k "chained" -x -q 2>&1 | tail -3 >> $O
RANDOM_DATA=1 ITERS=2000 timeout 300 python scripts/chain_micro.py 2>&1 | grep -i "chained" >> $O
for pair in 0 1 0 1; do
  if [ $pair = 1 ]; then export KJC_CHAIN_PAIR=1; else unset KJC_CHAIN_PAIR; fi
  timeout 600 python bench.py --no-index --no-cpu > gpurun_out/r2c30_bench_pair$pair.json 2> gpurun_out/r2c30_bench_pair$pair.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c30_bench_pair$pair.json'))
print('pair=$pair', d['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})" >> $O 2>&1
done
unset KJC_CHAIN_PAIR
cp kjarni_b200/libkjarni_cuda.so /tmp/new.so
cp gpurun_in_trace.so kjarni_b200/libkjarni_cuda.so
VARIANTS=16 KOS=0 KJC_LG_TRACE=1 timeout 300 python scripts/chain_trace.py > gpurun_out/r2c30_trace.txt 2>&1
cp /tmp/new.so kjarni_b200/libkjarni_cuda.so
cat $O
