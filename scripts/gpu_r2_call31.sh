#!/bin/bash
# CTA-pair chained kernels: relaxed accumulator-release arrive; which launches should pair? (KJC_CHAIN_PAIR mask: 1 = FFN-down chain, 2 = FFN-up chain)
mkdir -p gpurun_out
O=gpurun_out/r2c31_summary.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -k "chained" -x -q 2>&1 | tail -3 >> $O
RANDOM_DATA=1 ITERS=2000 timeout 300 python scripts/chain_micro.py 2>&1 | grep -i "chained" >> $O
for pair in 0 1 2 3 0 1 2 3; do
  KJC_CHAIN_PAIR=$pair timeout 600 python bench.py --no-index --no-cpu > gpurun_out/r2c31_bench_pair$pair.json 2> gpurun_out/r2c31_bench_pair$pair.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c31_bench_pair$pair.json'))
print('pair=$pair', d['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})" >> $O 2>&1
done
cp kjarni_b200/libkjarni_cuda.so /tmp/new.so
cp gpurun_in_trace.so kjarni_b200/libkjarni_cuda.so
VARIANTS=16 KOS=0 KJC_LG_TRACE=1 timeout 300 python scripts/chain_trace.py > gpurun_out/r2c31_trace.txt 2>&1
cp /tmp/new.so kjarni_b200/libkjarni_cuda.so
cat $O
