#!/bin/bash
# GELU with the 1/sqrt2 folded into the polynomial + packed bias adds in the phase-2 / wide-store epilogues (default build) vs gpurun_in_base.so
mkdir -p gpurun_out
O=gpurun_out/r2c60_summary.txt
: > $O
cp kjarni_b200/libkjarni_cuda.so /tmp/new.so
timeout 900 python -m pytest tests/test_gpu_kernels.py -k "gemm or chained" -x -q 2>&1 | tail -3 >> $O
timeout 900 python -m pytest tests/test_gpu_encoder.py -x -q 2>&1 | tail -3 >> $O
for v in base new base new; do
  if [ $v = new ]; then cp /tmp/new.so kjarni_b200/libkjarni_cuda.so; else cp gpurun_in_$v.so kjarni_b200/libkjarni_cuda.so; fi
  echo "== $v" >> $O
  RANDOM_DATA=1 ITERS=2000 timeout 300 python scripts/chain_micro.py 2>&1 | grep "pair" >> $O
  timeout 600 python bench.py --no-index --no-cpu > gpurun_out/r2c60_bench_${v}.json 2> gpurun_out/r2c60_bench_${v}.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c60_bench_${v}.json'))
print('$v', d['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()}, {k:c['value'] for k,c in d['configs'].items()})" >> $O 2>&1
done
cp /tmp/new.so kjarni_b200/libkjarni_cuda.so
cat $O
