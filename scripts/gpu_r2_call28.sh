#!/bin/bash
# chained kernels: wide-store phase-2 epilogue (one 32 x 64 TMA store per warp and tile) vs the two-chunk staging (gpurun_in_attnbase.so)
mkdir -p gpurun_out
O=gpurun_out/r2c28_summary.txt
: > $O
cp kjarni_b200/libkjarni_cuda.so /tmp/new.so
timeout 900 python -m pytest tests/test_gpu_kernels.py -k "chained" -x -q 2>&1 | tail -3 >> $O
timeout 900 python -m pytest tests/test_gpu_encoder.py -x -q 2>&1 | tail -3 >> $O
for v in attnbase new attnbase new; do
  if [ $v = new ]; then cp /tmp/new.so kjarni_b200/libkjarni_cuda.so; else cp gpurun_in_$v.so kjarni_b200/libkjarni_cuda.so; fi
  echo "== $v" >> $O
  RANDOM_DATA=1 ITERS=2000 timeout 300 python scripts/chain_micro.py 2>&1 | grep "chained" >> $O
  timeout 600 python bench.py --no-index --no-cpu > gpurun_out/r2c28_bench_${v}.json 2> gpurun_out/r2c28_bench_${v}.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c28_bench_${v}.json'))
print('$v', d['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()}, {k:c['value'] for k,c in d['configs'].items()})" >> $O 2>&1
done
cp /tmp/new.so kjarni_b200/libkjarni_cuda.so
KNOCKOUT=1 RANDOM_DATA=1 ITERS=1000 timeout 600 python scripts/chain_micro.py 2>&1 | grep "\[smem" >> $O
cat $O
