#!/bin/bash
# MMA issue style A/B on one box: round-1 `if (lane == 0)` issuer (gpurun_in_lane0.so) vs warp-uniform issuer with one elect per
# k-block (default build); each with x' of the chained launches in shared memory (KJC_CHAIN_TS=0) and in tensor memory (=1)
mkdir -p gpurun_out
O=gpurun_out/r2c25_summary.txt
: > $O
cp kjarni_b200/libkjarni_cuda.so /tmp/new.so
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -3 >> $O
timeout 600 python -m pytest tests/test_gpu_encoder.py -k "chained_launch_variants or embedding_matches_oracle" -x -q 2>&1 | tail -3 >> $O
timeout 600 python -m pytest tests/test_gpu_scan.py -x -q 2>&1 | tail -3 >> $O
for v in lane0 new lane0 new; do
  if [ $v = new ]; then cp /tmp/new.so kjarni_b200/libkjarni_cuda.so; else cp gpurun_in_$v.so kjarni_b200/libkjarni_cuda.so; fi
  echo "== $v" >> $O
  RANDOM_DATA=1 ITERS=2000 timeout 300 python scripts/chain_micro.py >> $O 2>&1
  for ts in 0 1; do
    KJC_CHAIN_TS=$ts timeout 600 python bench.py --no-cpu > gpurun_out/r2c25_bench_${v}_ts$ts.json 2> gpurun_out/r2c25_bench_${v}_ts$ts.err
    python -c "
import json
d=json.load(open('gpurun_out/r2c25_bench_${v}_ts$ts.json'))
print('$v ts=$ts', d['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()}, {k:c['value'] for k,c in d['configs'].items()}, {k:(c.get('value'), c.get('ms')) for k,c in d.get('index_topk',{}).items() if isinstance(c,dict)})" >> $O 2>&1
  done
done
cp /tmp/new.so kjarni_b200/libkjarni_cuda.so
cat $O
