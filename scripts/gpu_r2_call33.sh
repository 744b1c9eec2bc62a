#!/bin/bash
# pair chained kernels: early weight loads (before griddepcontrol.wait) + early residual prefetch vs without (gpurun_in_noearly.so)
mkdir -p gpurun_out
O=gpurun_out/r2c33_summary.txt
: > $O
cp kjarni_b200/libkjarni_cuda.so /tmp/new.so
timeout 900 python -m pytest tests/test_gpu_kernels.py -k "chained" -x -q 2>&1 | tail -3 >> $O
timeout 900 python -m pytest tests/test_gpu_encoder.py -x -q 2>&1 | tail -3 >> $O
for v in noearly new noearly new; do
  if [ $v = new ]; then cp /tmp/new.so kjarni_b200/libkjarni_cuda.so; else cp gpurun_in_$v.so kjarni_b200/libkjarni_cuda.so; fi
  echo "== $v" >> $O
  RANDOM_DATA=1 ITERS=2000 timeout 300 python scripts/chain_micro.py 2>&1 | grep "pair" >> $O
  timeout 600 python bench.py --no-index --no-cpu > gpurun_out/r2c33_bench_${v}.json 2> gpurun_out/r2c33_bench_${v}.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c33_bench_${v}.json'))
print('$v', d['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()}, {k:c['value'] for k,c in d['configs'].items()})" >> $O 2>&1
done
cp gpurun_in_trace.so kjarni_b200/libkjarni_cuda.so
VARIANTS=16 KOS=0 KJC_LG_TRACE=1 timeout 300 python scripts/chain_trace.py > gpurun_out/r2c33_trace.txt 2>&1
cp /tmp/new.so kjarni_b200/libkjarni_cuda.so
cat $O
