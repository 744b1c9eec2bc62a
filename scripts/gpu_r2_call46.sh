#!/bin/bash
# pair chained kernels: early residual prefetch alone (gpurun_in_earlyres.so, KJ_LG_EARLY=2) vs default, same box
mkdir -p gpurun_out
O=gpurun_out/r2c46_summary.txt
: > $O
cp kjarni_b200/libkjarni_cuda.so /tmp/new.so
for v in new earlyres new earlyres; do
  if [ $v = new ]; then cp /tmp/new.so kjarni_b200/libkjarni_cuda.so; else cp gpurun_in_$v.so kjarni_b200/libkjarni_cuda.so; fi
  timeout 600 python bench.py --no-index --no-cpu --no-extra > gpurun_out/r2c46_bench_${v}.json 2> gpurun_out/r2c46_bench_${v}.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c46_bench_${v}.json'))
print('$v', d['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})" >> $O 2>&1
done
cp gpurun_in_earlyres.so kjarni_b200/libkjarni_cuda.so
timeout 600 python -m pytest tests/test_gpu_kernels.py -k chained -x -q 2>&1 | tail -2 >> $O
cp /tmp/new.so kjarni_b200/libkjarni_cuda.so
cat $O
