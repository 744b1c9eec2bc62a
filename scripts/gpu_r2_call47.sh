#!/bin/bash
# attention max pass: two 32-key chunks per tcgen05.wait::ld (default build) vs one (gpurun_in_base.so), same box
mkdir -p gpurun_out
O=gpurun_out/r2c47_summary.txt
: > $O
cp kjarni_b200/libkjarni_cuda.so /tmp/new.so
timeout 600 python -m pytest tests/test_gpu_kernels.py -k attention -x -q 2>&1 | tail -2 >> $O
for v in base new base new; do
  if [ $v = new ]; then cp /tmp/new.so kjarni_b200/libkjarni_cuda.so; else cp gpurun_in_$v.so kjarni_b200/libkjarni_cuda.so; fi
  for shape in "148 128 384 12" "148 128 768 12" "74 256 384 12" "37 512 768 12"; do
    timeout 120 python scripts/attn_trace.py $shape 2>&1 | grep "us/launch" | sed "s/^/$v /" >> $O
  done
  timeout 600 python bench.py --no-index --no-cpu > gpurun_out/r2c47_bench_$v.json 2> gpurun_out/r2c47_bench_$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c47_bench_$v.json'))
print('$v', d['value'], d['roofline']['kernels']['attention']['ms_per_step'], {k:(c['value'], c['kernels']['attention']['ms']) for k,c in d['configs'].items()})" >> $O
done
cp /tmp/new.so kjarni_b200/libkjarni_cuda.so
cat $O
