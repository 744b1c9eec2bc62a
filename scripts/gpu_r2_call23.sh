#!/bin/bash
# attention: fraction of the exponentials on the FMA pipe (KJ_ATTN_POLY_PAIRS = 0 / 2 / 3 / 4 pairs of every 8), same box
mkdir -p gpurun_out
cp kjarni_b200/libkjarni_cuda.so /tmp/base.so
: > gpurun_out/r2c23_attn.txt
for v in base poly2 poly3 poly4; do
  if [ $v = base ]; then cp /tmp/base.so kjarni_b200/libkjarni_cuda.so; else cp gpurun_in_$v.so kjarni_b200/libkjarni_cuda.so; fi
  for shape in "148 128 384 12" "148 128 768 12" "74 256 384 12" "37 512 768 12"; do
    timeout 120 python scripts/attn_trace.py $shape 2>&1 | grep "us/launch" | sed "s/^/$v /" >> gpurun_out/r2c23_attn.txt
  done
  timeout 300 python -m pytest tests/test_gpu_kernels.py -k attention -x -q 2>&1 | tail -1 | sed "s/^/$v tests: /" >> gpurun_out/r2c23_attn.txt
  timeout 600 python bench.py --no-index --no-cpu > gpurun_out/r2c23_bench_$v.json 2> gpurun_out/r2c23_bench_$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c23_bench_$v.json'))
print('$v', d['value'], d['roofline']['kernels']['attention']['ms_per_step'], {k:(c['value'], c['kernels']['attention']['ms']) for k,c in d['configs'].items()})" >> gpurun_out/r2c23_attn.txt
done
cp /tmp/base.so kjarni_b200/libkjarni_cuda.so
cat gpurun_out/r2c23_attn.txt
