#!/bin/bash
# qkv rows laid over the FFN activation rows (one buffer, pitch I): parity, then A/B against separate buffers (KJC_NO_ALIAS_QKV=1) on one box
mkdir -p gpurun_out
O=gpurun_out/r2c74_summary.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_ffi.py -x -q -m gpu 2>&1 | tail -4 >> $O
for v in alias sep alias sep; do
  if [ $v = sep ]; then export KJC_NO_ALIAS_QKV=1; else unset KJC_NO_ALIAS_QKV; fi
  echo "== $v" >> $O
  timeout 600 python bench.py --no-index --no-cpu > gpurun_out/r2c74_bench_${v}.json 2> gpurun_out/r2c74_bench_${v}.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c74_bench_${v}.json'))
print('$v', d['value'], d['e2e']['value'], d['clocks'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()}, {k:c['value'] for k,c in d['configs'].items()})" >> $O 2>&1
done
cat $O
