#!/bin/bash
# Chunked micro-batches with the next chunk's ctx / x rows prefetched into L2 by the FFN-down launch (cp.async.bulk.prefetch.L2): A/B on one box
mkdir -p gpurun_out
O=gpurun_out/r2c62_summary.txt
: > $O
timeout 600 python -m pytest tests/test_gpu_encoder.py -x -q -k "micro_batch_wider or chained_launch" 2>&1 | tail -5 >> $O
for v in 18944 37888 37888np 75776 75776np 18944 37888 75776; do
  echo "== KJC_MICRO_TOKENS=$v" >> $O
  if [[ $v == *np ]]; then export KJC_NO_CHUNK_PREFETCH=1; else unset KJC_NO_CHUNK_PREFETCH; fi
  KJC_MICRO_TOKENS=${v%np} timeout 600 python bench.py --batch 4144 --no-index --no-cpu --no-extra > gpurun_out/r2c62_bench_${v}.json 2> gpurun_out/r2c62_bench_${v}.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c62_bench_${v}.json'))
print('$v', d['value'], d['e2e']['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})" >> $O 2>&1
done
cat $O
