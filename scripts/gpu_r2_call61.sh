#!/bin/bash
# Micro-batches wider than one tile per SM: attention / embedding / pooling cover 2 or 4 x 148 sequences per launch, the chained
# launches walk row chunks of 148 tiles (tile_base).  A/B on one box: KJC_MICRO_TOKENS = 18944 (one tile per SM), 37888, 75776.
mkdir -p gpurun_out
O=gpurun_out/r2c61_summary.txt
: > $O
timeout 600 python -m pytest tests/test_gpu_encoder.py -x -q -k "micro_batch_wider or chained_launch or embedding_matches" 2>&1 | tail -5 >> $O
for v in 18944 37888 75776 18944 37888 75776; do
  echo "== KJC_MICRO_TOKENS=$v" >> $O
  KJC_MICRO_TOKENS=$v timeout 600 python bench.py --batch 4144 --no-index --no-cpu --no-extra > gpurun_out/r2c61_bench_${v}.json 2> gpurun_out/r2c61_bench_${v}.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c61_bench_${v}.json'))
print('$v', d['value'], d['e2e']['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})" >> $O 2>&1
done
cat $O
