#!/bin/bash
# scan filter pass as CTA pairs (KJC_SG_PAIR=1, default) vs one CTA per row tile; last-layer FFN-down through the pair launch vs gemm_ln_kernel<1>
mkdir -p gpurun_out
O=gpurun_out/r2c41_summary.txt
: > $O
timeout 1200 python -m pytest tests/test_gpu_scan.py tests/test_gpu_index_dir.py -x -q 2>&1 | tail -3 >> $O
for sp in 0 1 0 1; do
  KJC_SG_PAIR=$sp timeout 900 python bench.py --no-cpu --no-extra > gpurun_out/r2c41_bench_sp$sp.json 2> gpurun_out/r2c41_bench_sp$sp.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c41_bench_sp$sp.json'))
i=d['index_topk']
print('sg_pair=$sp', d['value'], i['value'], i['ms_per_step'], i['roofline']['frac'], i['small_batch_8q']['value'], i['exact_scan_8q']['value'], i['unverified_queries'])" >> $O 2>&1
done
for one in 1 0 1 0; do
  if [ $one = 1 ]; then export KJC_LAST_LN_ONE_CTA=1; else unset KJC_LAST_LN_ONE_CTA; fi
  timeout 600 python bench.py --no-index --no-cpu --no-extra > gpurun_out/r2c41_bench_last$one.json 2> gpurun_out/r2c41_bench_last$one.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c41_bench_last$one.json'))
print('last_ln_one_cta=$one', d['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})" >> $O 2>&1
done
cat $O
