#!/bin/bash
# CTA-pair form of the stand-alone GEMM: parity, micro timing, bench A/B (KJC_GEMM_PAIR = 0 / 3)
mkdir -p gpurun_out
O=gpurun_out/r2c39_summary.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_kernels.py -k "gemm" -x -q 2>&1 | tail -3 >> $O
timeout 300 python scripts/gemm_pair_micro.py >> $O 2>&1
for gp in 0 3 0 3; do
  KJC_GEMM_PAIR=$gp timeout 600 python bench.py --no-index --no-cpu > gpurun_out/r2c39_bench_gp$gp.json 2> gpurun_out/r2c39_bench_gp$gp.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c39_bench_gp$gp.json'))
print('gemm_pair=$gp', d['value'], round(d['roofline']['kernels']['gemm_qkv']['ms_per_step'],3), {k:(c['value'], {kk:round(v['ms'],3) for kk,v in c['kernels'].items() if kk in ('gemm_qkv','gemm_ffn_up')}) for k,c in d['configs'].items()})" >> $O 2>&1
done
cat $O
