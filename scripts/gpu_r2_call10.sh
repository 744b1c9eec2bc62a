#!/bin/bash
# CTA-pair chained kernels: parity, micro timing, whole-step A/B on the same box
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "chained" > gpurun_out/r2c10_chain.log 2>&1; echo "chain tests rc=$?" > gpurun_out/r2c10_summary.txt
timeout 300 python scripts/chain_micro.py > gpurun_out/r2c10_chain_micro.txt 2>&1
timeout 600 python bench.py --no-index --no-cpu --no-extra > gpurun_out/r2c10_bench_base.json 2> gpurun_out/r2c10_bench_base.err
KJC_CHAIN_PAIR=1 timeout 600 python bench.py --no-index --no-cpu --no-extra > gpurun_out/r2c10_bench_pair.json 2> gpurun_out/r2c10_bench_pair.err
timeout 600 python bench.py --no-index --no-cpu --no-extra > gpurun_out/r2c10_bench_base2.json 2> gpurun_out/r2c10_bench_base2.err
KJC_CHAIN_PAIR=1 timeout 900 python -m pytest tests/test_gpu_encoder.py -x -q > gpurun_out/r2c10_enc_pair.log 2>&1; echo "encoder tests (pair) rc=$?" >> gpurun_out/r2c10_summary.txt
tail -5 gpurun_out/r2c10_chain.log; cat gpurun_out/r2c10_chain_micro.txt; cat gpurun_out/r2c10_summary.txt
for f in base pair base2; do python -c "
import json,sys
d=json.load(open('gpurun_out/r2c10_bench_$f.json')); k=d['roofline']['kernels']
print('$f', d['value'], d['roofline']['whole_step']['frac'], {n:(v['ms_per_step'],v['frac']) for n,v in k.items()})"; done
