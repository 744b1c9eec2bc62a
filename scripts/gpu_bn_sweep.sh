#!/bin/bash
# In-situ tile-width A/B for the QKV and FFN-up GEMMs (KJC_BN_QKV / KJC_BN_I tuning hooks).
mkdir -p gpurun_out
for cfg in "192 256" "256 256" "192 192" "256 192"; do
  set -- $cfg
  KJC_BN_QKV=$1 KJC_BN_I=$2 timeout 300 python bench.py --no-index --no-cpu --steps 10 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('bn_qkv=$1 bn_i=$2', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', d['roofline']['whole_step']['frac'], {k:round(v['ms_per_step']/v['launches_per_step']*1000,1) for k,v in d['roofline']['kernels'].items()})"
done | tee gpurun_out/bn_sweep.txt
