#!/bin/bash
mkdir -p gpurun_out
for cfg in "X=1" "KJC_BN_I=192" "KJC_NO_PDL=1 KJC_BN_I=192" "KJC_BN_I=128"; do
  echo "== $cfg"; env $cfg timeout 300 python bench.py --no-index --no-cpu --steps 10 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})"
done 2>&1 | tee gpurun_out/pdl.txt
