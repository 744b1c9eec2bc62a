#!/bin/bash
# Micro-batch size sweep (tokens per internal micro-batch) for the encoder bench: amortises per-kernel bubbles vs L2 residency.
mkdir -p gpurun_out
for mt in ${SWEEP:-18944 37888 56832 75776}; do
  KJC_MICRO_TOKENS=$mt timeout 300 python bench.py --no-index --no-cpu --steps 10 --batch 3552 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('micro_tokens=$mt', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', d['roofline']['whole_step']['frac'], {k:round(v['ms_per_step']/v['launches_per_step']*1000,1) for k,v in d['roofline']['kernels'].items()})"
done | tee gpurun_out/micro_sweep.txt
