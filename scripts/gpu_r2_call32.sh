#!/bin/bash
# full GPU suite + bench of the new default (both chained launches as CTA pairs with deep rings)
mkdir -p gpurun_out
O=gpurun_out/r2c32_summary.txt
: > $O
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $O
timeout 900 python bench.py > gpurun_out/r2c32_bench.json 2> gpurun_out/r2c32_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2c32_bench.json'))
print(d['value'], d['e2e'], d['roofline']['frac'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
print({k:(c['value'], c.get('roofline',{}).get('frac')) for k,c in d['configs'].items()})" >> $O 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> $O
cat $O
