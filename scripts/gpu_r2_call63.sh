#!/bin/bash
# final r02 validation of HEAD (chunked chained launches in tree, default one tile per SM): full GPU suite, smoke, default bench, ncu launch list + full captures
mkdir -p gpurun_out
O=gpurun_out/r2c63_summary.txt
: > $O
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> $O
timeout 900 python bench.py > gpurun_out/r2c63_bench.json 2> gpurun_out/r2c63_bench.err
tail -c 600 gpurun_out/r2c63_bench.err >> $O
python -c "
import json
d=json.load(open('gpurun_out/r2c63_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
print({k:(c['value'], c.get('roofline',{}).get('frac')) for k,c in d['configs'].items()})
print(d.get('index_topk'))" >> $O 2>&1
bash scripts/gpu_r2_profile.sh > gpurun_out/r2c63_profile.log 2>&1
cat $O
