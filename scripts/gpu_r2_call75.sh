#!/bin/bash
# validation of HEAD of the final HEAD (qkv over h): full GPU suite, smoke, default bench
mkdir -p gpurun_out
O=gpurun_out/r2c75_summary.txt
: > $O
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> $O
timeout 900 python bench.py > gpurun_out/r2c75_bench.json 2> gpurun_out/r2c75_bench.err
tail -c 600 gpurun_out/r2c75_bench.err >> $O
python -c "
import json
d=json.load(open('gpurun_out/r2c75_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_step'], d['clocks'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
print({k:(c['value'], c['e2e']['value'], c.get('roofline',{}).get('frac'), c['gpu_launches_per_step']) for k,c in d['configs'].items()})
i=d['index_topk']; print(i['value'], i['roofline']['frac'], i['small_batch_8q']['value'], i['exact_scan_8q']['value'])" >> $O 2>&1
cat $O
