#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_ffi.py tests/test_gpu_multi.py tests/test_gpu_scan.py -q > gpurun_out/r2c21_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2c21_summary.txt
timeout 900 python bench.py --no-cpu > gpurun_out/r2c21_bench.json 2> gpurun_out/r2c21_bench.err; echo "bench rc=$?" >> gpurun_out/r2c21_summary.txt
tail -5 gpurun_out/r2c21_tests.log; cat gpurun_out/r2c21_summary.txt
python -c "
import json
d=json.load(open('gpurun_out/r2c21_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['whole_step'])
for k,c in d['configs'].items(): print(k, c['value'], c['e2e']['value'], c['roofline']['frac'])
it=d['index_topk']; print(it['value'], it['small_batch_8q']['roofline']['frac'], it['exact_scan_8q']['roofline']['frac'])
"
