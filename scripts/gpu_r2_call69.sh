#!/bin/bash
# single-chunk host-buffer forward on the compute stream: parity tests through the host path + bench (C1 e2e)
mkdir -p gpurun_out
O=gpurun_out/r2c69_summary.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_ffi.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4 >> $O
timeout 900 python bench.py --no-index --no-cpu > gpurun_out/r2c69_bench.json 2> gpurun_out/r2c69_bench.err
tail -c 300 gpurun_out/r2c69_bench.err >> $O
python -c "
import json
d=json.load(open('gpurun_out/r2c69_bench.json'))
print(d['value'], d['e2e']['value'], d['roofline']['whole_step'], d['clocks'])
print({k:(c['value'], c['e2e']['value'], c['ms_per_step'], c['e2e']['ms_per_step']) for k,c in d['configs'].items()})" >> $O 2>&1
cat $O
