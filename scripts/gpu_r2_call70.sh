#!/bin/bash
# 4-GPU validation of HEAD: in-library multi-GPU tests through the C ABI, torchrun bench at N = 4 (both arms), in-process bench
mkdir -p gpurun_out
O=gpurun_out/r2c70_summary.txt
: > $O
nvidia-smi --query-gpu=index,name --format=csv,noheader >> $O
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4 >> $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2c70_bench_4gpu.json 2> gpurun_out/r2c70_bench_4gpu.err
tail -c 400 gpurun_out/r2c70_bench_4gpu.err >> $O
python -c "
import json
d=json.loads(open('gpurun_out/r2c70_bench_4gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['e2e']['value'], d['roofline']['whole_step'], d['clocks'])
i=d.get('index_topk') or {}
print({k:i.get(k) for k in ('value','rows_total','merge','unverified_queries')})" >> $O 2>&1
timeout 600 python scripts/inproc_bench.py --gpus 4 > gpurun_out/r2c70_inproc_4gpu.json 2> gpurun_out/r2c70_inproc_4gpu.err
tail -c 1500 gpurun_out/r2c70_inproc_4gpu.json >> $O
cat $O
