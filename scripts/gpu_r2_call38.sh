#!/bin/bash
# 2 GPUs: multi-GPU tests through the C ABI, torchrun bench, in-process bench
mkdir -p gpurun_out
O=gpurun_out/r2c38_summary.txt
: > $O
nvidia-smi -L >> $O
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3 >> $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu > gpurun_out/r2c38_bench_2gpu.json 2> gpurun_out/r2c38_bench_2gpu.err
python -c "
import json
d=json.load(open('gpurun_out/r2c38_bench_2gpu.json'))
print('2gpu', d['value'], d['e2e']['value'], d['n_gpus'], d.get('index_topk',{}).get('value'))" >> $O 2>&1
timeout 600 python scripts/inproc_bench.py --gpus 2 > gpurun_out/r2c38_inproc_2gpu.json 2> gpurun_out/r2c38_inproc.err
tail -c 800 gpurun_out/r2c38_inproc_2gpu.json >> $O
cat $O
