#!/bin/bash
# 2 GPUs: multi-GPU C-ABI tests on distinct devices, in-process bench, torchrun bench (packed NCCL gather)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c19_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2c19_multi.log 2>&1; echo "multi rc=$?" > gpurun_out/r2c19_summary.txt
timeout 900 python scripts/inproc_bench.py --gpus 2 > gpurun_out/r2c19_inproc.json 2> gpurun_out/r2c19_inproc.err; echo "inproc rc=$?" >> gpurun_out/r2c19_summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu --no-extra > gpurun_out/r2c19_bench_2gpu.json 2> gpurun_out/r2c19_bench_2gpu.err; echo "torchrun rc=$?" >> gpurun_out/r2c19_summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/r2c19_ref_2gpu.json 2> gpurun_out/r2c19_ref_2gpu.err; echo "ref rc=$?" >> gpurun_out/r2c19_summary.txt
tail -5 gpurun_out/r2c19_multi.log; cat gpurun_out/r2c19_inproc.json; cat gpurun_out/r2c19_summary.txt
