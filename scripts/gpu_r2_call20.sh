#!/bin/bash
# 4 GPUs: torchrun bench as the driver launches it (+ reference arm), in-process multi-GPU tests and bench
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2c20_bench_4gpu.json 2> gpurun_out/r2c20_bench_4gpu.err; echo "torchrun rc=$?" > gpurun_out/r2c20_summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --impl reference --steps 3 --warmup 1 > gpurun_out/r2c20_ref_4gpu.json 2> gpurun_out/r2c20_ref_4gpu.err; echo "ref rc=$?" >> gpurun_out/r2c20_summary.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2c20_multi.log 2>&1; echo "multi rc=$?" >> gpurun_out/r2c20_summary.txt
timeout 900 python scripts/inproc_bench.py --gpus 4 > gpurun_out/r2c20_inproc.json 2> gpurun_out/r2c20_inproc.err; echo "inproc rc=$?" >> gpurun_out/r2c20_summary.txt
cat gpurun_out/r2c20_summary.txt; tail -3 gpurun_out/r2c20_multi.log
