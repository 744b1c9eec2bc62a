#!/bin/bash
# 2 GPUs: multi-GPU C ABI tests + in-process bench; full suite on device 0
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c6_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2c6_multi.log 2>&1; echo "multi rc=$?" > gpurun_out/r2c6_summary.txt
timeout 900 python scripts/inproc_bench.py --gpus 2 > gpurun_out/r2c6_inproc.json 2> gpurun_out/r2c6_inproc.err; echo "inproc rc=$?" >> gpurun_out/r2c6_summary.txt
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/r2c6_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c6_summary.txt
tail -15 gpurun_out/r2c6_multi.log; cat gpurun_out/r2c6_inproc.json; tail -5 gpurun_out/r2c6_tests.log; cat gpurun_out/r2c6_summary.txt
