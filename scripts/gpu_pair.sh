#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "pair" --no-header -p no:cacheprovider 2>&1 | tail -15
timeout 200 python scripts/gemm_micro.py 2>&1 | tee gpurun_out/gemm_micro.txt
