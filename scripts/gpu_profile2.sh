#!/bin/bash
mkdir -p gpurun_out
python scripts/cublas_shapes.py > gpurun_out/cublas_shapes.txt 2>&1
CMD="python bench.py --steps 1 --warmup 3 --batch 296 --index-rows 2000000 --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 24 -c 4 -o gpurun_out/prof_gemm2 -f $CMD > gpurun_out/prof_gemm2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 2 -c 1 -o gpurun_out/prof_scan2 -f $CMD > gpurun_out/prof_scan2.log 2>&1
cat gpurun_out/cublas_shapes.txt
