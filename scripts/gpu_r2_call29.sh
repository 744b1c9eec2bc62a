#!/bin/bash
# timeline trace of the chained kernels (full and the three phase-2 knock-outs)
mkdir -p gpurun_out
cp kjarni_b200/libkjarni_cuda.so /tmp/new.so
cp gpurun_in_trace.so kjarni_b200/libkjarni_cuda.so
KJC_LG_TRACE=1 timeout 300 python scripts/chain_trace.py > gpurun_out/r2c29_trace.txt 2>&1
cp /tmp/new.so kjarni_b200/libkjarni_cuda.so
tail -c 3000 gpurun_out/r2c29_trace.txt
