#!/bin/bash
# phase-2 knock-outs of the chained kernels: what bounds the weight-streaming projection (loads / MMAs / epilogue)?
mkdir -p gpurun_out
O=gpurun_out/r2c27_summary.txt
: > $O
KNOCKOUT=1 RANDOM_DATA=1 ITERS=1000 timeout 600 python scripts/chain_micro.py >> $O 2>&1
cat $O
