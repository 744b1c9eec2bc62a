#!/bin/bash
# r02 ncu evidence for profiles/: launch list (durations) of a short default bench run + full captures of the hot encoder kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --batch 296 --index-rows 1000000 --no-cpu --no-extra"
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 700 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|gemm_ln_kernel|gemm_ln_gemm" -s 18 -c 6 -o gpurun_out/r2_prof_gemm -f $CMD > gpurun_out/r2_prof_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention -s 6 -c 1 -o gpurun_out/r2_prof_attn -f $CMD > gpurun_out/r2_prof_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pool_l2|embed_layernorm" -s 4 -c 2 -o gpurun_out/r2_prof_rows -f $CMD > gpurun_out/r2_prof_rows.log 2>&1
ls -la gpurun_out/r2_prof_* gpurun_out/r2_launches.csv
