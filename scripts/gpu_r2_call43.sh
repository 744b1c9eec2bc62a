#!/bin/bash
# compute-sanitizer over the kernels that changed in the second half of round 2 (CTA-pair chained / GEMM / scan filter, attention)
mkdir -p gpurun_out
O=gpurun_out/r2c43_sanitizer.txt
{ echo "== memcheck: chained kernels (all variants, small shapes) + CTA-pair GEMM"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "(chained and (77 or 128-384 or 300)) or (cta_pair and (130 or 128-512 or 300-1152))" 2>&1 | tail -6
  echo "== memcheck: attention (tcgen05 kernels, every shape of the parity suite)"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" 2>&1 | tail -6
  echo "== memcheck: scan filter path (pair filter pass) vs oracle"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py -q -x -k "gemm_filter" 2>&1 | tail -6
  echo "== memcheck: encoder forward (tiny models + chained launch variants on MiniLM-L6)"; timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_encoder.py -q -x -k "tiny-bert-6-16 or chained_launch_variants" 2>&1 | tail -6
  echo "== racecheck: chained pair kernel + pair GEMM (smallest shapes)"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "(chained and 77) or (cta_pair and 128-512-64) or (fused_gemm_residual_layernorm and 77-64-768)" 2>&1 | tail -6
} > $O 2>&1
cat $O
