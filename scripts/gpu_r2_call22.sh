#!/bin/bash
mkdir -p gpurun_out
{ echo "== memcheck: scan tests (small shapes)"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scan.py -q -x -k "test_search_matches_oracle or test_reference_vector_store_kats or test_segment_semantics_and_merge or zero_query" 2>&1 | tail -8
  echo "== memcheck: multi-GPU C ABI (duplicated device)"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_multi.py -q -x -k "sharded_index_matches or zero_norm" 2>&1 | tail -8
  echo "== memcheck: fused GEMM+LN 768 / chained pair (small)"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "fused_gemm_residual_layernorm and (128-768 or 77-64-768 or 1000-3072) or (chained and 77)" 2>&1 | tail -8
  echo "== memcheck: tiny encoder forward incl. fp32 residual"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_encoder.py -q -x -k "tiny-bert-6-16 or tiny-distilbert-6-16" 2>&1 | tail -8
} > gpurun_out/r2c22_sanitizer.txt 2>&1
cat gpurun_out/r2c22_sanitizer.txt
