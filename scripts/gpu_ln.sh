#!/bin/bash
# Fused GEMM + residual + LayerNorm kernel: parity, microbenchmark, whole-encoder bench.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "fused_gemm_residual_layernorm" --no-header -p no:cacheprovider 2>&1 | tail -5
timeout 120 python - <<'PY' 2>&1 | grep -v "^trace" | tee gpurun_out/ln_micro.txt
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np
from kjarni_b200 import _native as N
lib = N.lib()
M = 18944
for K in (384, 1536):
    a = np.zeros((M, K), np.uint16); w = np.zeros((384, K), np.uint16); r = np.zeros((M, 384), np.uint16); o = np.empty((M, 384), np.uint16)
    v = np.ones(384, np.float32); us = C.c_float()
    N.check(lib.kjc_dbg_gemm_ln(a.ctypes.data, w.ctypes.data, v.ctypes.data, v.ctypes.data, v.ctypes.data, 1e-12, r.ctypes.data, M, K,
                                o.ctypes.data, 30, C.byref(us)))
    print(f"gemm_ln384 K={K}: {us.value:.1f} us ({2.0*M*384*K/us.value/1e6:.0f} TF)")
PY
TESTS="test_gpu_encoder" bash scripts/gpu_tests.sh | tail -3
timeout 300 python bench.py --no-index --no-cpu --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['roofline']['whole_step']['frac'], {k:round(v['ms_per_step']/v['launches_per_step']*1000,1) for k,v in d['roofline']['kernels'].items()})" | tee gpurun_out/ln_bench.txt
