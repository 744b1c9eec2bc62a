"""GEMM microbenchmark matrix (kjc_dbg_gemm_time): which part of the pipeline bounds each encoder GEMM shape."""
import ctypes as C, sys
sys.path.insert(0, ".")
from kjarni_b200 import _native as N
lib = N.lib()
M = 18944
def t(Nn, K, epi, bn, flags, act=0):
    us = C.c_float()
    N.check(lib.kjc_dbg_gemm_time(M, Nn, K, epi, act, bn, flags, 30, C.byref(us)))
    return us.value
for name, Nn, K, epi in (("qkv", 1152, 384, 0), ("ffn_up", 1536, 384, 1), ("out", 384, 384, 2), ("ffn_down", 384, 1536, 2)):
    for bn in ((128, 192) if Nn % 192 == 0 else (128, 256)):
        if Nn % bn: continue
        row = {f: t(Nn, K, epi, bn, f) for f in (0, 1, 2, 3, 4, 5, 6, 7)}
        fl = 2.0 * M * Nn * K
        print(f"{name:9s} N={Nn} K={K} BN={bn}: full {row[0]:.1f}us ({fl/row[0]/1e6:.0f} TF) | no-epi {row[1]:.1f} | no-mma {row[2]:.1f} | "
              f"tma-only {row[3]:.1f} | no-tma {row[4]:.1f} | mma-only {row[5]:.1f} | epi-only {row[6]:.1f} | empty {row[7]:.1f}")

for name, Nn, K, epi, bn in (("qkv", 1152, 384, 0, 192), ("ffn_up", 1536, 384, 1, 256)):
    us = t(Nn, K, epi, 1000 + bn, 0)
    print(f"PAIR {name:9s} N={Nn} K={K} BN={bn}: {us:.1f}us ({2.0*M*Nn*K/us/1e6:.0f} TF)")

# fused out-proj/FFN-down + residual + LayerNorm kernel
import numpy as np
for K in (384, 1536):
    a = np.zeros((M, K), np.uint16); w = np.zeros((384, K), np.uint16); r = np.zeros((M, 384), np.uint16); o = np.empty((M, 384), np.uint16)
    v = np.ones(384, np.float32); us = C.c_float()
    N.check(lib.kjc_dbg_gemm_ln(a.ctypes.data, w.ctypes.data, v.ctypes.data, v.ctypes.data, v.ctypes.data, 1e-12, r.ctypes.data, M, K,
                                o.ctypes.data, 30, C.byref(us)))
    print(f"gemm_ln384 K={K}: {us.value:.1f} us ({2.0*M*384*K/us.value/1e6:.0f} TF)")
