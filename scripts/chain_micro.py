"""Chained GEMM+LN -> GEMM launch vs the two separate launches (kjc_dbg_* timing hooks, M = 18944)."""
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np
from kjarni_b200 import _native as N
lib = N.lib()
M = 18944
def ln(K):
    a = np.zeros((M, K), np.uint16); w = np.zeros((384, K), np.uint16); r = np.zeros((M, 384), np.uint16); o = np.empty((M, 384), np.uint16)
    v = np.ones(384, np.float32); us = C.c_float()
    N.check(lib.kjc_dbg_gemm_ln(a.ctypes.data, w.ctypes.data, v.ctypes.data, v.ctypes.data, v.ctypes.data, 1e-12, r.ctypes.data, M, K, o.ctypes.data, 30, C.byref(us)))
    return us.value
def gemm(Nn, epi, bn):
    us = C.c_float()
    N.check(lib.kjc_dbg_gemm_time(M, Nn, 384, epi, 0, bn, 0, 30, C.byref(us)))
    return us.value
def chain(K1, N2, epi2):
    a = np.zeros((M, K1), np.uint16); w = np.zeros((384, K1), np.uint16); r = np.zeros((M, 384), np.uint16); ox = np.empty((M, 384), np.uint16)
    w2 = np.zeros((N2, 384), np.uint16); b2 = np.zeros(N2, np.float32); o2 = np.empty((M, N2), np.uint16)
    v = np.ones(384, np.float32); us = C.c_float()
    N.check(lib.kjc_dbg_gemm_ln_gemm(a.ctypes.data, w.ctypes.data, v.ctypes.data, v.ctypes.data, v.ctypes.data, 1e-12, r.ctypes.data, M, K1,
                                     w2.ctypes.data, b2.ctypes.data, N2, epi2, 0, ox.ctypes.data, o2.ctypes.data, 30, C.byref(us)))
    return us.value
a, b, c = ln(384), gemm(1536, 1, 256), chain(384, 1536, 1)
print(f"out-proj+LN {a:.1f} us + FFN-up {b:.1f} us = {a+b:.1f} us   |  chained {c:.1f} us")
a, b, c = ln(1536), gemm(1152, 0, 192), chain(1536, 1152, 0)
print(f"FFN-down+LN {a:.1f} us + QKV {b:.1f} us = {a+b:.1f} us   |  chained {c:.1f} us")
