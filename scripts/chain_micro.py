"""Chained GEMM+LN -> GEMM launch vs the two separate launches (kjc_dbg_* timing hooks, M = 18944)."""
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np
from kjarni_b200 import _native as N
import os
lib = N.lib()
M = 18944
ITERS = int(os.environ.get("ITERS", "30"))
RANDOM = os.environ.get("RANDOM_DATA") == "1"  # zeros draw far less power than real activations: compare both
rng = np.random.default_rng(0)
def data(*shape, scale=1.0):
    if not RANDOM:
        return np.zeros(shape, np.uint16)
    x = (rng.standard_normal(shape) * scale).astype(np.float32)
    return (x.view(np.uint32) >> 16).astype(np.uint16)
def ln(K):
    a = data(M, K); w = data(384, K, scale=K ** -0.5); r = data(M, 384); o = np.empty((M, 384), np.uint16)
    v = np.ones(384, np.float32); us = C.c_float()
    N.check(lib.kjc_dbg_gemm_ln(a.ctypes.data, w.ctypes.data, v.ctypes.data, v.ctypes.data, v.ctypes.data, 1e-12, r.ctypes.data, M, K, o.ctypes.data, ITERS, C.byref(us)))
    return us.value
def gemm(Nn, epi, bn):
    us = C.c_float()
    N.check(lib.kjc_dbg_gemm_time(M, Nn, 384, epi, 0, bn, 0, 30, C.byref(us)))
    return us.value
def chain(K1, N2, epi2):
    a = data(M, K1); w = data(384, K1, scale=K1 ** -0.5); r = data(M, 384); ox = np.empty((M, 384), np.uint16)
    w2 = data(N2, 384, scale=384 ** -0.5); b2 = np.zeros(N2, np.float32); o2 = np.empty((M, N2), np.uint16)
    v = np.ones(384, np.float32); us = C.c_float()
    N.check(lib.kjc_dbg_gemm_ln_gemm(a.ctypes.data, w.ctypes.data, v.ctypes.data, v.ctypes.data, v.ctypes.data, 1e-12, r.ctypes.data, M, K1,
                                     w2.ctypes.data, b2.ctypes.data, N2, epi2, 0, ox.ctypes.data, o2.ctypes.data, ITERS, C.byref(us)))
    return us.value
a, b, c = ln(384), gemm(1536, 1, 256), chain(384, 1536, 1)
print(f"out-proj+LN {a:.1f} us + FFN-up {b:.1f} us = {a+b:.1f} us   |  chained {c:.1f} us")
a, b, c = ln(1536), gemm(1152, 0, 192), chain(1536, 1152, 0)
print(f"FFN-down+LN {a:.1f} us + QKV {b:.1f} us = {a+b:.1f} us   |  chained {c:.1f} us")
c1, c2 = chain(384, 1536, 1 + 16), chain(1536, 1152, 0 + 16)
print(f"CTA-pair chained (cta_group::2, half a weight tile per CTA): out-proj+LN1->FFN-up {c1:.1f} us, FFN-down+LN2->QKV {c2:.1f} us")
c1, c2 = chain(384, 1536, 1 + 32), chain(1536, 1152, 0 + 32)
print(f"x' in tensor memory (TS-form phase 2, 128-column tiles): out-proj+LN1->FFN-up {c1:.1f} us, FFN-down+LN2->QKV {c2:.1f} us")
if os.environ.get("KNOCKOUT") == "1":
    # phase-2 knock-outs (timing only): 64 = no W2 loads, 128 = no phase-2 epilogue work, 256 = no phase-2 MMAs
    for name, K1, N2, e in (("out-proj+LN1->FFN-up", 384, 1536, 1), ("FFN-down+LN2->QKV", 1536, 1152, 0)):
        for var, vname in ((0, "smem x'"), (32, "tmem x'")):
            row = []
            for ko, kname in ((0, "full"), (64, "no-W2-loads"), (128, "no-epi2"), (256, "no-mma2"), (64 + 128, "mma2-only"), (128 + 256, "loads-only"), (64 + 256, "epi2-only"), (64 + 128 + 256, "phase-1-only")):
                row.append(f"{kname} {chain(K1, N2, e + var + ko):.1f}")
            print(f"{name} [{vname}]: " + " | ".join(row))
