"""Timeline of one chained launch (gemm_ln_gemm.cuh built with -DKJ_LG_TRACE=1; KJC_LG_TRACE=1): clock64 stamps of CTAs 0 and 100.
Slots: 1 pdl_wait done | 2 first phase-1 operands | 3 phase-1 MMAs issued | 5 accumulator complete | 6 LN pass A | 7 statistics barrier |
8 x' published | 4 MMA warp sees x' | 16+2nb / 17+2nb MMA warp: accumulator free / tile nb issued | 48+8nb.. (warp 4) and 128+8nb.. (warp 11):
+0 accumulator full, +1 chunk 0 loaded, +2 chunk 0 math, +3 staging free, +4 chunk 1 loaded, +5 chunk 1 math, +6 fence, +7 store issued |
200+it producer: stage free, W2 load issued | 112+nb clocks the MMA warp waited for W2 stages in tile nb | 255 exit."""
import ctypes as C, os, sys
sys.path.insert(0, ".")
import numpy as np
from kjarni_b200 import _native as N
lib = N.lib()
M = 18944
rng = np.random.default_rng(0)
def data(*shape, scale=1.0):
    x = (rng.standard_normal(shape) * scale).astype(np.float32)
    return (x.view(np.uint32) >> 16).astype(np.uint16)
def chain(K1, N2, epi2):
    a = data(M, K1); w = data(384, K1, scale=K1 ** -0.5); r = data(M, 384); ox = np.empty((M, 384), np.uint16)
    w2 = data(N2, 384, scale=384 ** -0.5); b2 = np.zeros(N2, np.float32); o2 = np.empty((M, N2), np.uint16)
    v = np.ones(384, np.float32); us = C.c_float()
    N.check(lib.kjc_dbg_gemm_ln_gemm(a.ctypes.data, w.ctypes.data, v.ctypes.data, v.ctypes.data, v.ctypes.data, 1e-12, r.ctypes.data, M, K1,
                                     w2.ctypes.data, b2.ctypes.data, N2, epi2, 0, ox.ctypes.data, o2.ctypes.data, 0, C.byref(us)))
for var in [int(x) for x in os.environ.get("VARIANTS", "0").split(",")]:
    for ko in [int(x) for x in os.environ.get("KOS", "0,64,128,256").split(",")]:
        chain(384, 1536, 1 + var + ko)
        chain(1536, 1152, 0 + var + ko)
