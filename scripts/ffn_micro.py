"""Fused FFN kernel microbenchmark (kjc_dbg_ffn_ln): time per launch for several activations / intermediate sizes."""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, ".")
from kjarni_b200 import _native as N
lib = N.lib()
M, H = 18944, 384
rng = np.random.default_rng(0)
def bf(a): return (a.astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)
for I in (1536, 768):
    x = bf(rng.standard_normal((M, H))); w1 = bf(rng.standard_normal((I, H)) * 0.05); w2 = bf(rng.standard_normal((H, I)) * 0.03)
    b1 = np.zeros(I, np.float32); v = np.ones(H, np.float32); o = np.empty((M, H), np.uint16)
    for act, name in ((0, "erf-gelu"), (3, "none"), (2, "relu")):
        us = C.c_float()
        N.check(lib.kjc_dbg_ffn_ln(x.ctypes.data, w1.ctypes.data, b1.ctypes.data, w2.ctypes.data, v.ctypes.data, v.ctypes.data, v.ctypes.data, 1e-12,
                                   M, I, act, o.ctypes.data, 30, C.byref(us)))
        print(f"ffn_ln384 I={I} act={name}: {us.value:.1f} us ({4.0*M*H*I/us.value/1e6:.0f} TF)", flush=True)
