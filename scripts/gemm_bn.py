"""Tile-width A/B for the K=384 projections (kjc_dbg_gemm_time, GELU epilogue for FFN-up)."""
import ctypes as C, sys
sys.path.insert(0, ".")
from kjarni_b200 import _native as N
lib = N.lib()
M = 18944
def t(Nn, K, epi, bn, flags=0):
    us = C.c_float()
    N.check(lib.kjc_dbg_gemm_time(M, Nn, K, epi, 0, bn, flags, 30, C.byref(us)))
    return us.value
for name, Nn, K, epi in (("qkv", 1152, 384, 0), ("ffn_up", 1536, 384, 1)):
    for bn in (192, 256):
        print(f"{name} BN={bn}: full {t(Nn, K, epi, bn):.1f} us | no-epi {t(Nn, K, epi, bn, 1):.1f} | mma-only {t(Nn, K, epi, bn, 5):.1f}", flush=True)
for name, Nn, K, epi in (("qkv", 1152, 384, 0), ("ffn_up", 1536, 384, 1)):
    for bn in (192, 256):
        print(f"PAIR {name} BN={bn}: full {t(Nn, K, epi, 1000 + bn):.1f} us", flush=True)
