"""MMA-rate probe (kjc_dbg_gemm_time): long-K shapes so the k-loop dominates; clocks per 128xBNx16 MMA at an assumed SM clock."""
import ctypes as C, sys
sys.path.insert(0, ".")
from kjarni_b200 import _native as N
lib = N.lib()
M = 18944
MHZ = 1780.0
def t(Nn, K, epi, bn, flags):
    us = C.c_float()
    N.check(lib.kjc_dbg_gemm_time(M, Nn, K, epi, 0, bn, flags, 20, C.byref(us)))
    return us.value
for Nn, K in ((1536, 3072), (1536, 384)):
    for bn in (128, 192, 256):
        if Nn % bn: continue
        tiles = Nn // bn
        mmas = tiles * (K // 16)
        for fl, nm in ((69, "mma-raw"), (5, "mma-only"), (0, "full")):
            us = t(Nn, K, 0, bn, fl)
            print(f"N={Nn} K={K} BN={bn} {nm:8s}: {us:7.1f} us  -> {us*MHZ/mmas:6.1f} clk/MMA (floor {bn/2:.0f})  {2.0*M*Nn*K/us/1e6:.0f} TF", flush=True)
