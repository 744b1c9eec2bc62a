"""Per-CTA milestone trace of the GEMM kernel (kjc_dbg_gemm_time flag 8): where one launch spends its time."""
import ctypes as C, sys
sys.path.insert(0, ".")
from kjarni_b200 import _native as N
lib = N.lib()
M = 18944
for name, Nn, K, epi, bn in (("qkv", 1152, 384, 0, 192), ("ffn_up", 1536, 384, 1, 192), ("qkv-pair", 1152, 384, 0, 1192), ("ffn_up-pair", 1536, 384, 1, 1192), ("ffn_up-pair", 1536, 384, 1, 1256)):
    us = C.c_float()
    for fl in (8,):
        N.check(lib.kjc_dbg_gemm_time(M, Nn, K, epi, 0, bn, fl, 20, C.byref(us)))
        print(f"{name} N={Nn} K={K} BN={bn} flags={fl}: {us.value:.1f} us/launch", flush=True)
