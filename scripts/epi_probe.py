"""Epilogue decomposition under load (kjc_dbg_gemm_time flags): 16 no TMA store, 32 no proxy fence, 128 no tcgen05.ld, 256 no st.shared, 512 no bias/act math."""
import ctypes as C, sys
sys.path.insert(0, ".")
from kjarni_b200 import _native as N
lib = N.lib()
M = 18944
def t(Nn, K, epi, bn, flags):
    us = C.c_float()
    N.check(lib.kjc_dbg_gemm_time(M, Nn, K, epi, 0, bn, flags, 30, C.byref(us)))
    return us.value
for name, Nn, K, epi in (("qkv", 1152, 384, 0), ("ffn_up", 1536, 384, 1)):
    for fl, nm in ((0, "full"), (16, "no-store"), (48, "no-store,no-fence"), (128, "no-ldtm"), (256, "no-sts"), (512, "no-math"), (16 + 32 + 256, "ldtm+math only"),
                   (16 + 32 + 128 + 256 + 512, "nothing"), (1, "no-epi")):
        print(f"{name} BN=192 {nm:20s}: {t(Nn, K, epi, 192, fl):6.1f} us", flush=True)
