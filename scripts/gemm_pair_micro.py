"""One-CTA vs CTA-pair form of gemm_tcgen05_kernel (kjc_dbg_gemm_time; block_n + 2000 = pair) on the encoder's stand-alone projections."""
import ctypes as C, sys
sys.path.insert(0, ".")
from kjarni_b200 import _native as N
lib = N.lib()
def t(M, Nn, K, epi, bn, iters=200):
    us = C.c_float()
    N.check(lib.kjc_dbg_gemm_time(M, Nn, K, epi, 0, bn, 0, iters, C.byref(us)))
    return us.value
for name, M, Nn, K, epi, bn in (("MiniLM QKV (layer 0) 18944 x 1152 x 384", 18944, 1152, 384, 0, 192), ("MiniLM FFN-up 18944 x 1536 x 384", 18944, 1536, 384, 1, 256),
                                ("DistilBERT QKV 32768 x 2304 x 768", 32768, 2304, 768, 0, 256), ("DistilBERT FFN-up 32768 x 3072 x 768", 32768, 3072, 768, 1, 256),
                                ("BERT-base QKV 262144 x 2304 x 768", 262144, 2304, 768, 0, 256), ("BERT-base FFN-up 262144 x 3072 x 768", 262144, 3072, 768, 1, 256),
                                ("DistilBERT QKV on 192-column tiles", 32768, 2304, 768, 0, 192)):
    a, b = t(M, Nn, K, epi, bn), t(M, Nn, K, epi, bn + 2000)
    fl = 2.0 * M * Nn * K
    print(f"{name} (BN {bn}): one CTA {a:.1f} us ({fl / a / 1e6:.0f} TFLOP/s) | CTA pair {b:.1f} us ({fl / b / 1e6:.0f} TFLOP/s)")
