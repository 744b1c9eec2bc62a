"""kjarni_b200 -- B200 (sm_100a) backend for Kjarni's encoder forward + cosine top-k scan.

The product is `libkjarni_cuda.so` (C ABI in include/kjarni_cuda.h); `api` mirrors the
reference's operator interface over it via ctypes.  `synth` builds random-init model
directories / synthetic inputs and does not need the native library.
"""
__all__ = ["api", "synth", "_native"]
