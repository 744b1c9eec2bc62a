"""Milestone trace of the tcgen05 attention kernel (KJC_ATTN_TRACE through kjc_dbg_attention): where a softmax warpgroup's time goes.
The per-unit stamps need a library built with `make -C kjarni_b200/csrc clean all EXTRA=-DKJ_ATTN_TRACE_BUILD=1`; without it only the
average launch time is printed."""
import os, sys
sys.path.insert(0, ".")
os.environ["KJC_ATTN_TRACE"] = "1"
import numpy as np
from kjarni_b200 import _native as N
B, S, H, heads = 148, 128, 384, 12
rng = np.random.default_rng(0)
qkv = (rng.standard_normal((B * S, 3 * H)).astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)
ctx = np.empty((B * S, H), np.uint16)
N.check(N.lib().kjc_dbg_attention(qkv.ctypes.data, None, B, S, H, heads, 0, ctx.ctypes.data))
