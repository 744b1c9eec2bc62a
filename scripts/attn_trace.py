"""Average launch time of the attention kernel through kjc_dbg_attention (KJC_ATTN_TRACE): `python scripts/attn_trace.py [B S H heads]`.
KJC_ATTN=ts|tc|legacy selects the kernel (attention_ts.cuh is the default).  The round-1 kernel (tc) additionally prints per-unit
milestone stamps when the library is built with `make -C kjarni_b200/csrc clean all EXTRA=-DKJ_ATTN_TRACE_BUILD=1`."""
import os, sys
sys.path.insert(0, ".")
os.environ["KJC_ATTN_TRACE"] = "1"
import numpy as np
from kjarni_b200 import _native as N
B, S, H, heads = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (148, 128, 384, 12)
rng = np.random.default_rng(0)
qkv = (rng.standard_normal((B * S, 3 * H)).astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)
ctx = np.empty((B * S, H), np.uint16)
N.check(N.lib().kjc_dbg_attention(qkv.ctypes.data, None, B, S, H, heads, 0, ctx.ctypes.data))
