#!/usr/bin/env python
"""Batch-32 x 128 latency case (BASELINE configs[0]) under the kernel-selection switches: which launch structure is fastest when a
micro-batch holds far fewer 128-row tiles than the GPU has SMs.  usage: [KJC_...=..] python scripts/c1_latency_ab.py [batch ...]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from kjarni_b200 import _native as N  # noqa: E402
from kjarni_b200 import api, synth  # noqa: E402

lib = N.lib()
torch.cuda.set_stream(torch.cuda.Stream())  # stream 0 would mean "the encoder's own stream", which torch events do not see
enc = api.EncoderModel(bench._model_dir("minilm-l6", 0), device=0)
for B in [int(a) for a in sys.argv[1:]] or [32]:
    S = 128
    ids, mask, _ = synth.synth_tokens(B, S, enc.info.vocab_size, regime="T", seed=4242)
    ids_d = torch.from_numpy(ids.view(np.int32)).cuda()
    mask_d = torch.from_numpy(mask.astype(np.float32)).cuda()
    out_d = torch.empty((B, enc.info.hidden_size), dtype=torch.float32, device="cuda")
    opts = N.KjcForwardOptions(N.OUT_POOLED, N.POOL_MEAN, 1, N.MASK_AUTO)
    st = torch.cuda.current_stream().cuda_stream

    def step():
        N.check(lib.kjc_encoder_forward_device_async(enc._h, ids_d.data_ptr(), mask_d.data_ptr(), None, B, S, C.byref(opts), out_d.data_ptr(), st))

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 1000
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    # one isolated call (launch -> result ready), median of 50
    lat = []
    for _ in range(50):
        torch.cuda.synchronize()
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        lat.append(e0.elapsed_time(e1))
    env = {k: v for k, v in os.environ.items() if k.startswith("KJC_")}
    print(f"B={B} {env} back-to-back {ms * 1e3:.1f} us/step ({B / ms * 1e3:.0f} emb/s), isolated {np.median(lat) * 1e3:.1f} us, launches {enc.last_launch_count}, checksum {float(out_d.double().sum()):.6f}")
