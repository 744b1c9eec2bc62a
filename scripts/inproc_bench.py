#!/usr/bin/env python
"""In-process multi-GPU measurement through the C ABI alone (no torch, no torchrun): kjc_encoder_create_multi +
kjc_sharded_index_* on N GPUs of one box.  One host thread per GPU inside the library; everything is timed end to end with
HOST buffers (wall clock around synchronous C-ABI calls).
    python scripts/inproc_bench.py --gpus 2 [--steps 5] [--index-rows 6250000]"""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kjarni_b200 import _native as N  # noqa: E402
from kjarni_b200 import api, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--index-rows", type=int, default=6_250_000, help="rows per GPU")
    ap.add_argument("--dup", action="store_true", help="list device 0 N times (one-GPU box)")
    a = ap.parse_args()
    lib = N.lib()
    ndev = lib.kjc_device_count()
    devs = [0] * a.gpus if a.dup else list(range(a.gpus))
    assert a.dup or ndev >= a.gpus, (ndev, a.gpus)
    out = {"n_gpus": a.gpus, "devices": devs, "parallelism": "one process, one host thread per GPU inside libkjarni_cuda.so"}
    with tempfile.TemporaryDirectory() as td:
        d = synth.write_model_dir(os.path.join(td, "minilm-l6"), "minilm-l6")
        for n in sorted({1, a.gpus}):
            enc = api.EncoderModel(d, devices=devs[:n])
            B = n * 28 * enc.micro_batch(128)
            ids, mask, _ = synth.synth_tokens(B, 128, enc.info.vocab_size, regime="T", seed=42)
            maskf = mask.astype(np.float32)
            o = np.empty((B, enc.hidden_size), np.float32)
            opts = N.KjcForwardOptions(N.OUT_POOLED, N.POOL_MEAN, 1, N.MASK_AUTO)
            for _ in range(3):
                N.check(lib.kjc_encoder_forward(enc._h, ids.ctypes.data, maskf.ctypes.data, None, B, 128, C.byref(opts), o.ctypes.data))
            t0 = time.perf_counter()
            for _ in range(a.steps):
                N.check(lib.kjc_encoder_forward(enc._h, ids.ctypes.data, maskf.ctypes.data, None, B, 128, C.byref(opts), o.ctypes.data))
            dt = time.perf_counter() - t0
            out[f"encoder_e2e_{n}gpu"] = {"value": B * a.steps / dt, "unit": "embeddings/s", "batch": B, "ms_per_step": dt / a.steps * 1e3}
            enc.close()
    for n in sorted({1, a.gpus}):
        total = a.index_rows * n
        sh = api.ShardedIndex(384, total, devs[:n])
        sh.append_synthetic(7, total)
        from oracle import kjarni_oracle as ko  # query rows only (outside the timed region)
        for nq in (4096, 8):
            q = ko.synth_rows(11, 0, nq, 384)
            for _ in range(2):
                sh.search_batch(q, 10)
            t0 = time.perf_counter()
            reps = a.steps if nq > 8 else 20 * a.steps
            for _ in range(reps):
                ids, sc, cnt = sh.search_batch(q, 10)
            dt = time.perf_counter() - t0
            out[f"index_top10_{nq}q_{n}gpu"] = {"value": nq * reps / dt, "unit": "queries/s", "rows_total": total, "ms_per_step": dt / reps * 1e3}
        sh.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
