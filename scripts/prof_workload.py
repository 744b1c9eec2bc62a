#!/usr/bin/env python
"""One small workload of the hot path for ncu to wrap (no timing here).
    python scripts/prof_workload.py c2|c3|c5|c1 [steps]      encoder configs (bench.py EXTRA_CONFIGS shapes, one micro-batch set)
    python scripts/prof_workload.py scan_exact|scan_gemm8|scan_gemm4096 [rows]"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kjarni_b200 import _native as N  # noqa: E402
from kjarni_b200 import api, synth  # noqa: E402

what = sys.argv[1]
lib = N.lib()
if what.startswith("scan"):
    from oracle import kjarni_oracle as ko
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 6_250_000
    dim = int(os.environ.get("DIM", "384"))
    k = int(os.environ.get("K", "10"))
    sh = api.IndexShard(dim, n)
    sh.append_synthetic(7, 0, n)
    nq = 4096 if what.endswith("4096") else 8
    if what == "scan_exact":
        N.check(lib.kjc_dbg_index_set_filter(sh._h, 0.0045, 1 << 30))
    q = ko.synth_rows(11, 0, nq, dim)
    for _ in range(3):
        ids, sc, cnt = sh.search_batch(q, k)
    print(what, n, ids[0, :4], sh.last_launch_count)
else:
    cfg = {"c1": ("minilm-l6", 148, 128, False), "c2": ("distilbert-sst2", 256, 128, False), "c3": ("minilm-l6-cross-encoder", 148, 256, True),
           "c5": ("bert-base", 74, 512, False)}[what]
    arch, B, S, pair = cfg
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    with tempfile.TemporaryDirectory() as td:
        d = synth.write_model_dir(os.path.join(td, arch), arch)
        enc = api.EncoderModel(d)
        ids, mask, types = synth.synth_tokens(B, S, enc.info.vocab_size, regime="T", seed=1, pair=pair)
        for _ in range(steps):
            if enc.num_labels:
                out = enc.predict_logits(ids, mask, types if pair else None)
            else:
                out = enc.encode_batch_from_ids(ids, mask)
        print(what, out.shape, enc.last_launch_count)
