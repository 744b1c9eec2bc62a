"""Escalation of unproven queries (32 -> 128 candidates from the same filter buffer) vs going straight to the exact scan."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from kjarni_b200 import api
from oracle import kjarni_oracle as ko
n, dim, nq, k = 400_000, 384, 1024, 10
rows = None
q = ko.synth_rows(11, 0, nq, dim)
for eps in (0.0045, 0.008, 0.012, 0.02):
    out = {}
    for esc in (1, 0):
        if esc: os.environ.pop("KJC_SCAN_NO_ESCALATE", None)
        else: os.environ["KJC_SCAN_NO_ESCALATE"] = "1"
        sh = api.IndexShard(dim, n); sh.append_synthetic(7, 0, n)
        sh.set_filter(eps=eps, min_queries=1)
        sh.search_batch(q, k)
        t0 = time.perf_counter()
        ids, sc, cnt = sh.search_batch(q, k)
        dt = (time.perf_counter() - t0) * 1e3
        out[esc] = (ids, sc, sh.last_launch_count, dt)
        sh.close()
    same = np.array_equal(out[1][0], out[0][0]) and np.array_equal(out[1][1], out[0][1])
    print(f"eps {eps}: escalate: {out[1][2]} launches {out[1][3]:.2f} ms | straight to exact: {out[0][2]} launches {out[0][3]:.2f} ms | identical {same}", flush=True)
