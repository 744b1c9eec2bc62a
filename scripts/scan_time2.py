"""Times the always-exact device-pointer search (kjc_index_search_device) for a given DIM / K at several query-batch sizes, with
the tensor-core filter and with the exact scan alone (not a bench line).  usage: DIM=768 K=50 NQ=8,4096 python scripts/scan_time2.py rows"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from kjarni_b200 import _native as N, api
from oracle import kjarni_oracle as ko

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
dim, k = int(os.environ.get("DIM", "384")), int(os.environ.get("K", "10"))
sh = api.IndexShard(dim, n)
sh.append_synthetic(7, 0, n)
lib = N.lib()
torch.cuda.set_stream(torch.cuda.Stream())
st = torch.cuda.current_stream().cuda_stream
for nq in [int(x) for x in os.environ.get("NQ", "8,4096").split(",")]:
    q = torch.from_numpy(ko.synth_rows(11, 0, nq, dim)).cuda()
    ids = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    sc = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    cnt = torch.empty((nq,), dtype=torch.int32, device="cuda")
    res = {}
    for name, min_q in (("filter", 1), ("exact", 1 << 30)):
        if name == "exact" and nq > 256:
            continue
        sh.set_filter(min_queries=min_q)

        def step():
            N.check(lib.kjc_index_search_device(sh._h, q.data_ptr(), nq, k, 0, ids.data_ptr(), sc.data_ptr(), cnt.data_ptr(), C.c_void_p(st)))
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        res[name] = (e0.elapsed_time(e1) / reps, sh.last_launch_count)
    print(f"nq={nq:5d}: " + "  ".join(f"{nm} {ms:8.3f} ms ({nq / ms * 1e3:9.0f} q/s, {ln} launches)" for nm, (ms, ln) in res.items()), flush=True)
