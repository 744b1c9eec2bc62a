"""Replicates bench.py's index regimes in order (4096-query filter, 8-query filter, 8-query exact) to compare with stand-alone timing."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, ".")
from kjarni_b200 import _native as N, api
from oracle import kjarni_oracle as ko
n, dim, k = 6_250_000, 384, 10
sh = api.IndexShard(dim, n); sh.append_synthetic(7, 0, n)
lib = N.lib()
torch.cuda.set_stream(torch.cuda.Stream()); st = torch.cuda.current_stream().cuda_stream
def run(nq, exact, reps, sync_entry=True, tag=""):
    sh.set_filter(min_queries=(1 << 30) if exact else 1)
    q = torch.from_numpy(ko.synth_rows(11, 0, nq, dim)).cuda()
    ids = torch.empty((nq, k), dtype=torch.int64, device="cuda"); sc = torch.empty((nq, k), dtype=torch.float32, device="cuda"); cnt = torch.empty((nq,), dtype=torch.int32, device="cuda")
    fn = lib.kjc_index_search_device if sync_entry else lib.kjc_index_search_device_async
    def step(): N.check(fn(sh._h, q.data_ptr(), nq, k, 0, ids.data_ptr(), sc.data_ptr(), cnt.data_ptr(), C.c_void_p(st)))
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): step()
    e1.record(); torch.cuda.synchronize()
    print(f"{tag} nq={nq} exact={exact} reps={reps} sync_entry={sync_entry}: {e0.elapsed_time(e1)/reps:.3f} ms", flush=True)
if os.environ.get("ORDER", "bench") == "bench":
    run(4096, False, 10, tag="A"); run(8, False, 10, tag="B"); run(8, True, 10, tag="C"); run(8, True, 5, False, tag="D"); run(8, True, 50, True, tag="E")
else:
    run(8, True, 5, False, tag="D0"); run(8, True, 10, True, tag="C0"); run(4096, False, 10, tag="A"); run(8, True, 10, True, tag="C1")
