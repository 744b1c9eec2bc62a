"""Does the second shard of the 12.5 M-row synthetic index (rows 6.25 M..12.5 M) have queries the filter cannot prove?"""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from kjarni_b200 import _native as N, api
from oracle import kjarni_oracle as ko
n, dim, k, nq = 6_250_000, 384, 10, 4096
lib = N.lib()
torch.cuda.set_stream(torch.cuda.Stream()); st = torch.cuda.current_stream().cuda_stream
for base in (0, n, 2 * n, 3 * n):
    sh = api.IndexShard(dim, n, id_base=base); sh.append_synthetic(7, base, n)
    q = torch.from_numpy(ko.synth_rows(11, 0, nq, dim)).cuda()
    ids = torch.empty((nq, k), dtype=torch.int64, device="cuda"); sc = torch.empty((nq, k), dtype=torch.float32, device="cuda"); cnt = torch.empty((nq,), dtype=torch.int32, device="cuda")
    for entry, name in ((lib.kjc_index_search_device_async, "async"), (lib.kjc_index_search_device, "exact")):
        for _ in range(2): N.check(entry(sh._h, q.data_ptr(), nq, k, 0, ids.data_ptr(), sc.data_ptr(), cnt.data_ptr(), C.c_void_p(st)))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): N.check(entry(sh._h, q.data_ptr(), nq, k, 0, ids.data_ptr(), sc.data_ptr(), cnt.data_ptr(), C.c_void_p(st)))
        e1.record(); torch.cuda.synchronize()
        print(f"shard base {base}: {name} entry {e0.elapsed_time(e1)/5:.3f} ms, launches {sh.last_launch_count}, unverified so far {sh.unverified_count}", flush=True)
    sh.close()
