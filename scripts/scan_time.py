"""Times the index search through the device-pointer C ABI at several query-batch sizes (not a bench line)."""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from kjarni_b200 import _native as N, api
from oracle import kjarni_oracle as ko

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6_250_000
dim, k = 384, 10
sh = api.IndexShard(dim, n)
sh.append_synthetic(7, 0, n)
lib = N.lib()
torch.cuda.set_stream(torch.cuda.Stream())  # NULL would mean "the index's own stream" to the library
st = torch.cuda.current_stream().cuda_stream
import os
for nq in [int(x) for x in os.environ.get('NQ', '8,16,128,256,1024,4096').split(',')]:
    q = torch.from_numpy(ko.synth_rows(11, 0, nq, dim)).cuda()
    ids = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    sc = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    cnt = torch.empty((nq,), dtype=torch.int32, device="cuda")
    def step():
        N.check(lib.kjc_index_search_device_async(sh._h, q.data_ptr(), nq, k, 0, ids.data_ptr(), sc.data_ptr(), cnt.data_ptr(), C.c_void_p(st)))
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"nq={nq:5d}: {ms:8.3f} ms  {nq/ms*1e3:10.0f} q/s  launches={sh.last_launch_count}  unverified={sh.unverified_count}", flush=True)
