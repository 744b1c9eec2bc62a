"""One 4096-query search over a 6.25 M x 384 shard, for an ncu launch list (which kernel takes what at full size)."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from kjarni_b200 import _native as N, api
from oracle import kjarni_oracle as ko
n, dim, k, nq = 6_250_000, 384, 10, 4096
sh = api.IndexShard(dim, n); sh.append_synthetic(7, 0, n)
lib = N.lib()
st = torch.cuda.current_stream().cuda_stream
q = torch.from_numpy(ko.synth_rows(11, 0, nq, dim)).cuda()
ids = torch.empty((nq, k), dtype=torch.int64, device="cuda"); sc = torch.empty((nq, k), dtype=torch.float32, device="cuda"); cnt = torch.empty((nq,), dtype=torch.int32, device="cuda")
for _ in range(3):
    N.check(lib.kjc_index_search_device(sh._h, q.data_ptr(), nq, k, 0, ids.data_ptr(), sc.data_ptr(), cnt.data_ptr(), C.c_void_p(st)))
torch.cuda.synchronize()
