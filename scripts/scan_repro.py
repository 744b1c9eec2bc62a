import sys
import numpy as np
sys.path.insert(0, ".")
from kjarni_b200 import api
from oracle import kjarni_oracle as ko
n, dim, nq, k = [int(x) for x in sys.argv[1:5]]
rows = ko.synth_rows(7, 0, n, dim)
q = ko.synth_rows(11, 0, nq, dim)
sh = api.IndexShard(dim, n + 3, id_base=500)
sh.add_rows(rows)
sh.set_filter(min_queries=1 << 30)
ids, sc, cnt = sh.search_batch(q, k)
wi, ws = ko.batched_topk(rows, q, k, row_offset=500)
print("ok", (ids == wi.astype(np.uint64)).mean(), np.abs(sc - ws).max())
