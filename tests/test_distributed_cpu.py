"""world_size-2 gloo test (CPU) of the N>1 host logic: batch split, contiguous row sharding with global ids,
candidate all-gather and the merge order -- against the oracle's single-index search."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kjarni_b200 import distributed as kd
from oracle import kjarni_oracle as ko


def test_split_helpers_cover_everything_once():
    for n, w in ((10, 3), (8, 8), (5, 8), (50_000_000, 8), (1, 2)):
        parts = [kd.split_batch(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        for (s0, c0), (s1, _) in zip(parts, parts[1:]):
            assert s0 + c0 == s1
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    assert kd.shard_rows(50_000_000, 8, 3) == (18_750_000, 6_250_000)


class _CpuShardIndex(kd.DistributedIndex):
    """Same exchange code path; the two device steps are replaced by the oracle so the test runs without a GPU."""

    def __init__(self, rows, row0, group=None):
        self.rows, self.row0 = rows, row0
        self.group = group
        self.world = dist.get_world_size(group)

    def _search_local(self, queries, k, mode, rec):
        nq = queries.shape[0]
        ids, sc = ko.batched_topk(self.rows, queries.numpy(), k, row_offset=self.row0)
        buf = rec.numpy()  # the packed record: [nq,k] ids | [nq,k] scores
        buf[:nq * k * 8] = np.ascontiguousarray(ids, np.int64).view(np.uint8).reshape(-1)
        buf[nq * k * 8:nq * k * 12] = np.ascontiguousarray(sc, np.float32).view(np.uint8).reshape(-1)

    def _merge(self, gathered, nq, k):
        g = gathered.numpy()
        world = g.shape[0]
        g_ids = np.stack([g[r, :nq * k * 8].view(np.int64).reshape(nq, k) for r in range(world)])
        g_sc = np.stack([g[r, nq * k * 8:nq * k * 12].view(np.float32).reshape(nq, k) for r in range(world)])
        ids = g_ids.transpose(1, 0, 2).reshape(nq, -1)
        sc = g_sc.transpose(1, 0, 2).reshape(nq, -1)
        out_i = np.full((nq, k), -1, np.int64)
        out_s = np.full((nq, k), -np.inf, np.float32)
        for q in range(nq):
            valid = ids[q] >= 0
            order = np.lexsort((ids[q][valid], -sc[q][valid]))[:k]  # (score desc, id asc)
            out_i[q, :len(order)] = ids[q][valid][order]
            out_s[q, :len(order)] = sc[q][valid][order]
        return torch.from_numpy(out_i), torch.from_numpy(out_s)


def _worker(rank, world, port, n_total, dim, nq, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        row0, n = kd.shard_rows(n_total, world, rank)
        rows = ko.synth_rows(7, row0, n, dim)
        if rank == 1:
            rows[3] = ko.synth_rows(7, 5, 1, dim)[0]  # exact duplicate of global row 5 living on another shard
        q = torch.from_numpy(ko.synth_rows(11, 0, nq, dim))
        q[0] = torch.from_numpy(ko.synth_rows(7, 5, 1, dim)[0])
        ids, sc = _CpuShardIndex(rows, row0).search(q, k)
        ret[rank] = (ids.numpy().copy(), sc.numpy().copy())
    finally:
        dist.destroy_process_group()


def test_row_sharded_search_matches_single_index():
    world, n_total, dim, nq, k = 2, 3001, 64, 5, 10
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_total, dim, nq, k, ret), nprocs=world, join=True)
    full = ko.synth_rows(7, 0, n_total, dim)
    row0_1, _ = kd.shard_rows(n_total, world, 1)
    full[row0_1 + 3] = full[5]
    q = ko.synth_rows(11, 0, nq, dim)
    q[0] = full[5]
    want_i, want_s = ko.batched_topk(full, q, k)
    for r in range(world):
        ids, sc = ret[r]
        assert np.array_equal(ids, want_i), r  # identical on every rank, tie (rows 5 and row0_1+3) -> lower global id first
        assert np.allclose(sc, want_s, atol=1e-6)
    assert list(want_i[0, :2]) == [5, row0_1 + 3]
