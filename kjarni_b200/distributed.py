"""One process per GPU over torch.distributed: how the two halves of the hot path shard.

* Encoder (Embedder / Classifier / Reranker): sequences are independent -> the batch is split by
  contiguous rows, every rank holds a full copy of the weights, NO collective.
* Cosine top-k scan: the index is row-sharded in contiguous blocks (global id = shard base + local id,
  as IndexReader::local_to_global, kjarni-rag/src/index_reader.rs:313-319); every rank scans its shard
  for the same query batch, the per-shard [Q,k] candidates are all-gathered (NCCL over NVLink on GPU
  tensors) and merged by the library's merge kernel -- the per-segment top-k -> concat -> sort -> truncate
  of IndexReader::search_semantic (index_reader.rs:207-228).

torch is plumbing here (process group, device buffers); all numeric work is in libkjarni_cuda.so.
"""
from __future__ import annotations

from typing import Optional, Tuple

from . import _native as N


def split_batch(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [start, start+count) of a batch for `rank` (first `batch % world` ranks get one extra)."""
    base, rem = divmod(batch, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def shard_rows(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block (row0, n) of an n_total-row index for `rank`; row0 is the shard's id_base."""
    return split_batch(n_total, world, rank)


class DistributedIndex:
    """Row-sharded index: `shard` is this rank's kjarni_b200.api.IndexShard (id_base = its first global row)."""

    def __init__(self, shard, group=None):
        import torch.distributed as dist

        self.shard = shard
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    @classmethod
    def open_dir(cls, root: str, device: int, group=None) -> "DistributedIndex":
        """Every rank uploads its contiguous part of an on-disk index directory (kjc_index_open_dir: IndexReader::open,
        kjarni-rag/src/index_reader.rs:161-204, split by rows as `shard_rows` does); global ids stay those of the host's
        IndexReader, so rank-local results merge without remapping."""
        import torch.distributed as dist

        from . import api

        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        return cls(api.IndexShard.open_dir(root, device, rank, world), group)

    # -- the two device-side steps, overridable so the exchange logic can be exercised on CPU (gloo) in tests
    def _search_local(self, queries, k: int, mode: int):
        import torch

        nq = queries.shape[0]
        ids = torch.empty((nq, k), dtype=torch.int64, device=queries.device)
        sc = torch.empty((nq, k), dtype=torch.float32, device=queries.device)
        stream = torch.cuda.current_stream().cuda_stream
        # the synchronising entry: queries the tensor-core filter cannot prove are re-run on the exact scan, so the merged result is
        # the exact top-k in every case (IndexReader::search_semantic is exact)
        N.check(N.lib().kjc_index_search_device(self.shard._h, queries.data_ptr(), nq, k, mode, ids.data_ptr(), sc.data_ptr(),
                                                None, stream if stream else None))
        return ids, sc

    def _merge(self, g_ids, g_sc, nq: int, k: int):
        import torch

        ids = torch.empty((nq, k), dtype=torch.int64, device=g_ids.device)
        sc = torch.empty((nq, k), dtype=torch.float32, device=g_ids.device)
        stream = torch.cuda.current_stream().cuda_stream
        N.check(N.lib().kjc_topk_merge_device_async(self.shard.device, g_ids.data_ptr(), g_sc.data_ptr(), g_ids.shape[0], nq, k,
                                                    ids.data_ptr(), sc.data_ptr(), None, stream if stream else None))
        return ids, sc

    def search(self, queries, k: int, mode: int = N.SCAN_SEGMENT):
        """queries: [Q, dim] float32 tensor on this rank's device (identical on every rank).
        Returns (ids int64 [Q,k] global, -1 = empty; scores f32 [Q,k]) -- identical on every rank."""
        import torch
        import torch.distributed as dist

        nq = queries.shape[0]
        ids, sc = self._search_local(queries, k, mode)
        if self.world == 1:
            return ids, sc
        g_ids = torch.empty((self.world * nq, k), dtype=torch.int64, device=ids.device)  # rank-major concat
        g_sc = torch.empty((self.world * nq, k), dtype=torch.float32, device=ids.device)
        dist.all_gather_into_tensor(g_ids, ids.contiguous(), group=self.group)
        dist.all_gather_into_tensor(g_sc, sc.contiguous(), group=self.group)
        return self._merge(g_ids.view(self.world, nq, k), g_sc.view(self.world, nq, k), nq, k)
