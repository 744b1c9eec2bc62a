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
    @staticmethod
    def record_bytes(nq: int, k: int) -> int:
        """One rank's packed candidate record: [nq,k] u64 ids | [nq,k] f32 scores, rounded up to 16 bytes (kjc_packed_record_bytes)."""
        return (nq * k * 12 + 15) & ~15

    def _search_local(self, queries, k: int, mode: int, rec):
        """Per-shard exact top-k written straight into the packed record `rec` (uint8 [record_bytes])."""
        nq = queries.shape[0]
        import torch

        stream = torch.cuda.current_stream().cuda_stream
        # the synchronising entry: queries the tensor-core filter cannot prove are re-run on the exact scan, so the merged result is
        # the exact top-k in every case (IndexReader::search_semantic is exact)
        N.check(N.lib().kjc_index_search_device(self.shard._h, queries.data_ptr(), nq, k, mode, rec.data_ptr(), rec.data_ptr() + nq * k * 8,
                                                None, stream if stream else None))

    def _merge(self, gathered, nq: int, k: int):
        """gathered: uint8 [world, record_bytes] -> (ids int64 [nq,k], scores f32 [nq,k])."""
        import torch

        ids = torch.empty((nq, k), dtype=torch.int64, device=gathered.device)
        sc = torch.empty((nq, k), dtype=torch.float32, device=gathered.device)
        stream = torch.cuda.current_stream().cuda_stream
        N.check(N.lib().kjc_topk_merge_packed_device_async(self.shard.device, gathered.data_ptr(), gathered.shape[0], nq, k,
                                                           ids.data_ptr(), sc.data_ptr(), None, stream if stream else None))
        return ids, sc

    def search(self, queries, k: int, mode: int = N.SCAN_SEGMENT):
        """queries: [Q, dim] float32 tensor on this rank's device (identical on every rank).
        Returns (ids int64 [Q,k] global, -1 = empty; scores f32 [Q,k]) -- identical on every rank.
        ONE collective per search: every rank's ids and scores travel as one packed record."""
        import torch
        import torch.distributed as dist

        nq = queries.shape[0]
        rb = self.record_bytes(nq, k)
        rec = torch.empty((rb,), dtype=torch.uint8, device=queries.device)
        self._search_local(queries, k, mode, rec)
        if self.world == 1:
            gathered = rec.view(1, rb)
        else:
            flat = torch.empty((self.world * rb,), dtype=torch.uint8, device=queries.device)  # rank-major concat
            dist.all_gather_into_tensor(flat, rec, group=self.group)
            gathered = flat.view(self.world, rb)
        return self._merge(gathered, nq, k)
