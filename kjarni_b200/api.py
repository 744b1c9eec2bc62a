"""Host-side mirror of the reference's operator interface for the hot path, over the C ABI.

Same names and argument meaning as the Rust API the CUDA backend sits behind
(citations into /root/reference/crates; KT = kjarni-transformers/src, KM = kjarni-models/src,
KS = kjarni-search/src, KR = kjarni-rag/src):

  EncoderModel.get_hidden_states_batch_from_ids   KT/cpu/encoder/traits.rs:66-139
  EncoderModel.encode_batch_from_ids              KT/cpu/encoder/traits.rs:204-225 (SentenceEncoder::encode_batch on ids)
  EncoderModel.predict_logits                     KM/models/sequence_classifier/mod.rs:266-346
  EncoderModel.classify_scores_batch              KM/models/sequence_classifier/mod.rs:248-263
  EncoderModel.predict_pairs / rerank             KM/models/cross_encoder/model.rs:170-255
  Segment.search_vectors / get_embedding          KR/segment.rs:240-262,307-337
  VectorStore.search                              KS/vector.rs:150-165
  IndexReader.search_semantic                     KR/index_reader.rs:207-228

Everything numeric happens inside libkjarni_cuda.so; numpy is only the container for
host buffers.  No torch import here.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


class EncoderModel:
    """One encoder checkpoint resident on one GPU, or one replica per GPU of `devices` (KjcEncoder; a host batch is then split by
    contiguous rows over the replicas inside the library, no collective)."""

    def __init__(self, model_dir: str, device: int = 0, devices: Optional[Sequence[int]] = None):
        self._h = C.c_void_p()
        if devices is None:
            N.check(N.lib().kjc_encoder_create(str(model_dir).encode(), int(device), C.byref(self._h)))
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            N.check(N.lib().kjc_encoder_create_multi(str(model_dir).encode(), arr, len(devices), C.byref(self._h)))
        self.n_devices = int(N.lib().kjc_encoder_device_count(self._h))
        info = N.KjcEncoderInfo()
        N.check(N.lib().kjc_encoder_info(self._h, C.byref(info)))
        self.info = info
        self.arch = N.ARCH_NAMES[info.arch]
        self.head_kind = N.HEAD_NAMES[info.head_kind]
        self.hidden_size = info.hidden_size
        self.num_labels = info.num_labels
        self.device = info.device
        self.labels: List[str] = []
        i = 0
        while True:
            s = N.lib().kjc_encoder_label(self._h, i)
            if s is None:
                break
            self.labels.append(s.decode())
            i += 1

    def set_fp32_residual(self, mode: int) -> None:
        """0 = bf16 residual stream everywhere, 1 = fp32 for hidden-state output (default), 2 = fp32 for every output."""
        N.check(N.lib().kjc_encoder_set_fp32_residual(self._h, int(mode)))

    @classmethod
    def from_pretrained(cls, model_dir: str, device: int = 0) -> "EncoderModel":
        return cls(model_dir, device)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().kjc_encoder_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ core
    def _forward(self, ids, mask, type_ids, output, pooling=N.POOL_MEAN, normalize=True, mask_convention=N.MASK_AUTO) -> np.ndarray:
        ids = _c(ids, np.uint32)
        if ids.ndim != 2:
            raise ValueError("input_ids must be [batch, seq]")
        B, S = ids.shape
        maskf = None if mask is None else _c(mask, np.float32)
        types = None if type_ids is None else _c(type_ids, np.uint32)
        if maskf is not None and maskf.shape != ids.shape:
            raise ValueError("attention_mask shape mismatch")
        if types is not None and types.shape != ids.shape:
            raise ValueError("token_type_ids shape mismatch")
        if output == N.OUT_HIDDEN:
            out = np.empty((B, S, self.hidden_size), np.float32)
        elif output == N.OUT_POOLED:
            out = np.empty((B, self.hidden_size), np.float32)
        else:
            out = np.empty((B, max(self.num_labels, 0)), np.float32)
        opts = N.KjcForwardOptions(output, pooling, 1 if normalize else 0, mask_convention)
        N.check(N.lib().kjc_encoder_forward(self._h, _ptr(ids), _ptr(maskf), _ptr(types), B, S, C.byref(opts), _ptr(out)))
        return out

    def get_hidden_states_batch_from_ids(self, input_ids, attention_mask, mask_convention=N.MASK_AUTO) -> np.ndarray:
        return self._forward(input_ids, attention_mask, None, N.OUT_HIDDEN, mask_convention=mask_convention)

    def forward_tokens(self, input_ids, attention_mask, token_type_ids=None, mask_convention=N.MASK_ALLOC) -> np.ndarray:
        """KT/cpu/encoder/traits.rs:295-313 (embed with type ids -> embed LN -> encoder.forward, alloc path)."""
        return self._forward(input_ids, attention_mask, token_type_ids, N.OUT_HIDDEN, mask_convention=mask_convention)

    def encode_batch_from_ids(self, input_ids, attention_mask, pooling: str = "mean", normalize: bool = True) -> np.ndarray:
        pool = {"mean": N.POOL_MEAN, "cls": N.POOL_CLS, "max": N.POOL_MAX, "last": N.POOL_LAST}[pooling]
        return self._forward(input_ids, attention_mask, None, N.OUT_POOLED, pool, normalize)

    def predict_logits(self, input_ids, attention_mask, token_type_ids=None) -> np.ndarray:
        if self.info.type_vocab_size > 0 and token_type_ids is None:
            token_type_ids = np.zeros_like(_c(input_ids, np.uint32))
        return self._forward(input_ids, attention_mask, token_type_ids, N.OUT_LOGITS)

    def classify_scores_batch(self, input_ids, attention_mask, token_type_ids=None) -> np.ndarray:
        logits = self.predict_logits(input_ids, attention_mask, token_type_ids)
        N.lib().kjc_softmax_rows(_ptr(logits), logits.shape[0], logits.shape[1])
        return logits

    def classify_multi_label(self, input_ids, attention_mask, token_type_ids=None) -> np.ndarray:
        """ClassificationMode::MultiLabel: raw logits -> sigmoid per label (kjarni/src/classifier/model.rs:313-335,528-531)."""
        logits = self.predict_logits(input_ids, attention_mask, token_type_ids)
        N.lib().kjc_sigmoid_rows(_ptr(logits), logits.shape[0], logits.shape[1])
        return logits

    def predict_pairs(self, input_ids, attention_mask, token_type_ids) -> np.ndarray:
        """Cross-encoder relevance scores = logits[:, 0] (raw) (KM/models/cross_encoder/model.rs:239)."""
        return self.predict_logits(input_ids, attention_mask, token_type_ids)[:, 0].copy()

    def rerank(self, input_ids, attention_mask, token_type_ids) -> List[Tuple[int, float]]:
        """Stable sort by score descending (KM/models/cross_encoder/model.rs:243-255)."""
        s = self.predict_pairs(input_ids, attention_mask, token_type_ids)
        order = np.argsort(-s, kind="stable")
        return [(int(i), float(s[i])) for i in order]

    def micro_batch(self, seq_len: int) -> int:
        return int(N.lib().kjc_encoder_micro_batch(self._h, int(seq_len)))

    @property
    def last_launch_count(self) -> int:
        return int(N.lib().kjc_encoder_last_launch_count(self._h))


class Tokenizer:
    """tokenizer.json -> ids with the reference's settings (KjcTokenizer; KT/pipeline/encoder/loader.rs:99-115):
    truncation to `max_length`, BatchLongest padding with id 0, special tokens from the file's post-processor."""

    def __init__(self, tokenizer_json: str, max_length: int = 512):
        self._h = C.c_void_p()
        N.check(N.lib().kjc_tokenizer_create(str(tokenizer_json).encode(), int(max_length), C.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().kjc_tokenizer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def token_to_id(self, token: str) -> Optional[int]:
        out = C.c_uint32()
        return int(out.value) if N.lib().kjc_tokenizer_token_to_id(self._h, token.encode(), C.byref(out)) else None

    def encode_batch(self, texts: Sequence[str], pairs: Optional[Sequence[str]] = None, add_special_tokens: bool = True):
        """(ids u32 [n,S], attention_mask f32 [n,S], type_ids u32 [n,S])"""
        n = len(texts)
        ta = (C.c_char_p * max(n, 1))(*[t.encode("utf-8") for t in texts])
        tb = None if pairs is None else (C.c_char_p * max(n, 1))(*[t.encode("utf-8") for t in pairs])
        S = C.c_int()
        N.check(N.lib().kjc_tokenizer_encode_batch(self._h, ta, tb, n, 1 if add_special_tokens else 0, None, None, None, 0, C.byref(S)))
        s = max(S.value, 1)
        ids = np.zeros((n, s), np.uint32)
        mask = np.zeros((n, s), np.float32)
        types = np.zeros((n, s), np.uint32)
        N.check(N.lib().kjc_tokenizer_encode_batch(self._h, ta, tb, n, 1 if add_special_tokens else 0, _ptr(ids), _ptr(mask), _ptr(types), s, C.byref(S)))
        return ids[:, :S.value], mask[:, :S.value], types[:, :S.value]


def scores_to_top_k(probs: np.ndarray, labels: Sequence[str], k: int) -> List[Tuple[str, float]]:
    """Stable sort descending => ties resolve to the lowest label index (KM/models/sequence_classifier/mod.rs:369-390)."""
    order = np.argsort(-np.asarray(probs, np.float32), kind="stable")[:k]
    return [(labels[i] if i < len(labels) else f"LABEL_{i}", float(probs[i])) for i in order]


class ShardedIndex:
    """The whole index row-sharded over the GPUs of `devices` inside this process (KjcShardedIndex): per-shard exact top-k, peer-copy
    candidate gather on devices[0], merge kernel -- IndexReader::search_semantic over all rows (KR/index_reader.rs:207-228)."""

    def __init__(self, dim: int, capacity_rows: int, devices: Sequence[int], mode: int = N.SCAN_SEGMENT):
        self._h = C.c_void_p()
        self.mode = mode
        self.dim = dim
        arr = (C.c_int * len(devices))(*[int(d) for d in devices])
        N.check(N.lib().kjc_sharded_index_create(int(dim), int(capacity_rows), arr, len(devices), C.byref(self._h)))

    @classmethod
    def open_dir(cls, root: str, devices: Sequence[int]) -> "ShardedIndex":
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        self.mode = N.SCAN_SEGMENT
        arr = (C.c_int * len(devices))(*[int(d) for d in devices])
        N.check(N.lib().kjc_sharded_index_open_dir(str(root).encode(), arr, len(devices), C.byref(self._h)))
        self.dim = int(N.lib().kjc_sharded_index_dim(self._h))
        return self

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().kjc_sharded_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(N.lib().kjc_sharded_index_len(self._h))

    @property
    def shard_lens(self) -> List[int]:
        return [int(N.lib().kjc_sharded_index_shard_len(self._h, p)) for p in range(int(N.lib().kjc_sharded_index_shards(self._h)))]

    def add_rows(self, rows) -> None:
        rows = _c(rows, np.float32)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise ValueError("rows must be [n, dim]")
        N.check(N.lib().kjc_sharded_index_add_rows(self._h, _ptr(rows), rows.shape[0]))

    def append_synthetic(self, seed: int, n: int) -> None:
        N.check(N.lib().kjc_sharded_index_append_synthetic(self._h, int(seed), int(n)))

    def search_batch(self, queries, k: int, mode: Optional[int] = None):
        """(ids u64 [Q,k] global, scores f32 [Q,k], counts i32 [Q]); order (score desc, id asc)."""
        q = _c(queries, np.float32)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError("queries must be [nq, dim]")
        nq = q.shape[0]
        ids = np.empty((nq, k), np.uint64)
        sc = np.empty((nq, k), np.float32)
        cnt = np.empty((nq,), np.int32)
        N.check(N.lib().kjc_sharded_index_search(self._h, _ptr(q), nq, int(k), self.mode if mode is None else mode, _ptr(ids), _ptr(sc), _ptr(cnt)))
        return ids, sc, cnt

    @property
    def last_launches(self) -> int:
        return int(N.lib().kjc_sharded_index_last_launch_count(self._h))


class IndexShard:
    """A contiguous block of index rows resident in HBM on one GPU (KjcIndex)."""

    def __init__(self, dim: int, capacity_rows: int, id_base: int = 0, device: int = 0, mode: int = N.SCAN_SEGMENT):
        self._h = C.c_void_p()
        self.mode = mode
        self.dim = dim
        self.id_base = id_base
        self.device = device
        N.check(N.lib().kjc_index_create(int(dim), int(capacity_rows), int(id_base), int(device), C.byref(self._h)))

    @classmethod
    def open_dir(cls, root: str, device: int = 0, part: int = 0, parts: int = 1) -> "IndexShard":
        """Rows of part `part` of `parts` of an on-disk index directory (kjc_index_open_dir; IndexReader::open,
        KR/index_reader.rs:161-204): all segments in file-name order, contiguous global-id range, id_base = first row."""
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        self.mode = N.SCAN_SEGMENT
        self.device = device
        N.check(N.lib().kjc_index_open_dir(str(root).encode(), int(device), int(part), int(parts), C.byref(self._h)))
        self.dim = int(N.lib().kjc_index_dim(self._h))
        self.id_base = int(N.lib().kjc_index_id_base(self._h))
        return self

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().kjc_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(N.lib().kjc_index_len(self._h))

    def add_rows(self, rows) -> None:
        rows = _c(rows, np.float32)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise ValueError("rows must be [n, dim]")
        N.check(N.lib().kjc_index_add_rows(self._h, _ptr(rows), rows.shape[0]))

    def load_vectors_bin(self, path: str) -> None:
        N.check(N.lib().kjc_index_load_vectors_bin(self._h, str(path).encode()))

    def append_synthetic(self, seed: int, row0: int, n: int) -> None:
        N.check(N.lib().kjc_index_append_synthetic(self._h, int(seed), int(row0), int(n)))

    def get_embedding(self, doc_id: int) -> np.ndarray:
        out = np.empty((self.dim,), np.float32)
        N.check(N.lib().kjc_index_get_rows(self._h, int(doc_id), 1, _ptr(out)))
        return out

    def get_rows(self, row: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dim), np.float32)
        N.check(N.lib().kjc_index_get_rows(self._h, int(row), int(n), _ptr(out)))
        return out

    def search_batch(self, queries, k: int, mode: Optional[int] = None):
        """Top-k of every query: (ids u64 [Q,k], scores f32 [Q,k], counts i32 [Q]); order (score desc, id asc)."""
        q = _c(queries, np.float32)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError("queries must be [nq, dim]")
        nq = q.shape[0]
        ids = np.empty((nq, k), np.uint64)
        sc = np.empty((nq, k), np.float32)
        cnt = np.empty((nq,), np.int32)
        N.check(N.lib().kjc_index_search(self._h, _ptr(q), nq, int(k), self.mode if mode is None else mode, _ptr(ids), _ptr(sc), _ptr(cnt)))
        return ids, sc, cnt

    def search_vectors(self, query, limit: int) -> List[Tuple[int, float]]:
        """Segment::search_vectors: [] on a dimension mismatch or zero-norm query; local doc ids."""
        q = np.asarray(query, np.float32)
        if q.ndim != 1 or q.shape[0] != self.dim or limit <= 0:
            return []
        ids, sc, cnt = self.search_batch(q[None, :], min(int(limit), 256), N.SCAN_SEGMENT)
        return [(int(ids[0, j] - self.id_base), float(sc[0, j])) for j in range(int(cnt[0]))]

    @property
    def last_launch_count(self) -> int:
        return int(N.lib().kjc_index_last_launch_count(self._h))

    @property
    def unverified_count(self) -> int:
        """Queries of async searches whose tensor-core filter result could not be proven exact (see kjarni_cuda.h)."""
        return int(N.lib().kjc_index_unverified_count(self._h))

    def set_filter(self, eps: float = 0.0045, min_queries: int = 1) -> None:
        """Test hook (kjarni_cuda_debug.h): proof margin and smallest batch that takes the tensor-core filter path."""
        N.check(N.lib().kjc_dbg_index_set_filter(self._h, float(eps), int(min_queries)))


class VectorStore(IndexShard):
    """In-memory store semantics (KS/vector.rs): scores use max(|q||r|, 1e-9) as denominator."""

    def __init__(self, dim: int, capacity_rows: int, device: int = 0):
        super().__init__(dim, capacity_rows, 0, device, N.SCAN_VECTORSTORE)

    def search(self, query_embedding, limit: int) -> List[Tuple[int, float]]:
        q = np.asarray(query_embedding, np.float32)
        if len(self) == 0 or q.ndim != 1 or q.shape[0] != self.dim or limit <= 0:
            return []
        ids, sc, cnt = self.search_batch(q[None, :], min(int(limit), 256), N.SCAN_VECTORSTORE)
        return [(int(ids[0, j]), float(sc[0, j])) for j in range(int(cnt[0]))]


class IndexReader:
    """Several shards (segments) searched one after another on their GPUs and merged on the host the way
    IndexReader::search_semantic does: per-shard top-`limit`, concat in shard order, stable sort desc, truncate;
    global id = shard id_base + local id.  (The multi-process NCCL variant lives in kjarni_b200.distributed.)"""

    def __init__(self, shards: Sequence[IndexShard]):
        self.shards = list(shards)

    @classmethod
    def open(cls, root: str, devices: Sequence[int] = (0,)) -> "IndexReader":
        """IndexReader::open (KR/index_reader.rs:161-204) onto GPU shards: one contiguous part per device."""
        return cls([IndexShard.open_dir(root, dev, i, len(devices)) for i, dev in enumerate(devices)])

    def __len__(self) -> int:
        return sum(len(s) for s in self.shards)

    def search_semantic(self, query, limit: int) -> List[Tuple[int, float]]:
        allr: List[Tuple[int, float]] = []
        for sh in self.shards:
            allr.extend((i + sh.id_base, s) for i, s in sh.search_vectors(query, limit))
        order = np.argsort(-np.array([r[1] for r in allr], np.float32), kind="stable") if allr else []
        return [allr[i] for i in order[:limit]]


def index_dir_info(root: str) -> dict:
    """config.json + segment table of an on-disk index (host-only, kjc_index_dir_info / kjc_index_dir_segment_lens)."""
    info = N.KjcIndexDirInfo()
    N.check(N.lib().kjc_index_dir_info(str(root).encode(), C.byref(info)))
    lens = np.zeros((max(info.n_segments, 1),), np.uint64)
    n = N.lib().kjc_index_dir_segment_lens(str(root).encode(), _ptr(lens), int(lens.shape[0]))
    if n < 0:
        N.check(N.KJC_LOAD_FAILED)
    return {"dimension": info.dimension, "n_segments": info.n_segments, "n_skipped": info.n_skipped, "total_rows": int(info.total_rows),
            "max_docs_per_segment": int(info.max_docs_per_segment), "segment_lens": [int(x) for x in lens[:n]]}


def index_part_range(total_rows: int, part: int, parts: int) -> Tuple[int, int]:
    lo, hi = C.c_uint64(), C.c_uint64()
    N.check(N.lib().kjc_index_part_range(int(total_rows), int(part), int(parts), C.byref(lo), C.byref(hi)))
    return int(lo.value), int(hi.value)


def cosine_similarity(a, b) -> float:
    a = _c(a, np.float32)
    b = _c(b, np.float32)
    if a.shape != b.shape:
        return 0.0
    return float(N.lib().kjc_cosine_similarity(_ptr(a), _ptr(b), a.size))
