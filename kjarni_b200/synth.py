"""Random-init model directories and synthetic inputs (no network, no checkpoints).

Writes exactly what Kjarni's own loader accepts (`EncoderLoader::load_from_pretrained`,
kjarni-transformers/src/pipeline/encoder/loader.rs:82-141): `model.safetensors`
(F32, tensor names per the three layouts in SURVEY.md Appendix B), `config.json`
and a minimal WordLevel `tokenizer.json` -- the same recipe as the reference's
fixture test kjarni-models/src/models/sentence_encoder/tests.rs:16-176.
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, Optional, Tuple

import numpy as np

ARCHS = {
    # name: (family, hidden, layers, heads, intermediate, vocab, max_pos, type_vocab, num_labels)
    "minilm-l6": ("bert", 384, 6, 12, 1536, 30522, 512, 2, 0),
    "minilm-l6-cross-encoder": ("bert_prefixed", 384, 6, 12, 1536, 30522, 512, 2, 1),
    "distilbert-sst2": ("distilbert", 768, 6, 12, 3072, 30522, 512, 0, 2),
    "bert-base": ("bert", 768, 12, 12, 3072, 30522, 512, 2, 0),
    # RoBERTa-family classifier (distilroberta-emotion shape: 7 labels, classifier.dense + out_proj head)
    # and MPNet sentence encoder (all-mpnet-base-v2 shape); both take positions from row 2 of the table
    "distilroberta-emotion": ("roberta", 768, 6, 12, 3072, 50265, 514, 1, 7),
    "mpnet-base": ("mpnet", 768, 12, 12, 3072, 30527, 514, 0, 0),
    "tiny-roberta": ("roberta", 64, 2, 4, 256, 1000, 66, 1, 3),
    "tiny-mpnet": ("mpnet", 64, 2, 4, 256, 1000, 66, 0, 0),
    # tiny shapes for CPU-speed tests
    "tiny-bert": ("bert", 64, 2, 4, 256, 1000, 64, 2, 0),      # head_dim 16: the mma.sync attention kernel (the tcgen05 kernels need 32 / 64)
    "tiny-bert32": ("bert", 128, 2, 4, 512, 1000, 64, 2, 0),   # head_dim 32: the tcgen05 attention kernel at a tiny size
    "tiny-cross-encoder": ("bert_prefixed", 64, 2, 4, 256, 1000, 64, 2, 1),
    "tiny-distilbert": ("distilbert", 128, 2, 2, 512, 1000, 64, 0, 2),
    # same shape as tiny-cross-encoder with wider-initialised layer weights (ARCH_STD): the token content reaches the CLS row, so
    # candidate scores are spread far beyond the bf16 error and rank-order checks have clearly separated pairs to decide
    "tiny-reranker": ("bert_prefixed", 64, 2, 4, 256, 1000, 64, 2, 1),
}
ARCH_STD = {"tiny-reranker": 0.15}  # std of the projection weights (default 0.02)


def write_safetensors(path: str, tensors: Dict[str, np.ndarray]) -> None:
    """Standard safetensors container: u64 LE header size, JSON header, raw LE data."""
    header = {}
    off = 0
    order = sorted(tensors)
    for name in order:
        a = np.ascontiguousarray(tensors[name], dtype="<f4")
        header[name] = {"dtype": "F32", "shape": list(a.shape), "data_offsets": [off, off + a.nbytes]}
        off += a.nbytes
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for name in order:
            f.write(np.ascontiguousarray(tensors[name], dtype="<f4").tobytes())


def make_weights(arch: str, seed: int = 1234) -> Tuple[Dict[str, np.ndarray], dict]:
    family, H, L, heads, I, vocab, max_pos, type_vocab, num_labels = ARCHS[arch]
    rng = np.random.default_rng(seed)
    t: Dict[str, np.ndarray] = {}

    wstd = ARCH_STD.get(arch, 0.02)

    def mat(*shape, std=None):
        std = wstd if std is None else std
        return (rng.standard_normal(shape) * std).astype(np.float32)

    def bias(n):
        return (rng.standard_normal(n) * 0.05).astype(np.float32)

    def gamma(n):
        return (1.0 + rng.standard_normal(n) * 0.05).astype(np.float32)

    if family == "distilbert":
        ep, lp = "distilbert.embeddings.", "distilbert.transformer.layer.{}."
        n = dict(q="attention.q_lin", k="attention.k_lin", v="attention.v_lin", o="attention.out_lin",
                 ln1="sa_layer_norm", f1="ffn.lin1", f2="ffn.lin2", ln2="output_layer_norm")
    elif family == "mpnet":
        ep, lp = "embeddings.", "encoder.layer.{}."
        n = dict(q="attention.attn.q", k="attention.attn.k", v="attention.attn.v", o="attention.attn.o",
                 ln1="attention.LayerNorm", f1="intermediate.dense", f2="output.dense", ln2="output.LayerNorm")
    else:
        pre = {"bert_prefixed": "bert.", "roberta": "roberta."}.get(family, "")
        ep, lp = pre + "embeddings.", pre + "encoder.layer.{}."
        n = dict(q="attention.self.query", k="attention.self.key", v="attention.self.value",
                 o="attention.output.dense", ln1="attention.output.LayerNorm",
                 f1="intermediate.dense", f2="output.dense", ln2="output.LayerNorm")
    t[ep + "word_embeddings.weight"] = mat(vocab, H)
    t[ep + "position_embeddings.weight"] = mat(max_pos, H)
    if type_vocab:
        t[ep + "token_type_embeddings.weight"] = mat(type_vocab, H)
    t[ep + "LayerNorm.weight"] = gamma(H)
    t[ep + "LayerNorm.bias"] = bias(H)
    for i in range(L):
        p = lp.format(i)
        for key, (o, k) in dict(q=(H, H), k=(H, H), v=(H, H), o=(H, H), f1=(I, H), f2=(H, I)).items():
            t[p + n[key] + ".weight"] = mat(o, k)
            t[p + n[key] + ".bias"] = bias(o)
        for key in ("ln1", "ln2"):
            t[p + n[key] + ".weight"] = gamma(H)
            t[p + n[key] + ".bias"] = bias(H)
    if family == "distilbert":
        t["pre_classifier.weight"] = mat(H, H, std=0.05)
        t["pre_classifier.bias"] = bias(H)
        t["classifier.weight"] = mat(num_labels, H, std=0.5)
        t["classifier.bias"] = bias(num_labels)
        cfg = dict(model_type="distilbert", activation="gelu", dim=H, hidden_dim=I, n_layers=L, n_heads=heads,
                   max_position_embeddings=max_pos, vocab_size=vocab,
                   id2label={"0": "NEGATIVE", "1": "POSITIVE"}, label2id={"NEGATIVE": 0, "POSITIVE": 1})
    elif family == "roberta":
        t["classifier.dense.weight"] = mat(H, H, std=0.05)
        t["classifier.dense.bias"] = bias(H)
        t["classifier.out_proj.weight"] = mat(num_labels, H, std=0.5)
        t["classifier.out_proj.bias"] = bias(num_labels)
        cfg = dict(model_type="roberta", hidden_size=H, num_hidden_layers=L, num_attention_heads=heads, intermediate_size=I,
                   vocab_size=vocab, layer_norm_eps=1e-5, hidden_act="gelu", type_vocab_size=type_vocab,
                   max_position_embeddings=max_pos, position_embedding_type="absolute", pad_token_id=1,
                   id2label={str(i): f"LABEL_{i}" for i in range(num_labels)})
    elif family == "mpnet":
        cfg = dict(model_type="mpnet", hidden_size=H, num_hidden_layers=L, num_attention_heads=heads, intermediate_size=I,
                   vocab_size=vocab, layer_norm_eps=1e-5, hidden_act="gelu", max_position_embeddings=max_pos)
    else:
        cfg = dict(model_type="bert", hidden_size=H, num_hidden_layers=L, num_attention_heads=heads,
                   intermediate_size=I, vocab_size=vocab, layer_norm_eps=1e-12, hidden_act="gelu",
                   type_vocab_size=type_vocab, max_position_embeddings=max_pos)
        if family == "bert_prefixed":
            t["bert.pooler.dense.weight"] = mat(H, H, std=0.05)
            t["bert.pooler.dense.bias"] = bias(H)
            t["classifier.weight"] = mat(num_labels, H, std=0.5)
            t["classifier.bias"] = bias(num_labels)
            cfg["num_labels"] = num_labels
            cfg["id2label"] = {str(i): f"LABEL_{i}" for i in range(num_labels)}
            cfg["_name_or_path"] = "cross-encoder/ms-marco-MiniLM-L-6-v2"
    return t, cfg


def write_model_dir(path: str, arch: str, seed: int = 1234) -> str:
    """Creates `path` with model.safetensors + config.json + tokenizer.json."""
    os.makedirs(path, exist_ok=True)
    t, cfg = make_weights(arch, seed)
    write_safetensors(os.path.join(path, "model.safetensors"), t)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f, indent=1)
    vocab = {"[PAD]": 0, "[UNK]": 100, "[CLS]": 101, "[SEP]": 102}
    tok = {"version": "1.0", "truncation": None, "padding": None, "added_tokens": [], "normalizer": None,
           "pre_tokenizer": {"type": "Whitespace"}, "post_processor": None, "decoder": None,
           "model": {"type": "WordLevel", "vocab": vocab, "unk_token": "[UNK]"}}
    with open(os.path.join(path, "tokenizer.json"), "w") as f:
        json.dump(tok, f)
    return path


def synth_tokens(batch: int, seq: int, vocab: int = 30522, *, regime: str = "T", seed: int = 42,
                 pair: bool = False) -> Tuple[np.ndarray, np.ndarray, Optional[np.ndarray]]:
    """Synthetic token ids / mask / type ids.

    regime 'T' (throughput): every row has length `seq`.  regime 'P' (parity):
    lengths uniform in [seq/4, seq], plus one length-1 row.  [CLS]=101 first,
    [SEP]=102 last, body uniform in [lo, vocab), pad id 0."""
    rng = np.random.default_rng(seed)
    lo = 1000 if vocab > 2000 else 200
    ids = np.zeros((batch, seq), dtype=np.uint32)
    mask = np.zeros((batch, seq), dtype=np.uint32)
    types = np.zeros((batch, seq), dtype=np.uint32) if pair else None
    if regime == "T":
        lens = np.full(batch, seq)
    else:
        lens = rng.integers(max(seq // 4, 2), seq + 1, size=batch)
        lens[batch // 2] = 1
    for b in range(batch):
        n = int(lens[b])
        row = rng.integers(lo, vocab, size=n)
        row[0] = 101
        if n > 1:
            row[n - 1] = 102
        ids[b, :n] = row
        mask[b, :n] = 1
        if pair and n > 3:
            cut = int(rng.integers(2, n - 1))
            ids[b, cut - 1] = 102
            types[b, cut:n] = 1
    return ids, mask, types


def bm25_bincode(texts) -> bytes:
    """bm25.bin of a segment holding `texts`: the Bm25Index that SegmentBuilder::add builds document by document
    (kjarni-search/src/bm25.rs:115-147), serialised as bincode 1.3 does (kjarni-rag/src/segment.rs:163-165): u64 lengths,
    little-endian fixed-width integers, struct fields in declaration order."""
    import struct

    def tokenize(text):  # bm25.rs:192-198
        out, cur = [], []
        for ch in text.lower():
            if ch.isalnum():
                cur.append(ch)
            else:
                if cur:
                    out.append("".join(cur))
                cur = []
        if cur:
            out.append("".join(cur))
        return [t for t in out if len(t.encode()) >= 2]

    df, inv, lens, total_len = {}, {}, [], 0
    for i, t in enumerate(texts):
        toks = tokenize(t)
        lens.append(len(toks))
        total_len += len(toks)
        counts = {}
        for k in toks:
            counts[k] = counts.get(k, 0) + 1
        for k, c in counts.items():
            inv.setdefault(k, []).append((i, c))
            df[k] = df.get(k, 0) + 1
    n = len(texts)
    avg = np.float32(total_len) / np.float32(n) if n else np.float32(0)
    u64 = lambda v: struct.pack("<Q", v)
    st = lambda s: u64(len(s.encode())) + s.encode()
    b = u64(len(df)) + b"".join(st(k) + u64(v) for k, v in df.items())
    b += u64(len(lens)) + b"".join(u64(v) for v in lens)
    b += struct.pack("<f", float(avg)) + u64(n)
    b += u64(len(inv)) + b"".join(st(k) + u64(len(v)) + b"".join(u64(d) + u64(c) for d, c in v) for k, v in inv.items())
    b += struct.pack("<fff", 1.2, 0.75, 0.25) + u64(0) + u64(total_len)
    return b


def write_index_dir(root: str, segments, *, dimension: int = None, max_docs_per_segment: int = 10_000, broken=(), docs=None,
                    metadata=None) -> str:
    """`docs` / `metadata`: optional per-segment lists of texts / {str: str} dicts (defaults: generated text, {}).
    Writes an index directory in the layout IndexWriter::commit leaves (kjarni-rag/src/index_writer.rs:128-170,
    segment.rs:140-193): config.json, index.json, segments/seg_%06d/{segment.json, vectors.bin, docs.bin, docs.idx,
    metadata.jsonl, bm25.bin}.  `segments` = list of float32 [n_i, dim] arrays.  docs.idx is a bincode Vec<u64>
    (u64 length + offsets); bm25.bin is the bincode Bm25Index of the segment's texts (`bm25_bincode`).  `broken` = segment indices written WITHOUT bm25.bin, which IndexReader::open skips."""
    import json
    import struct
    import time

    segs = [np.ascontiguousarray(x, np.float32) for x in segments]
    dim = int(dimension if dimension is not None else segs[0].shape[1])
    os.makedirs(os.path.join(root, "segments"), exist_ok=True)
    with open(os.path.join(root, "config.json"), "w") as f:
        json.dump({"dimension": dim, "max_docs_per_segment": max_docs_per_segment, "max_segment_memory": 100 * 1024 * 1024,
                   "embedding_model": None, "model_name": None, "created_at": None, "version": 1}, f, indent=2)
    total = 0
    for i, rows in enumerate(segs):
        sd = os.path.join(root, "segments", "seg_%06d" % i)
        os.makedirs(sd, exist_ok=True)
        rows.astype("<f4").tofile(os.path.join(sd, "vectors.bin"))
        seg_docs = [("doc %d of segment %d" % (j, i)).encode() for j in range(rows.shape[0])] if docs is None else [t.encode() for t in docs[i]]
        offs, cur = [], 0
        with open(os.path.join(sd, "docs.bin"), "wb") as f:
            for d in seg_docs:
                offs.append(cur)
                f.write(d + b"\n")
                cur += len(d) + 1
        with open(os.path.join(sd, "docs.idx"), "wb") as f:
            f.write(struct.pack("<Q", len(offs)) + b"".join(struct.pack("<Q", o) for o in offs))
        with open(os.path.join(sd, "metadata.jsonl"), "w") as f:
            if metadata is None:
                f.write("{}\n" * rows.shape[0])
            else:
                for m in metadata[i]:
                    f.write(json.dumps(m) + "\n")
        if i not in broken:
            with open(os.path.join(sd, "bm25.bin"), "wb") as f:
                f.write(bm25_bincode([d.decode() for d in seg_docs]))
        with open(os.path.join(sd, "segment.json"), "w") as f:
            json.dump({"id": i, "doc_count": int(rows.shape[0]), "dimension": int(rows.shape[1]), "created_at": int(time.time()),
                       "total_bytes": int(rows.nbytes + cur)}, f, indent=2)
        if i not in broken:
            total += rows.shape[0]
    with open(os.path.join(root, "index.json"), "w") as f:
        json.dump({"total_docs": total, "segment_count": len(segs), "dimension": dim}, f, indent=2)
    return root
