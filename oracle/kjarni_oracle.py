"""CPU fp32 restatement of Kjarni's encoder + cosine-scan hot path (numpy).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`kjarni_b200/`) may
import this module.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` use it, and there only
as the checker (or the timed CPU baseline), never as the thing shipped.

Parity status: PINNED.  The reference is Rust and cannot be compiled in this
image (no cargo/rustc), so this restatement is pinned against the reference's
own golden vectors instead (tests/test_oracle_goldens.py):
  * encoder layer post-/pre-norm goldens  cpu/encoder/encoder_layer.rs:349-448,693-780
  * softmax goldens                        activations.rs:438-505
  * GELU scalars                           activations.rs:311-329
  * FFN GELU golden                        cpu/feedforward/standard_new.rs:154-191
  * LayerNorm KATs                         cpu/normalization/layer_norm.rs:223-330
  * pooling + L2 goldens                   cpu/encoder/traits.rs:780-895
  * classifier-head goldens                cpu/encoder/classifier.rs:595-735
  * cosine / search KATs                   kjarni-search/src/vector.rs:201-309
and cross-checked against HuggingFace BertModel / DistilBert on random-init
weights (tests/golden/make_hf_goldens.py, run in the build container).

All file:line citations are relative to /root/reference/crates/ with
KT = kjarni-transformers/src, KM = kjarni-models/src, KS = kjarni-search/src,
KR = kjarni-rag/src.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32
MASK_VALUE = F32(-1e9)  # KT/utils/masks.rs:4


# --------------------------------------------------------------------------- #
# Scalar / row primitives
# --------------------------------------------------------------------------- #
def _erf32(x: np.ndarray) -> np.ndarray:
    """libm::erff stand-in (KT/activations.rs:5): double erf rounded to f32."""
    flat = np.asarray(x, dtype=np.float64).ravel()
    try:
        from scipy.special import erf as _erf

        out = _erf(flat)
    except Exception:  # pragma: no cover - scipy is in the image
        out = np.array([math.erf(v) for v in flat], dtype=np.float64)
    return out.reshape(np.shape(x)).astype(F32)


def gelu_erf(x: np.ndarray) -> np.ndarray:
    """gelu_scalar, KT/activations.rs:57-59: 0.5*x*(1+erff(x*SQRT_2_INV))."""
    x = np.asarray(x, dtype=F32)
    return (F32(0.5) * x * (F32(1.0) + _erf32(x * F32(0.7071067811865475)))).astype(F32)


def gelu_tanh(x: np.ndarray) -> np.ndarray:
    """gelu_new_scalar, KT/activations.rs:62-66."""
    x = np.asarray(x, dtype=F32)
    inner = F32(0.7978845608) * (x + F32(0.044715) * x * x * x)
    return (F32(0.5) * x * (F32(1.0) + np.tanh(inner).astype(F32))).astype(F32)


def softmax_rows(x: np.ndarray) -> np.ndarray:
    """softmax_inplace over the last axis, KT/activations.rs:223-242.

    max-subtract, exp, sum; divide only if sum > 0 (so an all -inf row stays
    NaN, an all -1e9 row becomes uniform, exactly as the reference)."""
    x = np.asarray(x, dtype=F32)
    if x.shape[-1] == 0:
        return x.copy()
    m = np.max(x, axis=-1, keepdims=True)
    with np.errstate(invalid="ignore", over="ignore"):
        e = np.exp((x - m).astype(F32)).astype(F32)
        s = np.sum(e, axis=-1, keepdims=True, dtype=F32)
        scale = np.where(s > 0, F32(1.0) / s, F32(1.0)).astype(F32)
    return (e * scale).astype(F32)


def layer_norm(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray, eps: float) -> np.ndarray:
    """LayerNorm over the last axis: biased variance, eps inside the sqrt.
    KT/cpu/normalization/layer_norm.rs:37-134 (no-alloc) and :203-215 (alloc)."""
    x = np.asarray(x, dtype=F32)
    h = F32(x.shape[-1])
    mean = (np.sum(x, axis=-1, keepdims=True, dtype=F32) / h).astype(F32)
    d = (x - mean).astype(F32)
    var = (np.sum(d * d, axis=-1, keepdims=True, dtype=F32) / h).astype(F32)
    inv_std = (F32(1.0) / np.sqrt(var + F32(eps))).astype(F32)
    return (d * inv_std * gamma.astype(F32) + beta.astype(F32)).astype(F32)


def linear(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray]) -> np.ndarray:
    """LinearLayer: y = x @ W^T + b with W stored [out, in] row-major.
    KT/linear_layer/linear_layer.rs:160-282 -> KT/cpu/ops/matmul.rs:370-479."""
    y = np.matmul(np.asarray(x, dtype=F32), np.asarray(w, dtype=F32).T)
    if b is not None:
        y = y + np.asarray(b, dtype=F32)
    return y.astype(F32)


# --------------------------------------------------------------------------- #
# Encoder blocks
# --------------------------------------------------------------------------- #
@dataclass
class LayerWeights:
    wq: np.ndarray
    bq: Optional[np.ndarray]
    wk: np.ndarray
    bk: Optional[np.ndarray]
    wv: np.ndarray
    bv: Optional[np.ndarray]
    wo: np.ndarray
    bo: Optional[np.ndarray]
    ln1_g: np.ndarray
    ln1_b: np.ndarray
    w1: np.ndarray
    b1: Optional[np.ndarray]
    w2: np.ndarray
    b2: Optional[np.ndarray]
    ln2_g: np.ndarray
    ln2_b: np.ndarray


def self_attention(
    x: np.ndarray,
    mask: np.ndarray,
    lw: LayerWeights,
    heads: int,
    *,
    noalloc: bool,
    position_bias: Optional[np.ndarray] = None,
) -> np.ndarray:
    """EncoderSelfAttention::forward / forward_noalloc.
    KT/cpu/encoder/encoder_self_attention.rs:61-140 (alloc, mask -> -1e9 via
    utils/masks.rs:7-36) and :143-307 (no-alloc, mask -> -inf at :311-325)."""
    b, s, hdim = x.shape
    d = hdim // heads
    x2 = x.reshape(b * s, hdim)
    q = linear(x2, lw.wq, lw.bq).reshape(b, s, heads, d).transpose(0, 2, 1, 3)
    k = linear(x2, lw.wk, lw.bk).reshape(b, s, heads, d).transpose(0, 2, 1, 3)
    v = linear(x2, lw.wv, lw.bv).reshape(b, s, heads, d).transpose(0, 2, 1, 3)
    scores = np.matmul(q, k.transpose(0, 1, 3, 2)).astype(F32)
    scores = (scores * F32(1.0 / math.sqrt(d))).astype(F32)  # scale_factor, :244-249
    if position_bias is not None:
        scores = (scores + position_bias.astype(F32)).astype(F32)
    fill = F32(-np.inf) if noalloc else MASK_VALUE
    keymask = (np.asarray(mask, dtype=F32) == 0.0)[:, None, None, :]
    scores = np.where(keymask, fill, scores).astype(F32)
    p = softmax_rows(scores)
    ctx = np.matmul(p, v).astype(F32)  # [b, heads, s, d]
    ctx = ctx.transpose(0, 2, 1, 3).reshape(b * s, hdim)
    return linear(ctx, lw.wo, lw.bo).reshape(b, s, hdim)


def feed_forward(x2: np.ndarray, lw: LayerWeights, act: str = "gelu") -> np.ndarray:
    """StdFeedForwardNew, KT/cpu/feedforward/standard_new.rs:28-80."""
    t = linear(x2, lw.w1, lw.b1)
    if act == "gelu":
        t = gelu_erf(t)
    elif act == "gelu_new":
        t = gelu_tanh(t)
    elif act == "relu":
        t = np.maximum(t, F32(0.0))
    else:
        raise ValueError(act)
    return linear(t, lw.w2, lw.b2)


def encoder_layer(
    x: np.ndarray,
    mask: np.ndarray,
    lw: LayerWeights,
    heads: int,
    eps: float,
    *,
    noalloc: bool,
    prenorm: bool = False,
    position_bias: Optional[np.ndarray] = None,
    act: str = "gelu",
) -> np.ndarray:
    """EncoderLayer post-norm (BERT family) and pre-norm.
    KT/cpu/encoder/encoder_layer.rs:113-179 (no-alloc) / :216-232 (alloc);
    pre-norm :196-214."""
    b, s, h = x.shape
    if prenorm:
        n = layer_norm(x, lw.ln1_g, lw.ln1_b, eps)
        x = (x + self_attention(n, mask, lw, heads, noalloc=noalloc, position_bias=position_bias)).astype(F32)
        n = layer_norm(x, lw.ln2_g, lw.ln2_b, eps)
        return (x + feed_forward(n.reshape(b * s, h), lw, act).reshape(b, s, h)).astype(F32)
    a = self_attention(x, mask, lw, heads, noalloc=noalloc, position_bias=position_bias)
    x = layer_norm((x + a).astype(F32), lw.ln1_g, lw.ln1_b, eps)
    f = feed_forward(x.reshape(b * s, h), lw, act).reshape(b, s, h)
    return layer_norm((x + f).astype(F32), lw.ln2_g, lw.ln2_b, eps)


def embeddings_forward(
    ids: np.ndarray,
    word: np.ndarray,
    pos: Optional[np.ndarray],
    typ: Optional[np.ndarray],
    type_ids: Optional[np.ndarray],
    position_offset: int = 0,
) -> np.ndarray:
    """Embeddings::forward, KT/cpu/embeddings/mod.rs:181-225.
    ids >= vocab -> zero row (:227-246); positions clipped to the table
    (:199-214); no type ids -> row 0 added to every token (:216-223)."""
    ids = np.asarray(ids)
    b, s = ids.shape
    vocab, h = word.shape
    out = np.zeros((b, s, h), dtype=F32)
    ok = ids < vocab
    out[ok] = word[ids[ok]].astype(F32)
    if pos is not None:
        end = min(position_offset + s, pos.shape[0])
        n = max(end - position_offset, 0)
        if n > 0:
            out[:, :n, :] += pos[position_offset : position_offset + n].astype(F32)
    if typ is not None and typ.shape[0] > 0:
        if type_ids is not None:
            tt = np.asarray(type_ids)
            if (tt >= typ.shape[0]).any():
                raise ValueError("Token type ID out of range")  # panics in the reference (:312-317)
            out += typ[tt].astype(F32)
        else:
            out += typ[0].astype(F32)
    return out.astype(F32)


def mean_pool(hidden: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """mean_pool, KT/pooling/mod.rs:11-34 (count 0 -> token-0 row)."""
    hidden = np.asarray(hidden, dtype=F32)
    mask = np.asarray(mask, dtype=F32)
    summed = np.sum(hidden * mask[:, :, None], axis=1, dtype=F32)
    count = np.sum(mask, axis=1, dtype=F32)
    safe = np.where(count == 0, F32(1.0), count)[:, None]
    out = (summed / safe).astype(F32)
    zero = count == 0
    out[zero] = hidden[zero, 0, :]
    return out


def cls_pool(hidden: np.ndarray) -> np.ndarray:
    """KT/pooling/mod.rs:36-38."""
    return np.asarray(hidden, dtype=F32)[:, 0, :].copy()


def max_pool(hidden: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """KT/pooling/mod.rs:41-52."""
    hidden = np.asarray(hidden, dtype=F32).copy()
    hidden[np.asarray(mask) == 0] = MASK_VALUE
    return np.maximum(hidden.max(axis=1), MASK_VALUE).astype(F32)


def last_token_pool(hidden: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """KT/pooling/mod.rs:54-68: last index with mask > 0, else index 0."""
    hidden = np.asarray(hidden, dtype=F32)
    out = np.zeros((hidden.shape[0], hidden.shape[2]), dtype=F32)
    for i in range(hidden.shape[0]):
        nz = np.nonzero(np.asarray(mask[i]) > 0.0)[0]
        out[i] = hidden[i, nz[-1] if len(nz) else 0]
    return out


def l2_normalize(e: np.ndarray) -> np.ndarray:
    """l2_normalize_inplace, KT/cpu/encoder/traits.rs:529-536 (iff norm > 0)."""
    e = np.asarray(e, dtype=F32).copy()
    n = np.sqrt(np.sum(e * e, axis=1, dtype=F32)).astype(F32)
    nz = n > 0
    e[nz] = (e[nz] / n[nz, None]).astype(F32)
    return e


def classification_head(
    hidden: np.ndarray,
    kind: str,
    w_pre: Optional[np.ndarray],
    b_pre: Optional[np.ndarray],
    w_cls: np.ndarray,
    b_cls: Optional[np.ndarray],
) -> np.ndarray:
    """CpuSequenceClassificationHead::forward, KT/cpu/encoder/classifier.rs:210-258.
    kind: 'pooler_tanh' (bert.pooler.dense, :170-187), 'pre_relu'
    (pre_classifier, :151-168), 'dense_tanh' (classifier.dense, :135-150),
    'none' (:189-200)."""
    z = np.asarray(hidden, dtype=F32)[:, 0, :]
    if kind in ("pooler_tanh", "dense_tanh"):
        z = np.tanh(linear(z, w_pre, b_pre)).astype(F32)
    elif kind == "pre_relu":
        z = np.maximum(linear(z, w_pre, b_pre), F32(0.0))
    elif kind != "none":
        raise ValueError(kind)
    return linear(z, w_cls, b_cls)


# --------------------------------------------------------------------------- #
# Model directory -> weights (the three Kjarni tensor-name layouts)
# --------------------------------------------------------------------------- #
@dataclass
class EncoderModel:
    arch: str  # 'bert' | 'bert_prefixed' | 'distilbert' | 'roberta' | 'mpnet'
    hidden: int
    layers: int
    heads: int
    eps: float
    word: np.ndarray
    pos: np.ndarray
    typ: Optional[np.ndarray]
    emb_g: np.ndarray
    emb_b: np.ndarray
    layer: List[LayerWeights]
    head_kind: Optional[str] = None
    w_pre: Optional[np.ndarray] = None
    b_pre: Optional[np.ndarray] = None
    w_cls: Optional[np.ndarray] = None
    b_cls: Optional[np.ndarray] = None
    position_offset: int = 0
    act: str = "gelu"  # 'gelu' (erf) | 'gelu_new' (tanh) | 'relu'
    labels: List[str] = field(default_factory=list)


def _read_safetensors(path: str) -> Dict[str, np.ndarray]:
    """Standard safetensors container: u64 LE header length, JSON header,
    raw little-endian data (KT/weights/safetensors_loader.rs:131-176).
    F32/F16/BF16 are accepted and up-cast (KT/linear_layer/builder.rs:108-138)."""
    import struct

    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(n))
        base = 8 + n
        blob = np.memmap(path, dtype=np.uint8, mode="r", offset=base)
    out = {}
    for name, meta in header.items():
        if name == "__metadata__":
            continue
        a, b = meta["data_offsets"]
        raw = np.asarray(blob[a:b])
        dt = meta["dtype"]
        if dt == "F32":
            arr = raw.view("<f4")
        elif dt == "F16":
            arr = raw.view("<f2").astype(F32)
        elif dt == "BF16":
            arr = (raw.view("<u2").astype(np.uint32) << 16).view("<f4")
        else:
            continue
        out[name] = np.array(arr, dtype=F32).reshape(meta["shape"])
    return out


def load_model_dir(model_dir: str) -> EncoderModel:
    """EncoderLoader::load_from_pretrained restated (KT/pipeline/encoder/loader.rs:82-141):
    config.json:model_type selects the layout (KM/models/sentence_encoder/model.rs:41-54);
    names per KM/models/sentence_encoder/configs.rs:271-364,638-687 and
    KM/models/sequence_classifier/configs.rs:99-143."""
    with open(os.path.join(model_dir, "config.json")) as f:
        cfg = json.load(f)
    t = _read_safetensors(os.path.join(model_dir, "model.safetensors"))
    mt = str(cfg.get("model_type", "bert")).lower()
    act, pos_off = "gelu", 0
    if mt in ("roberta", "distilroberta"):
        # RobertaConfig, KM/models/sequence_classifier/configs.rs:147-283 (extra_pos_embeddings 2 :223; act map :217-222)
        arch = "roberta"
        hidden, layers, heads = cfg["hidden_size"], cfg["num_hidden_layers"], cfg["num_attention_heads"]
        eps = float(cfg["layer_norm_eps"])
        act = {"gelu": "gelu", "gelu_new": "gelu_new", "relu": "relu"}.get(cfg["hidden_act"], "gelu")
        pos_off = 2
        ep = "roberta.embeddings."
        lp = "roberta.encoder.layer.{}."
        names = dict(
            q="attention.self.query", k="attention.self.key", v="attention.self.value",
            o="attention.output.dense", ln1="attention.output.LayerNorm",
            f1="intermediate.dense", f2="output.dense", ln2="output.LayerNorm",
        )
    elif mt == "mpnet":
        # MpnetConfig, KM/models/sentence_encoder/configs.rs:370-468: GeluNew hard-coded (:408), offset 2 (:415), no token types
        arch = "mpnet"
        hidden, layers, heads = cfg["hidden_size"], cfg["num_hidden_layers"], cfg["num_attention_heads"]
        eps = float(cfg["layer_norm_eps"])
        act, pos_off = "gelu_new", 2
        ep = "embeddings."
        lp = "encoder.layer.{}."
        names = dict(
            q="attention.attn.q", k="attention.attn.k", v="attention.attn.v", o="attention.attn.o",
            ln1="attention.LayerNorm", f1="intermediate.dense", f2="output.dense", ln2="output.LayerNorm",
        )
    elif mt == "distilbert":
        arch = "distilbert"
        hidden, layers, heads = cfg["dim"], cfg["n_layers"], cfg["n_heads"]
        eps = 1e-12  # hard-coded, configs.rs:620
        ep = "distilbert.embeddings."
        lp = "distilbert.transformer.layer.{}."
        names = dict(
            q="attention.q_lin", k="attention.k_lin", v="attention.v_lin", o="attention.out_lin",
            ln1="sa_layer_norm", f1="ffn.lin1", f2="ffn.lin2", ln2="output_layer_norm",
        )
    else:
        # is_hf_classification(): id2label or num_labels present => `bert.` prefix (configs.rs:92-95)
        prefixed = (cfg.get("id2label") is not None) or (cfg.get("num_labels") is not None)
        arch = "bert_prefixed" if prefixed else "bert"
        hidden, layers, heads = cfg["hidden_size"], cfg["num_hidden_layers"], cfg["num_attention_heads"]
        eps = float(cfg.get("layer_norm_eps", 1e-12))
        act = {"gelu": "gelu", "gelu_new": "gelu_new", "relu": "relu"}[cfg.get("hidden_act", "gelu")]  # configs.rs:194-200
        pre = "bert." if arch == "bert_prefixed" else ""
        ep = pre + "embeddings."
        lp = pre + "encoder.layer.{}."
        names = dict(
            q="attention.self.query", k="attention.self.key", v="attention.self.value",
            o="attention.output.dense", ln1="attention.output.LayerNorm",
            f1="intermediate.dense", f2="output.dense", ln2="output.LayerNorm",
        )

    def opt(name):
        return t.get(name)

    ls = []
    for i in range(layers):
        p = lp.format(i)
        ls.append(
            LayerWeights(
                wq=t[p + names["q"] + ".weight"], bq=opt(p + names["q"] + ".bias"),
                wk=t[p + names["k"] + ".weight"], bk=opt(p + names["k"] + ".bias"),
                wv=t[p + names["v"] + ".weight"], bv=opt(p + names["v"] + ".bias"),
                wo=t[p + names["o"] + ".weight"], bo=opt(p + names["o"] + ".bias"),
                ln1_g=t[p + names["ln1"] + ".weight"], ln1_b=t[p + names["ln1"] + ".bias"],
                w1=t[p + names["f1"] + ".weight"], b1=opt(p + names["f1"] + ".bias"),
                w2=t[p + names["f2"] + ".weight"], b2=opt(p + names["f2"] + ".bias"),
                ln2_g=t[p + names["ln2"] + ".weight"], ln2_b=t[p + names["ln2"] + ".bias"],
            )
        )
    m = EncoderModel(
        arch=arch, hidden=hidden, layers=layers, heads=heads, eps=eps,
        word=t[ep + "word_embeddings.weight"], pos=t[ep + "position_embeddings.weight"],
        typ=None if arch == "mpnet" else opt(ep + "token_type_embeddings.weight"),
        emb_g=t[ep + "LayerNorm.weight"], emb_b=t[ep + "LayerNorm.bias"], layer=ls,
        position_offset=pos_off, act=act,
    )
    # head auto-detection, first match wins (classifier.rs:113-206)
    if "classifier.dense.weight" in t:
        m.head_kind = "dense_tanh"
        m.w_pre, m.b_pre = t["classifier.dense.weight"], opt("classifier.dense.bias")
        m.w_cls, m.b_cls = t["classifier.out_proj.weight"], opt("classifier.out_proj.bias")
    elif "pre_classifier.weight" in t:
        m.head_kind = "pre_relu"
        m.w_pre, m.b_pre = t["pre_classifier.weight"], opt("pre_classifier.bias")
        m.w_cls, m.b_cls = t["classifier.weight"], opt("classifier.bias")
    elif "bert.pooler.dense.weight" in t:
        m.head_kind = "pooler_tanh"
        m.w_pre, m.b_pre = t["bert.pooler.dense.weight"], opt("bert.pooler.dense.bias")
        m.w_cls, m.b_cls = t["classifier.weight"], opt("classifier.bias")
    elif "classifier.weight" in t:
        m.head_kind = "none"
        m.w_cls, m.b_cls = t["classifier.weight"], opt("classifier.bias")
    id2 = cfg.get("id2label")
    if isinstance(id2, dict):
        m.labels = [id2[k] for k in sorted(id2, key=lambda s: int(s))]
    return m


def encoder_forward(
    m: EncoderModel,
    ids: np.ndarray,
    mask: np.ndarray,
    type_ids: Optional[np.ndarray] = None,
    *,
    noalloc: bool,
) -> np.ndarray:
    """embed -> embed LN -> L post-norm layers.
    KT/cpu/encoder/traits.rs:66-139, transformer_encoder.rs:240-368."""
    x = embeddings_forward(ids, m.word, m.pos, m.typ, type_ids, m.position_offset)
    x = layer_norm(x, m.emb_g, m.emb_b, m.eps)
    maskf = np.asarray(mask, dtype=F32)
    for lw in m.layer:
        x = encoder_layer(x, maskf, lw, m.heads, m.eps, noalloc=noalloc, act=m.act)
    return x


def use_noalloc(tokens: int) -> bool:
    """ComputeStrategy::select, KT/cpu/strategy.rs:29-47."""
    return tokens <= 1 or tokens >= 1000


def embed(m: EncoderModel, ids, mask, *, pooling="mean", normalize=True) -> np.ndarray:
    """SentenceEncoder::encode_batch path on token ids: no token-type ids
    (row 0 added), strategy-selected mask value, pool, L2.
    KM/models/sentence_encoder/model.rs:201-218; KT/cpu/encoder/traits.rs:204-225."""
    ids = np.asarray(ids)
    hidden = encoder_forward(m, ids, mask, None, noalloc=use_noalloc(ids.size))
    maskf = np.asarray(mask, dtype=F32)
    if pooling == "mean":
        e = mean_pool(hidden, maskf)
    elif pooling == "cls":
        e = cls_pool(hidden)
    elif pooling == "max":
        e = max_pool(hidden, maskf)
    elif pooling == "last":
        e = last_token_pool(hidden, maskf)
    else:
        raise ValueError(pooling)
    return l2_normalize(e) if normalize else e


def predict_logits(m: EncoderModel, ids, mask, type_ids=None) -> np.ndarray:
    """SequenceClassifier::predict_logits / CrossEncoder::predict_pairs on ids:
    always the alloc path (mask -1e9).  KM/models/sequence_classifier/mod.rs:266-346,
    KM/models/cross_encoder/model.rs:170-240."""
    if m.typ is not None and type_ids is None:
        type_ids = np.zeros_like(np.asarray(ids))
    hidden = encoder_forward(m, ids, mask, type_ids if m.typ is not None else None, noalloc=False)
    return classification_head(hidden, m.head_kind, m.w_pre, m.b_pre, m.w_cls, m.b_cls)


def classify_probs(logits: np.ndarray) -> np.ndarray:
    """softmax_inplace per row, KM/models/sequence_classifier/mod.rs:256-262."""
    return softmax_rows(logits)


def stable_argsort_desc(scores: np.ndarray) -> np.ndarray:
    """Stable sort descending => ties resolve to the lowest index
    (KM/models/sequence_classifier/mod.rs:369-372, KS/vector.rs:162,
    KR/index_reader.rs:223, KM/models/cross_encoder/model.rs:251-252)."""
    return np.argsort(-np.asarray(scores, dtype=F32), kind="stable")


# --------------------------------------------------------------------------- #
# Cosine scan
# --------------------------------------------------------------------------- #
def cosine_similarity(a: np.ndarray, b: np.ndarray) -> float:
    """VectorStore::cosine_similarity, KS/vector.rs:131-148 (sequential fp32)."""
    if len(a) != len(b):
        return 0.0
    dot = na = nb = F32(0.0)
    for x, y in zip(np.asarray(a, dtype=F32), np.asarray(b, dtype=F32)):
        dot = F32(dot + x * y)
        na = F32(na + x * x)
        nb = F32(nb + y * y)
    den = max(F32(np.sqrt(na) * np.sqrt(nb)), F32(1e-9))
    return float(F32(dot / den))


def vector_store_search(rows: np.ndarray, q: np.ndarray, limit: int) -> List[Tuple[int, float]]:
    """VectorStore::search, KS/vector.rs:150-165: score all, stable sort desc, truncate."""
    rows = np.asarray(rows, dtype=F32)
    if rows.size == 0 or rows.shape[1] != len(q):
        return []
    q = np.asarray(q, dtype=F32)
    dot = rows @ q
    den = np.maximum(np.sqrt(np.sum(q * q, dtype=F32)) * np.sqrt(np.sum(rows * rows, axis=1, dtype=F32)), F32(1e-9))
    s = (dot / den).astype(F32)
    order = stable_argsort_desc(s)[:limit]
    return [(int(i), float(s[i])) for i in order]


def segment_scores(rows: np.ndarray, q: np.ndarray) -> Optional[np.ndarray]:
    """cosine_similarity_with_norm for every row, KR/segment.rs:307-325,355-370:
    None if |q| < 1e-9; 0 for rows with |r| < 1e-9."""
    q = np.asarray(q, dtype=F32)
    rows = np.asarray(rows, dtype=F32)
    qn = np.sqrt(np.sum(q * q, dtype=F32))
    if qn < 1e-9:
        return None
    dot = rows @ q
    rn = np.sqrt(np.sum(rows * rows, axis=1, dtype=F32)).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = (dot / (qn * rn)).astype(F32)
    s[rn < 1e-9] = F32(0.0)
    return s


def segment_search(rows: np.ndarray, q: np.ndarray, limit: int) -> List[Tuple[int, float]]:
    """Segment::search_vectors, KR/segment.rs:307-337.  The reference truncates
    with select_nth_unstable before a stable sort, so exact ties straddling the
    cut are unspecified there; the canonical order here is (score desc, id asc)."""
    if rows.shape[1] != len(q):
        return []
    s = segment_scores(rows, q)
    if s is None:
        return []
    order = stable_argsort_desc(s)[:limit]
    return [(int(i), float(s[i])) for i in order]


def index_search_semantic(segments: Sequence[np.ndarray], q: np.ndarray, limit: int) -> List[Tuple[int, float]]:
    """IndexReader::search_semantic + local_to_global, KR/index_reader.rs:207-228,313-319:
    per-segment top-`limit`, concat in segment order, stable sort desc, truncate;
    global id = sum of preceding segment lengths + local id."""
    allr: List[Tuple[int, float]] = []
    off = 0
    for seg in segments:
        for i, sc in segment_search(seg, q, limit):
            allr.append((off + i, sc))
        off += seg.shape[0]
    order = stable_argsort_desc(np.array([r[1] for r in allr], dtype=F32)) if allr else []
    return [allr[i] for i in order[:limit]]


def batched_topk(rows: np.ndarray, queries: np.ndarray, k: int, row_offset: int = 0):
    """Vectorised form of `segment_search` for Q queries over one shard:
    returns (ids int64 [Q,k], scores f32 [Q,k]) ordered (score desc, id asc);
    zero-norm queries yield ids -1 / scores -inf (the reference returns [])."""
    rows = np.asarray(rows, dtype=F32)
    queries = np.asarray(queries, dtype=F32)
    qn = np.sqrt(np.sum(queries * queries, axis=1, dtype=F32)).astype(F32)
    rn = np.sqrt(np.sum(rows * rows, axis=1, dtype=F32)).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = ((queries @ rows.T) / (qn[:, None] * rn[None, :])).astype(F32)
    s[:, rn < 1e-9] = F32(0.0)
    kk = min(k, rows.shape[0])
    ids = np.full((queries.shape[0], k), -1, dtype=np.int64)
    sc = np.full((queries.shape[0], k), -np.inf, dtype=F32)
    for qi in range(queries.shape[0]):
        if qn[qi] < 1e-9:
            continue
        order = stable_argsort_desc(s[qi])[:kk]
        ids[qi, :kk] = order + row_offset
        sc[qi, :kk] = s[qi, order]
    return ids, sc


# --------------------------------------------------------------------------- #
# Counter-based synthetic index generator; shared definition
# with the CUDA generator kernel so shards can be produced on device and
# re-derived row by row on the CPU.
# --------------------------------------------------------------------------- #
def hash32(seed: int, r: np.ndarray, c: np.ndarray) -> np.ndarray:
    """32-bit mix of (seed, row, col): two rounds of a murmur3-style finaliser."""
    with np.errstate(over="ignore"):
        x = (np.asarray(r, dtype=np.uint64) * np.uint64(0x9E3779B1) + np.asarray(c, dtype=np.uint64) * np.uint64(0x85EBCA77)
             + np.uint64(seed) * np.uint64(0xC2B2AE3D)) & np.uint64(0xFFFFFFFF)
        x = x.astype(np.uint32)
        x ^= x >> np.uint32(16)
        x = (x * np.uint32(0x85EBCA6B)).astype(np.uint32)
        x ^= x >> np.uint32(13)
        x = (x * np.uint32(0xC2B2AE35)).astype(np.uint32)
        x ^= x >> np.uint32(16)
    return x


def synth_rows(seed: int, row0: int, n: int, dim: int) -> np.ndarray:
    """x[r,c] = (hash32(seed,r,c) >> 8) * 2^-24 - 0.5 for rows row0..row0+n."""
    r = np.arange(row0, row0 + n, dtype=np.uint64)[:, None]
    c = np.arange(dim, dtype=np.uint64)[None, :]
    h = hash32(seed, r, c)
    return ((h >> np.uint32(8)).astype(F32) * F32(2.0 ** -24) - F32(0.5)).astype(F32)


# --------------------------------------------------------------------------- #
# Keyword (BM25) and hybrid (RRF) search, text splitting -- CPU in the reference too
# --------------------------------------------------------------------------- #
def bm25_tokenize(text: str) -> List[str]:
    """kjarni-search/src/bm25.rs:192-198: to_lowercase, split on non-alphanumeric chars, keep tokens of >= 2 BYTES."""
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
    return [t for t in out if len(t.encode("utf-8")) >= 2]


class Bm25:
    """Bm25Index, kjarni-search/src/bm25.rs:43-190 (k1 1.2, b 0.75; all score arithmetic in f32)."""

    def __init__(self):
        self.doc_frequencies: Dict[str, int] = {}
        self.doc_lengths: List[int] = []
        self.avg_doc_length = F32(0.0)
        self.total_docs = 0
        self.inverted_index: Dict[str, List[Tuple[int, int]]] = {}
        self.k1, self.b, self.epsilon = F32(1.2), F32(0.75), F32(0.25)
        self.total_length = 0

    def add_document(self, doc_id: int, text: str) -> None:  # :115-147
        tokens = bm25_tokenize(text)
        if doc_id >= len(self.doc_lengths):
            self.doc_lengths.extend([0] * (doc_id + 1 - len(self.doc_lengths)))
        self.doc_lengths[doc_id] = len(tokens)
        counts: Dict[str, int] = {}
        for t in tokens:
            counts[t] = counts.get(t, 0) + 1
        for term, c in counts.items():
            self.inverted_index.setdefault(term, []).append((doc_id, c))
            self.doc_frequencies[term] = self.doc_frequencies.get(term, 0) + 1
        self.total_docs = max(self.total_docs, doc_id + 1)
        self.total_length += len(tokens)
        self.avg_doc_length = F32(F32(self.total_length) / F32(self.total_docs))

    def score(self, query_tokens: Sequence[str], doc_id: int) -> np.float32:  # calculate_score, :149-176
        score = F32(0.0)
        doc_length = F32(self.doc_lengths[doc_id])
        length_norm = F32(F32(1.0) - self.b + self.b * F32(doc_length / self.avg_doc_length))
        for term in query_tokens:
            tf = F32(next((f for d, f in self.inverted_index.get(term, []) if d == doc_id), 0))
            if tf == 0.0:
                continue
            df = F32(self.doc_frequencies.get(term, 0))
            if df == 0.0:
                continue
            idf = F32(np.log(F32(F32(F32(F32(self.total_docs) - df) + F32(0.5)) / F32(df + F32(0.5)) + F32(1.0))))
            ntf = F32(F32(tf * F32(self.k1 + F32(1.0))) / F32(tf + F32(self.k1 * length_norm)))
            score = F32(score + F32(idf * ntf))
        return score

    def search(self, query: str, limit: int) -> List[Tuple[int, float]]:  # :85-113; equal scores: ascending doc id (reference: HashMap order)
        if self.total_docs == 0:
            return []
        q = bm25_tokenize(query)
        if not q:
            return []
        res = [(d, float(self.score(q, d))) for d in range(self.total_docs)]
        res = [(d, s) for d, s in res if s > 0.0]
        res.sort(key=lambda t: (-t[1], t[0]))
        return res[:limit]


def bm25_from_bincode(buf: bytes) -> Bm25:
    """bm25.bin as SegmentBuilder::flush writes it (bincode 1.3 default options), kjarni-rag/src/segment.rs:163-165."""
    import struct

    p = 0

    def u64():
        nonlocal p
        v = struct.unpack_from("<Q", buf, p)[0]
        p += 8
        return v

    def f32():
        nonlocal p
        v = struct.unpack_from("<f", buf, p)[0]
        p += 4
        return F32(v)

    def s():
        nonlocal p
        n = u64()
        v = buf[p:p + n].decode("utf-8")
        p += n
        return v

    ix = Bm25()
    for _ in range(u64()):
        k = s()
        ix.doc_frequencies[k] = u64()
    ix.doc_lengths = [u64() for _ in range(u64())]
    ix.avg_doc_length = f32()
    ix.total_docs = u64()
    for _ in range(u64()):
        k = s()
        ix.inverted_index[k] = [(u64(), u64()) for _ in range(u64())]
    ix.k1, ix.b, ix.epsilon = f32(), f32(), f32()
    for _ in range(u64()):
        s()
        p += 8 * u64()
    ix.total_length = u64()
    assert p == len(buf), (p, len(buf))
    return ix


def rrf_hybrid(keyword: Sequence[Tuple[int, float]], semantic: Sequence[Tuple[int, float]], limit: int) -> List[Tuple[int, float]]:
    """hybrid_search, kjarni-search/src/hybrid.rs:3-31: reciprocal-rank fusion with k = 60 (f32); equal scores: ascending id."""
    sc: Dict[int, np.float32] = {}
    for lst in (keyword, semantic):
        for rank, (idx, _s) in enumerate(lst):
            v = F32(F32(1.0) / F32(F32(60.0) + F32(rank + 1)))
            sc[idx] = F32(sc[idx] + v) if idx in sc else v
    out = sorted(((i, float(v)) for i, v in sc.items()), key=lambda t: (-t[1], t[0]))
    return out[:limit]


def index_search_keywords(segment_texts: Sequence[Sequence[str]], query: str, limit: int) -> List[Tuple[int, float]]:
    """IndexReader::search_keywords, kjarni-rag/src/index_reader.rs:230-245: per-segment BM25 top-limit, segment order, stable sort."""
    allr, base = [], 0
    for texts in segment_texts:
        ix = Bm25()
        for i, t in enumerate(texts):
            ix.add_document(i, t)
        allr += [(base + d, s) for d, s in ix.search(query, limit)]
        base += len(texts)
    allr.sort(key=lambda t: -t[1])  # stable
    return allr[:limit]


def split_text(text: str, chunk_size: int = 1000, chunk_overlap: int = 200, separator: str = "\n\n") -> List[str]:
    """TextSplitter::split, kjarni-rag/src/splitter.rs:59-163 (section sizes in bytes, overlap / hard split in chars)."""
    def blen(s):
        return len(s.encode("utf-8"))

    def split_large(t):
        out, start = [], 0
        while start < len(t):
            end = min(start + chunk_size, len(t))
            out.append(t[start:end])
            if end >= len(t):
                break
            step = chunk_size - chunk_overlap if 0 < chunk_overlap < chunk_size else chunk_size
            start = start + step if step > 0 else start + 1
        return out

    if not text:
        return []
    chunks, cur = [], ""
    for section in text.split(separator):
        if not section:
            continue
        if blen(section) > chunk_size:
            if cur:
                chunks.append(cur)
                cur = ""
            chunks += split_large(section)
            continue
        would = blen(section) if not cur else blen(cur) + blen(separator) + blen(section)
        if would > chunk_size and cur:
            chunks.append(cur)
            cur = (cur if len(cur) <= chunk_overlap else cur[len(cur) - chunk_overlap:]) if chunk_overlap > 0 else ""
        if cur:
            cur += separator
        cur += section
    if cur:
        chunks.append(cur)
    return chunks
