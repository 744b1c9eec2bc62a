"""The public kjarni-ffi C ABI (include/kjarni_ffi.h) served by the B200 backend, called the way the
reference's Python binding does (ctypes on libkjarni_ffi.so, struct layouts of kjarni-ffi/src/*.rs), checked against the
fp32 oracle fed with the same token ids (the tokenizer itself is pinned separately by tests/test_tokenizer_cpu.py)."""
import ctypes as C
import json
import os
import shutil

import numpy as np
import pytest

from kjarni_b200 import _native as N
from kjarni_b200 import api, synth
from oracle import kjarni_oracle as ko
from kj_testutil import cosine_rows

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOK = os.path.join(HERE, "golden", "tokenizers", "bert_uncased.tokenizer.json")
GPU, CPU = 1, 0


class FloatArray(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_float)), ("len", C.c_size_t)]


class Float2DArray(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_float)), ("rows", C.c_size_t), ("cols", C.c_size_t)]


class StringArray(C.Structure):
    _fields_ = [("strings", C.POINTER(C.c_char_p)), ("len", C.c_size_t)]


class EmbedderConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("cache_dir", C.c_char_p), ("model_name", C.c_char_p), ("model_path", C.c_char_p),
                ("normalize", C.c_int32), ("quiet", C.c_int32)]


class ClassResult(C.Structure):
    _fields_ = [("label", C.c_char_p), ("score", C.c_float)]


class ClassResults(C.Structure):
    _fields_ = [("results", C.POINTER(ClassResult)), ("len", C.c_size_t)]


class ClassifierConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("cache_dir", C.c_char_p), ("model_name", C.c_char_p), ("model_path", C.c_char_p),
                ("labels", C.POINTER(C.c_char_p)), ("num_labels", C.c_size_t), ("multi_label", C.c_int32), ("quiet", C.c_int32)]


class RerankResult(C.Structure):
    _fields_ = [("index", C.c_size_t), ("score", C.c_float)]


class RerankResults(C.Structure):
    _fields_ = [("results", C.POINTER(RerankResult)), ("len", C.c_size_t)]


class RerankerConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("cache_dir", C.c_char_p), ("model_name", C.c_char_p), ("model_path", C.c_char_p), ("quiet", C.c_int32)]


class SearchResult(C.Structure):
    _fields_ = [("score", C.c_float), ("document_id", C.c_size_t), ("text", C.c_char_p), ("metadata_json", C.c_char_p)]


class SearchResults(C.Structure):
    _fields_ = [("results", C.POINTER(SearchResult)), ("len", C.c_size_t)]


class SearchOptions(C.Structure):
    _fields_ = [("mode", C.c_int32), ("top_k", C.c_size_t), ("use_reranker", C.c_int32), ("threshold", C.c_float),
                ("source_pattern", C.c_char_p), ("filter_key", C.c_char_p), ("filter_value", C.c_char_p)]


class SearcherConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("cache_dir", C.c_char_p), ("model_name", C.c_char_p), ("rerank_model", C.c_char_p),
                ("default_mode", C.c_int), ("default_top_k", C.c_size_t), ("quiet", C.c_int32)]


@pytest.fixture(scope="module")
def ffi():
    lib = C.CDLL(os.path.join(os.path.dirname(N.LIB_PATH), "libkjarni_ffi.so"))  # the reference's library name
    lib.kjarni_last_error_message.restype = C.c_char_p
    lib.kjarni_error_name.restype = C.c_char_p
    lib.kjarni_version.restype = C.c_char_p
    lib.kjarni_embedder_config_default.restype = EmbedderConfig
    lib.kjarni_classifier_config_default.restype = ClassifierConfig
    lib.kjarni_reranker_config_default.restype = RerankerConfig
    lib.kjarni_searcher_config_default.restype = SearcherConfig
    lib.kjarni_search_options_default.restype = SearchOptions
    lib.kjarni_embedder_dim.restype = C.c_size_t
    lib.kjarni_classifier_num_labels.restype = C.c_size_t
    lib.kjarni_searcher_default_top_k.restype = C.c_size_t
    lib.kjarni_searcher_has_reranker.restype = C.c_bool
    for fn in ("kjarni_embedder_free", "kjarni_classifier_free", "kjarni_reranker_free", "kjarni_searcher_free"):
        getattr(lib, fn).argtypes = [C.c_void_p]
    lib.kjarni_embedder_encode.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(FloatArray)]
    lib.kjarni_embedder_encode_batch.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_size_t, C.POINTER(Float2DArray)]
    lib.kjarni_embedder_similarity.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_float)]
    lib.kjarni_embedder_dim.argtypes = [C.c_void_p]
    lib.kjarni_classifier_classify.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(ClassResults)]
    lib.kjarni_classifier_labels.argtypes = [C.c_void_p, C.POINTER(StringArray)]
    lib.kjarni_classifier_num_labels.argtypes = [C.c_void_p]
    lib.kjarni_reranker_score.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_float)]
    lib.kjarni_reranker_rerank.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_size_t, C.POINTER(RerankResults)]
    lib.kjarni_reranker_rerank_top_k.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_size_t, C.c_size_t, C.POINTER(RerankResults)]
    lib.kjarni_searcher_search.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(SearchResults)]
    lib.kjarni_searcher_search_with_options.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(SearchOptions), C.POINTER(SearchResults)]
    for fn in ("kjarni_searcher_has_reranker", "kjarni_searcher_default_mode", "kjarni_searcher_default_top_k"):
        getattr(lib, fn).argtypes = [C.c_void_p]
    assert lib.kjarni_init() == 0
    return lib


@pytest.fixture(scope="module")
def models(tmp_path_factory):
    """Model directories in a cache laid out like ~/.cache/kjarni (repo id with '/' -> '_'), each with a WordPiece tokenizer.json."""
    cache = tmp_path_factory.mktemp("kjarni-cache")
    out = {"cache": str(cache)}
    for arch, sub in (("tiny-bert", "sentence-transformers_all-MiniLM-L6-v2"), ("tiny-distilbert", "distilbert_distilbert-base-uncased-finetuned-sst-2-english"),
                      ("tiny-reranker", "cross-encoder_ms-marco-MiniLM-L-6-v2")):
        d = synth.write_model_dir(str(cache / sub), arch)
        shutil.copy(TOK, os.path.join(d, "tokenizer.json"))
        out[arch] = d
    return out


TEXTS = ["Hello world", "The quick brown fox jumps over the lazy dog.", "unbelievable tokenization of tokenizers!", "Café naïve résumé", "search the vector index",
         "a", "ranking documents by cosine similarity score"]


def strs(xs):
    return (C.c_char_p * len(xs))(*[x.encode() for x in xs])


def test_embedder_matches_oracle(ffi, models):
    cfg = ffi.kjarni_embedder_config_default()
    assert (cfg.device, cfg.normalize, cfg.quiet) == (CPU, 1, 0) and cfg.model_name is None
    h = C.c_void_p()
    assert ffi.kjarni_embedder_new(C.byref(cfg), C.byref(h)) == 7  # KJARNI_DEVICE_CPU: no CPU path here, loudly
    assert b"KJARNI_DEVICE_GPU" in ffi.kjarni_last_error_message()
    cfg.device = GPU
    cfg.cache_dir = models["cache"].encode()
    assert ffi.kjarni_embedder_new(C.byref(cfg), C.byref(h)) == 0, ffi.kjarni_last_error_message()  # default name "minilm-l6-v2" resolved in the cache
    assert ffi.kjarni_embedder_dim(h) == 64
    out = Float2DArray()
    assert ffi.kjarni_embedder_encode_batch(h, strs(TEXTS), len(TEXTS), C.byref(out)) == 0
    got = np.ctypeslib.as_array(out.data, shape=(out.rows, out.cols)).copy()
    ffi.kjarni_float_2d_array_free(C.byref(out))
    assert got.shape == (len(TEXTS), 64)
    tok = api.Tokenizer(TOK, 64)
    ids, mask, _ = tok.encode_batch(TEXTS)
    m = ko.load_model_dir(models["tiny-bert"])
    want = ko.embed(m, ids, mask)
    assert cosine_rows(got, want).min() >= 0.9995 and np.abs(got - want).max() <= 2e-2
    one = FloatArray()
    assert ffi.kjarni_embedder_encode(h, TEXTS[1].encode(), C.byref(one)) == 0 and one.len == 64
    v = np.ctypeslib.as_array(one.data, shape=(64,)).copy()
    ffi.kjarni_float_array_free(C.byref(one))
    i1, m1, _ = tok.encode_batch([TEXTS[1]])
    assert cosine_rows(v[None], ko.embed(m, i1, m1)).min() >= 0.9995
    sim = C.c_float()
    assert ffi.kjarni_embedder_similarity(h, TEXTS[0].encode(), TEXTS[4].encode(), C.byref(sim)) == 0
    i2, m2, _ = tok.encode_batch([TEXTS[0], TEXTS[4]])
    w2 = ko.embed(m, i2, m2)
    assert abs(sim.value - float(ko.cosine_similarity(w2[0], w2[1]))) < 5e-3
    # empty batch, null pointers, invalid UTF-8
    assert ffi.kjarni_embedder_encode_batch(h, strs(TEXTS), 0, C.byref(out)) == 0 and out.rows == 0 and not out.data
    assert ffi.kjarni_embedder_encode(h, None, C.byref(one)) == 1
    assert ffi.kjarni_embedder_encode(h, b"\xff\xfe", C.byref(one)) == 2
    ffi.kjarni_embedder_free(h)
    # un-normalised embeddings, model_path instead of the registry name
    cfg.normalize = 0
    cfg.model_path = models["tiny-bert"].encode()
    assert ffi.kjarni_embedder_new(C.byref(cfg), C.byref(h)) == 0
    assert ffi.kjarni_embedder_encode_batch(h, strs(TEXTS), len(TEXTS), C.byref(out)) == 0
    raw = np.ctypeslib.as_array(out.data, shape=(out.rows, out.cols)).copy()
    ffi.kjarni_float_2d_array_free(C.byref(out))
    assert cosine_rows(raw, ko.embed(m, ids, mask, normalize=False)).min() >= 0.9995 and np.abs(np.linalg.norm(raw, axis=1) - 1).max() > 1e-2
    ffi.kjarni_embedder_free(h)
    # unknown registry name / not downloaded
    cfg.model_path = None
    cfg.model_name = b"no-such-model"
    assert ffi.kjarni_embedder_new(C.byref(cfg), C.byref(h)) == 3 and not h.value
    cfg.model_name = b"mpnet-base-v2"
    assert ffi.kjarni_embedder_new(C.byref(cfg), C.byref(h)) == 3 and b"never downloads" in ffi.kjarni_last_error_message()
    assert ffi.kjarni_embedder_new(C.byref(cfg), None) == 1


def test_classifier_matches_oracle(ffi, models):
    cfg = ffi.kjarni_classifier_config_default()
    cfg.device = GPU
    cfg.cache_dir = models["cache"].encode()
    h = C.c_void_p()
    assert ffi.kjarni_classifier_new(C.byref(cfg), C.byref(h)) == 0, ffi.kjarni_last_error_message()  # default preset "sentiment"
    assert ffi.kjarni_classifier_num_labels(h) == 2
    sa = StringArray()
    assert ffi.kjarni_classifier_labels(h, C.byref(sa)) == 0 and [sa.strings[i] for i in range(sa.len)] == [b"NEGATIVE", b"POSITIVE"]
    ffi.kjarni_string_array_free(C.byref(sa))
    tok = api.Tokenizer(TOK, 64)
    m = ko.load_model_dir(models["tiny-distilbert"])
    for text in TEXTS[:4]:
        res = ClassResults()
        assert ffi.kjarni_classifier_classify(h, text.encode(), C.byref(res)) == 0 and res.len == 2
        got = {res.results[i].label.decode(): res.results[i].score for i in range(2)}
        assert res.results[0].score >= res.results[1].score  # sorted descending
        ffi.kjarni_class_results_free(C.byref(res))
        ids, mask, _ = tok.encode_batch([text])
        p = ko.classify_probs(ko.predict_logits(m, ids, mask))[0]
        assert abs(got["NEGATIVE"] - p[0]) < 2e-2 and abs(got["POSITIVE"] - p[1]) < 2e-2 and abs(sum(got.values()) - 1) < 1e-5
    ffi.kjarni_classifier_free(h)
    # custom labels + multi-label (sigmoid per logit)
    labels = strs(["bad", "good"])
    cfg.labels = C.cast(labels, C.POINTER(C.c_char_p))
    cfg.num_labels = 2
    cfg.multi_label = 1
    assert ffi.kjarni_classifier_new(C.byref(cfg), C.byref(h)) == 0
    res = ClassResults()
    assert ffi.kjarni_classifier_classify(h, TEXTS[1].encode(), C.byref(res)) == 0
    got = {res.results[i].label.decode(): res.results[i].score for i in range(2)}
    ffi.kjarni_class_results_free(C.byref(res))
    ids, mask, _ = tok.encode_batch([TEXTS[1]])
    lg = ko.predict_logits(m, ids, mask)[0]
    assert abs(got["bad"] - 1 / (1 + np.exp(-lg[0]))) < 2e-2 and abs(got["good"] - 1 / (1 + np.exp(-lg[1]))) < 2e-2
    ffi.kjarni_classifier_free(h)
    cfg.num_labels = 1  # label count mismatch
    assert ffi.kjarni_classifier_new(C.byref(cfg), C.byref(h)) == 7
    cfg.labels, cfg.num_labels, cfg.model_name = None, 0, b"minilm-l6-v2"  # an encoder without a head
    assert ffi.kjarni_classifier_new(C.byref(cfg), C.byref(h)) == 7


def test_roberta_classifier_from_text(ffi, models, tmp_path):
    """A RoBERTa-family classifier preset end to end from TEXT: byte-level BPE tokenizer (tokenizer.hpp, pinned to the `tokenizers`
    crate in tests/test_tokenizer_cpu.py) -> position ids offset by 2 -> roberta layout -> classifier.dense / out_proj head."""
    cache = tmp_path / "cache"
    d = synth.write_model_dir(str(cache / "olafuraron_emotion-english-distilroberta-base-safetensors"), "tiny-roberta")
    bpe = os.path.join(HERE, "golden", "tokenizers", "roberta_bpe.tokenizer.json")
    shutil.copy(bpe, os.path.join(d, "tokenizer.json"))
    cfg = ffi.kjarni_classifier_config_default()
    cfg.device, cfg.cache_dir, cfg.model_name = GPU, str(cache).encode(), b"distilroberta-emotion"
    h = C.c_void_p()
    assert ffi.kjarni_classifier_new(C.byref(cfg), C.byref(h)) == 0, ffi.kjarni_last_error_message()
    assert ffi.kjarni_classifier_num_labels(h) == 3
    tok = api.Tokenizer(bpe, 66)
    m = ko.load_model_dir(d)
    for text in TEXTS[:5] + ["It's we'll   spaced\nout, isn't it?"]:
        res = ClassResults()
        assert ffi.kjarni_classifier_classify(h, text.encode(), C.byref(res)) == 0 and res.len == 3
        got = {res.results[i].label.decode(): res.results[i].score for i in range(3)}
        ffi.kjarni_class_results_free(C.byref(res))
        ids, mask, _ = tok.encode_batch([text])
        assert ids[0, 0] == 0 and ids[0, int(mask.sum()) - 1] == 2  # <s> ... </s>
        p = ko.classify_probs(ko.predict_logits(m, ids, mask))[0]
        for i in range(3):
            assert abs(got["LABEL_%d" % i] - p[i]) < 2e-2
    ffi.kjarni_classifier_free(h)


def test_reranker_matches_oracle(ffi, models):
    cfg = ffi.kjarni_reranker_config_default()
    cfg.device = GPU
    cfg.cache_dir = models["cache"].encode()
    h = C.c_void_p()
    assert ffi.kjarni_reranker_new(C.byref(cfg), C.byref(h)) == 0, ffi.kjarni_last_error_message()
    query, docs = "search the index", TEXTS
    tok = api.Tokenizer(TOK, 64)
    ids, mask, types = tok.encode_batch([query] * len(docs), docs)
    assert types.max() == 1  # [CLS] q [SEP] d [SEP] with type ids 0 / 1
    m = ko.load_model_dir(models["tiny-reranker"])
    want = ko.predict_logits(m, ids, mask, types)[:, 0]
    res = RerankResults()
    assert ffi.kjarni_reranker_rerank(h, query.encode(), strs(docs), len(docs), C.byref(res)) == 0 and res.len == len(docs)
    got = np.array([res.results[i].score for i in range(res.len)])
    idx = [res.results[i].index for i in range(res.len)]
    ffi.kjarni_rerank_results_free(C.byref(res))
    assert sorted(idx) == list(range(len(docs))) and (np.diff(got) <= 0).all()
    scale = max(1.0, float(np.abs(want).max()))
    assert np.abs(got - want[idx]).max() <= 5e-2 * scale
    # every pair of documents whose oracle scores differ by more than 4x the measured error is ranked as the oracle ranks it
    err = float(np.abs(got - want[idx]).max())
    rank = np.empty(len(docs), np.int64)
    rank[np.asarray(idx)] = np.arange(len(docs))
    decided = (want[:, None] - want[None, :]) > 4 * err + 1e-6
    assert decided.sum() >= len(docs), (decided.sum(), err)  # tiny-reranker spreads its scores: most pairs are clearly separated
    ii, jj = np.nonzero(decided)
    assert (rank[ii] < rank[jj]).all()
    assert ffi.kjarni_reranker_rerank_top_k(h, query.encode(), strs(docs), len(docs), 3, C.byref(res)) == 0 and res.len == 3
    assert [res.results[i].index for i in range(3)] == idx[:3]
    ffi.kjarni_rerank_results_free(C.byref(res))
    sc = C.c_float()
    assert ffi.kjarni_reranker_score(h, query.encode(), docs[2].encode(), C.byref(sc)) == 0
    i1, m1, t1 = tok.encode_batch([query], [docs[2]])
    assert abs(sc.value - float(ko.predict_logits(m, i1, m1, t1)[0, 0])) <= 5e-2 * scale
    assert ffi.kjarni_reranker_rerank(h, query.encode(), strs(docs), 0, C.byref(res)) == 0 and res.len == 0
    ffi.kjarni_reranker_free(h)


def test_searcher_semantic_search_over_an_index_directory(ffi, models, tmp_path):
    # index the texts with the embedder's own vectors, in two segments, then search through the public ABI
    enc = api.EncoderModel(models["tiny-bert"])
    tok = api.Tokenizer(TOK, 64)
    docs = TEXTS + ["completely different words about gpu kernels", "water and oil and trees"]
    ids, mask, _ = tok.encode_batch(docs)
    emb = enc.encode_batch_from_ids(ids, mask)
    enc.close()
    segs = [emb[:5], emb[5:]]
    meta = [[{"source": "/data/a/doc%d.txt" % i, "lang": "en"} for i in range(5)], [{"source": "/data/b/note%d.md" % i, "lang": "is"} for i in range(len(docs) - 5)]]
    root = synth.write_index_dir(str(tmp_path / "idx"), segs, docs=[docs[:5], docs[5:]], metadata=meta)
    cfg = ffi.kjarni_searcher_config_default()
    assert cfg.default_mode == 2 and cfg.default_top_k == 10  # Hybrid, 10
    cfg.device, cfg.cache_dir, cfg.default_mode, cfg.default_top_k = GPU, models["cache"].encode(), 1, 3
    h = C.c_void_p()
    assert ffi.kjarni_searcher_new(C.byref(cfg), C.byref(h)) == 0, ffi.kjarni_last_error_message()
    assert not ffi.kjarni_searcher_has_reranker(h) and ffi.kjarni_searcher_default_top_k(h) == 3
    res = SearchResults()
    for qi in (1, 4, 7):
        assert ffi.kjarni_searcher_search(h, root.encode(), docs[qi].encode(), C.byref(res)) == 0, ffi.kjarni_last_error_message()
        assert res.len == 3
        top = res.results[0]
        assert top.document_id == qi and top.text.decode() == docs[qi] and abs(top.score - 1.0) < 1e-3  # the document itself
        md = json.loads(top.metadata_json)
        assert md == meta[0 if qi < 5 else 1][qi if qi < 5 else qi - 5]
        scores = [res.results[i].score for i in range(res.len)]
        assert scores == sorted(scores, reverse=True)
        ffi.kjarni_search_results_free(C.byref(res))
    # options: top_k, metadata filter, source pattern, threshold
    o = ffi.kjarni_search_options_default()
    assert (o.mode, o.top_k, o.use_reranker) == (-1, 0, -1)
    o.top_k = 2
    o.filter_key, o.filter_value = b"lang", b"is"
    assert ffi.kjarni_searcher_search_with_options(h, root.encode(), docs[1].encode(), C.byref(o), C.byref(res)) == 0
    assert res.len == 2 and all(res.results[i].document_id >= 5 for i in range(res.len))
    ffi.kjarni_search_results_free(C.byref(res))
    o.filter_key = o.filter_value = None
    o.source_pattern = b"*.md"
    assert ffi.kjarni_searcher_search_with_options(h, root.encode(), docs[1].encode(), C.byref(o), C.byref(res)) == 0
    assert res.len == 2 and all(json.loads(res.results[i].metadata_json)["source"].endswith(".md") for i in range(res.len))
    ffi.kjarni_search_results_free(C.byref(res))
    o.source_pattern = None
    o.threshold = 0.999
    assert ffi.kjarni_searcher_search_with_options(h, root.encode(), docs[2].encode(), C.byref(o), C.byref(res)) == 0 and res.len == 1
    ffi.kjarni_search_results_free(C.byref(res))
    # keyword mode (0): BM25 over the segments' bm25.bin, per-segment top-k -> stable merge (index_reader.rs:230-245)
    def hits(mode, query, top_k):
        oo = ffi.kjarni_search_options_default()
        oo.mode, oo.top_k = mode, top_k
        assert ffi.kjarni_searcher_search_with_options(h, root.encode(), query.encode(), C.byref(oo), C.byref(res)) == 0, ffi.kjarni_last_error_message()
        out = [(res.results[i].document_id, res.results[i].score) for i in range(res.len)]
        ffi.kjarni_search_results_free(C.byref(res))
        return out

    kq = "ranking the vector index"
    want_kw = ko.index_search_keywords([docs[:5], docs[5:]], kq, 4)
    got_kw = hits(0, kq, 4)
    assert [d for d, _ in got_kw] == [d for d, _ in want_kw] and len(got_kw) >= 2
    assert np.allclose([v for _, v in got_kw], [v for _, v in want_kw], rtol=1e-5)
    # hybrid mode (2, the default of the reference's SearcherConfig): 2 x limit keyword and semantic hits fused by reciprocal rank
    # (index_reader.rs:248-289, hybrid.rs:3-31); the semantic half is this library's own GPU scan, the fusion is checked against the oracle
    top_k = 4
    sem = hits(1, kq, 2 * top_k)
    want_h = ko.rrf_hybrid(ko.index_search_keywords([docs[:5], docs[5:]], kq, 2 * top_k), sem, top_k)
    got_h = hits(2, kq, top_k)
    assert [d for d, _ in got_h] == [d for d, _ in want_h]
    assert np.allclose([v for _, v in got_h], [v for _, v in want_h], rtol=1e-6)
    # missing index; dimension mismatch
    assert ffi.kjarni_searcher_search(h, str(tmp_path / "nope").encode(), b"x", C.byref(res)) == 3
    bad = synth.write_index_dir(str(tmp_path / "idx32"), [np.ones((3, 32), np.float32)])
    assert ffi.kjarni_searcher_search(h, bad.encode(), b"x", C.byref(res)) == 7 and b"Dimension mismatch" in ffi.kjarni_last_error_message()
    ffi.kjarni_searcher_free(h)
    # with a reranker: scores become cross-encoder logits, order follows them
    cfg.rerank_model = b"minilm-l6-v2-cross-encoder"
    assert ffi.kjarni_searcher_new(C.byref(cfg), C.byref(h)) == 0, ffi.kjarni_last_error_message()
    assert ffi.kjarni_searcher_has_reranker(h)
    assert ffi.kjarni_searcher_search(h, root.encode(), docs[1].encode(), C.byref(res)) == 0 and res.len == 3
    scores = [res.results[i].score for i in range(res.len)]
    assert scores == sorted(scores, reverse=True)
    ffi.kjarni_search_results_free(C.byref(res))
    ffi.kjarni_searcher_free(h)


def test_length_bucketed_batching_matches_batch_longest(ffi, models, monkeypatch):
    """texts of very different lengths run in per-length groups; the rows equal the BatchLongest forward
    (padding never reaches a valid token) and come back in the caller's order."""
    cfg = ffi.kjarni_embedder_config_default()
    cfg.device, cfg.cache_dir = GPU, models["cache"].encode()
    h = C.c_void_p()
    assert ffi.kjarni_embedder_new(C.byref(cfg), C.byref(h)) == 0
    texts = ["a", "the quick brown fox " * 20, "hello world", "search " * 30, "vector database retrieval of relevant passage", "dog",
             "ranking documents by cosine similarity score " * 3, "water"]
    out = Float2DArray()

    def run():
        assert ffi.kjarni_embedder_encode_batch(h, strs(texts), len(texts), C.byref(out)) == 0
        v = np.ctypeslib.as_array(out.data, shape=(out.rows, out.cols)).copy()
        ffi.kjarni_float_2d_array_free(C.byref(out))
        return v

    bucketed = run()
    monkeypatch.setenv("KJC_NO_LENGTH_BUCKETS", "1")
    padded = run()
    monkeypatch.delenv("KJC_NO_LENGTH_BUCKETS")
    assert cosine_rows(bucketed, padded).min() >= 0.99999 and np.abs(bucketed - padded).max() < 2e-3
    tok = api.Tokenizer(TOK, 64)
    ids, mask, _ = tok.encode_batch(texts)
    assert mask.sum(1).min() == 3 and mask.sum(1).max() == 64  # [CLS] a [SEP] ... truncated to the 64-position table
    want = ko.embed(ko.load_model_dir(models["tiny-bert"]), ids, mask)
    assert cosine_rows(bucketed, want).min() >= 0.9995 and np.abs(bucketed - want).max() <= 2e-2
    ffi.kjarni_embedder_free(h)


# ------------------------------------------------------------------ indexer
class IndexerConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("cache_dir", C.c_char_p), ("model_name", C.c_char_p), ("chunk_size", C.c_size_t),
                ("chunk_overlap", C.c_size_t), ("batch_size", C.c_size_t), ("extensions", C.c_char_p), ("exclude_patterns", C.c_char_p),
                ("recursive", C.c_int32), ("include_hidden", C.c_int32), ("max_file_size", C.c_size_t), ("quiet", C.c_int32)]


class IndexStats(C.Structure):
    _fields_ = [("documents_indexed", C.c_size_t), ("chunks_created", C.c_size_t), ("dimension", C.c_size_t), ("size_bytes", C.c_uint64),
                ("files_processed", C.c_size_t), ("files_skipped", C.c_size_t), ("elapsed_ms", C.c_uint64)]


class Progress(C.Structure):
    _fields_ = [("stage", C.c_int), ("current", C.c_size_t), ("total", C.c_size_t), ("message", C.c_char_p)]


PROGRESS_FN = C.CFUNCTYPE(None, Progress, C.c_void_p)


def test_indexer_builds_the_index_the_reference_would(ffi, models, tmp_path):
    """kjarni_indexer_create / _add (Indexer::create / add, kjarni/src/indexer/model.rs:169-300,466-570): discover -> split -> embed on the
    GPU -> segments.  Checked against the oracle's TextSplitter and BM25 restatements (pinned to the reference's unit vectors in
    tests/test_search_cpu.py), against the embedder's own rows (vectors.bin is written from the pinned output buffer) and the fp32
    oracle embedding, and by reading the finished directory back through IndexReader's path (searcher, keywords, index_info)."""
    import struct

    ffi.kjarni_indexer_config_default.restype = IndexerConfig
    ffi.kjarni_indexer_free.argtypes = [C.c_void_p]
    ffi.kjarni_indexer_create.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_size_t, C.c_int32, C.POINTER(IndexStats)]
    ffi.kjarni_indexer_add.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_size_t, C.POINTER(C.c_size_t)]
    ffi.kjarni_indexer_create_with_callback.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_size_t, C.c_int32, PROGRESS_FN, C.c_void_p,
                                                        C.c_void_p, C.POINTER(IndexStats)]
    ffi.kjarni_indexer_dimension.restype = C.c_size_t
    ffi.kjarni_indexer_dimension.argtypes = [C.c_void_p]
    ffi.kjarni_indexer_chunk_size.restype = C.c_size_t
    ffi.kjarni_indexer_chunk_size.argtypes = [C.c_void_p]
    ffi.kjarni_search_keywords.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(SearchResults)]
    ffi.kjarni_cancel_token_new.restype = C.c_void_p
    ffi.kjarni_cancel_token_cancel.argtypes = [C.c_void_p]
    ffi.kjarni_cancel_token_free.argtypes = [C.c_void_p]

    corpus = tmp_path / "corpus"
    (corpus / "sub").mkdir(parents=True)
    para = ["Paragraph %d talks about %s and the way %s changes retrieval quality." % (i, w, w)
            for i, w in enumerate(["gpu kernels", "vector search", "tokenizers", "layer norm", "attention heads", "bm25 ranking", "cosine scores"])]
    files = {"a.txt": "\n\n".join(para[:4]), "sub/b.md": "A short note about trees and rivers.", "c.txt": "\n\n".join(para[4:]) + "\n\n" + "x" * 30,
             ".hidden.txt": "hidden files are skipped by default", "d.bin": "unsupported extension"}
    for name, text in files.items():
        (corpus / name).write_text(text)
    dflt = ffi.kjarni_indexer_config_default()
    assert (dflt.device, dflt.chunk_size, dflt.chunk_overlap, dflt.batch_size, dflt.recursive, dflt.include_hidden) == (CPU, 512, 50, 32, 1, 0)
    cfg = ffi.kjarni_indexer_config_default()
    cfg.device, cfg.cache_dir, cfg.chunk_size, cfg.chunk_overlap, cfg.batch_size, cfg.quiet = GPU, models["cache"].encode(), 120, 20, 3, 1
    h = C.c_void_p()
    assert ffi.kjarni_indexer_new(C.byref(cfg), C.byref(h)) == 0, ffi.kjarni_last_error_message()
    assert ffi.kjarni_indexer_dimension(h) == 64 and ffi.kjarni_indexer_chunk_size(h) == 120
    root = str(tmp_path / "built")
    inputs = strs([str(corpus)])
    stats = IndexStats()
    assert ffi.kjarni_indexer_create(h, root.encode(), inputs, 1, 0, C.byref(stats)) == 0, ffi.kjarni_last_error_message()
    # expected chunks: supported, non-hidden files in path order, split by the oracle's TextSplitter
    order = ["a.txt", "c.txt", "sub/b.md"]
    chunks, src = [], []
    for name in order:
        cs = ko.split_text(files[name], 120, 20)
        chunks += cs
        src += [(str(corpus / name), i, len(cs)) for i in range(len(cs))]
    assert len(chunks) > 6
    assert (stats.documents_indexed, stats.chunks_created, stats.dimension, stats.files_processed, stats.files_skipped) == (len(chunks), len(chunks), 64, 3, 0)
    # the directory parses as an index; vectors.bin holds the embedder's rows in chunk order
    info = N.KjcIndexDirInfo()
    N.check(N.lib().kjc_index_dir_info(root.encode(), C.byref(info)))
    assert (info.total_rows, info.dimension, info.n_segments) == (len(chunks), 64, 1)
    seg = os.path.join(root, "segments", "seg_000000")
    vec = np.fromfile(os.path.join(seg, "vectors.bin"), np.float32).reshape(-1, 64)
    assert vec.shape[0] == len(chunks)
    tok = api.Tokenizer(TOK, 64)
    ids, mask, _ = tok.encode_batch(chunks)
    want = ko.embed(ko.load_model_dir(models["tiny-bert"]), ids, mask)
    cos = (vec * want).sum(1) / (np.linalg.norm(vec, axis=1) * np.linalg.norm(want, axis=1))
    assert cos.min() >= 0.9995 and np.abs(vec - want).max() <= 2e-2
    enc = api.EncoderModel(models["tiny-bert"])
    assert np.abs(vec - enc.encode_batch_from_ids(ids, mask)).max() <= 2e-3  # same kernels; grouping by length only changes the padded length
    enc.close()
    # docs / metadata / bm25.bin as SegmentBuilder::flush leaves them
    offs = struct.unpack("<Q%dQ" % len(chunks), open(os.path.join(seg, "docs.idx"), "rb").read())
    assert offs[0] == len(chunks)
    blob = open(os.path.join(seg, "docs.bin"), "rb").read()
    bounds = list(offs[1:]) + [len(blob)]
    assert [blob[bounds[i]:bounds[i + 1]].decode() for i in range(len(chunks))] == [c + "\n" for c in chunks]  # text + '\n' (segment.rs:105-110)
    metas = [json.loads(line) for line in open(os.path.join(seg, "metadata.jsonl"))]
    assert [(m["source"], int(m["chunk_index"]), int(m["total_chunks"])) for m in metas] == src
    bm = ko.bm25_from_bincode(open(os.path.join(seg, "bm25.bin"), "rb").read())
    ref = ko.Bm25()
    for i, t in enumerate(chunks):
        ref.add_document(i, t)
    assert bm.total_docs == ref.total_docs and list(bm.doc_lengths) == list(ref.doc_lengths) and bm.doc_frequencies == ref.doc_frequencies
    assert bm.search("retrieval quality of vector search", 5) == ref.search("retrieval quality of vector search", 5)
    # read back through the searcher: a chunk's own text finds that chunk first, keyword search agrees with the oracle
    scfg = ffi.kjarni_searcher_config_default()
    scfg.device, scfg.cache_dir, scfg.default_mode, scfg.default_top_k = GPU, models["cache"].encode(), 1, 3
    sh = C.c_void_p()
    assert ffi.kjarni_searcher_new(C.byref(scfg), C.byref(sh)) == 0
    res = SearchResults()
    for ci in (0, len(chunks) // 2, len(chunks) - 1):
        assert ffi.kjarni_searcher_search(sh, root.encode(), chunks[ci].encode(), C.byref(res)) == 0, ffi.kjarni_last_error_message()
        assert res.len == 3 and res.results[0].document_id == ci and res.results[0].text.decode() == chunks[ci]
        ffi.kjarni_search_results_free(C.byref(res))
    assert ffi.kjarni_search_keywords(root.encode(), b"trees rivers", 5, C.byref(res)) == 0
    assert [res.results[i].document_id for i in range(res.len)] == [d for d, _ in ko.index_search_keywords([chunks], "trees rivers", 5)]
    ffi.kjarni_search_results_free(C.byref(res))
    # create over an existing index needs force; add appends a new segment and the searcher sees it (fingerprint check)
    assert ffi.kjarni_indexer_create(h, root.encode(), inputs, 1, 0, C.byref(stats)) != 0
    extra = tmp_path / "extra.txt"
    extra.write_text("Completely new material about submarines and lighthouses.")
    added = C.c_size_t()
    assert ffi.kjarni_indexer_add(h, root.encode(), strs([str(extra)]), 1, C.byref(added)) == 0, ffi.kjarni_last_error_message()
    assert added.value == 1
    N.check(N.lib().kjc_index_dir_info(root.encode(), C.byref(info)))
    assert (info.total_rows, info.n_segments) == (len(chunks) + 1, 2)
    assert ffi.kjarni_searcher_search(sh, root.encode(), b"Completely new material about submarines and lighthouses.", C.byref(res)) == 0
    assert res.results[0].document_id == len(chunks)
    ffi.kjarni_search_results_free(C.byref(res))
    ffi.kjarni_searcher_free(sh)
    assert ffi.kjarni_indexer_create(h, root.encode(), inputs, 1, 1, C.byref(stats)) == 0 and stats.documents_indexed == len(chunks)
    # progress callback and cancellation
    stages = []
    cb = PROGRESS_FN(lambda p, _u: stages.append(p.stage))
    root2 = str(tmp_path / "built2")
    assert ffi.kjarni_indexer_create_with_callback(h, root2.encode(), inputs, 1, 0, cb, None, None, C.byref(stats)) == 0
    assert {0, 1, 2, 4} <= set(stages)  # scanning, loading, embedding, committing
    token = ffi.kjarni_cancel_token_new()
    ffi.kjarni_cancel_token_cancel(token)
    assert ffi.kjarni_indexer_create_with_callback(h, str(tmp_path / "built3").encode(), inputs, 1, 0, cb, None, token, C.byref(stats)) == 8  # Cancelled
    ffi.kjarni_cancel_token_free(token)
    ffi.kjarni_indexer_free(h)
