"""Host-side halves of the Searcher / Indexer (no GPU): BM25 keyword search, reciprocal-rank fusion, text splitting, the bm25.bin
bincode format and the load-compatibility of the public C ABI.  The oracle restatements are pinned to the reference's own unit-test
vectors (kjarni-search/src/bm25.rs:199-400, hybrid.rs:34-61, kjarni-rag/src/splitter.rs:330-500), then the library's
`kjarni_search_keywords` is compared with the oracle on an index directory."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from kjarni_b200 import _native as N
from kjarni_b200 import synth
from oracle import kjarni_oracle as ko

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FFI_SO = os.path.join(os.path.dirname(N.LIB_PATH), "libkjarni_ffi.so")


class SearchResult(C.Structure):
    _fields_ = [("score", C.c_float), ("document_id", C.c_size_t), ("text", C.c_char_p), ("metadata_json", C.c_char_p)]


class SearchResults(C.Structure):
    _fields_ = [("results", C.POINTER(SearchResult)), ("len", C.c_size_t)]


class IndexInfo(C.Structure):
    _fields_ = [("path", C.c_char_p), ("document_count", C.c_size_t), ("segment_count", C.c_size_t), ("dimension", C.c_size_t),
                ("size_bytes", C.c_uint64), ("embedding_model", C.c_char_p)]


# ---------------------------------------------------------------- oracle pinned to the reference's unit tests
def test_bm25_tokenize_reference_vectors():
    # bm25.rs:217-246
    assert ko.bm25_tokenize("Hello World") == ["hello", "world"]
    assert ko.bm25_tokenize("I am a test") == ["am", "test"]
    assert ko.bm25_tokenize("hello, world! how are you?") == ["hello", "world", "how", "are", "you"]
    assert ko.bm25_tokenize("") == [] and ko.bm25_tokenize("   ") == []


def test_bm25_search_reference_vectors():
    ix = ko.Bm25()
    assert ix.search("test query", 10) == []  # empty index, bm25.rs:249-253
    # test_bm25_index_with_documents, bm25.rs:267-309: hand-built postings
    ix.total_docs, ix.doc_lengths, ix.avg_doc_length = 3, [5, 3, 7], np.float32(5.0)
    ix.doc_frequencies = {"rust": 2, "programming": 2, "language": 1, "python": 1, "fast": 1, "safe": 1}
    ix.inverted_index = {"rust": [(0, 1), (2, 1)], "programming": [(0, 1), (1, 1)], "language": [(0, 1)], "python": [(1, 1)], "fast": [(2, 1)],
                         "safe": [(2, 1)]}
    ids = [d for d, _ in ix.search("rust", 10)]
    assert 0 in ids and 2 in ids and 1 not in ids
    assert ix.search("", 10) == []
    # test_bm25_score_ordering, bm25.rs:311-331: more occurrences rank first
    ix = ko.Bm25()
    ix.total_docs, ix.doc_lengths, ix.avg_doc_length = 2, [10, 10], np.float32(10.0)
    ix.doc_frequencies, ix.inverted_index = {"test": 2}, {"test": [(0, 1), (1, 3)]}
    r = ix.search("test", 10)
    assert [d for d, _ in r] == [1, 0] and r[0][1] > r[1][1]
    # the closed form for one term: idf = ln((N - df + 0.5) / (df + 0.5) + 1), tf' = tf (k1 + 1) / (tf + k1 (1 - b + b len/avg))
    idf = np.log(np.float32((2 - 2 + 0.5) / 2.5 + 1.0))
    assert abs(r[1][1] - idf * (1 * 2.2) / (1 + 1.2)) < 1e-6 and abs(r[0][1] - idf * (3 * 2.2) / (3 + 1.2)) < 1e-6


def test_rrf_reference_vectors():
    # hybrid.rs:39-61
    r = ko.rrf_hybrid([(0, 1.0), (1, 0.5)], [(1, 0.9), (2, 0.4)], 10)
    assert r[0][0] == 1 and abs(r[0][1] - (1 / 62 + 1 / 61)) < 1e-7
    assert ko.rrf_hybrid([], [], 10) == []
    assert len(ko.rrf_hybrid([(0, 1.0), (1, 0.9), (2, 0.8)], [(3, 0.9), (4, 0.8), (5, 0.7)], 2)) == 2


def test_split_text_reference_vectors():
    # splitter.rs tests: empty, short, on separator, exceeding the chunk size, hard split with overlap
    assert ko.split_text("") == []
    assert ko.split_text("This is a short text.") == ["This is a short text."]
    t = "First paragraph.\n\nSecond paragraph.\n\nThird paragraph."
    assert ko.split_text(t, 100, 0) == [t]
    assert len(ko.split_text(t, 30, 0)) >= 3
    big = "x" * 250
    parts = ko.split_text(big, 100, 20)
    assert parts[0] == "x" * 100 and all(len(p) <= 100 for p in parts) and "".join(p[20:] if i else p for i, p in enumerate(parts)) == big


def test_bm25_bincode_round_trip():
    texts = ["Rust is a fast and safe programming language", "Python programming is popular", "Rust is fast", "", "Ünïcode tökens naïve café"]
    ix = ko.bm25_from_bincode(synth.bm25_bincode(texts))
    ref = ko.Bm25()
    for i, t in enumerate(texts):
        ref.add_document(i, t)
    assert ix.total_docs == 5 and ix.doc_lengths == ref.doc_lengths and ix.doc_frequencies == ref.doc_frequencies
    for q in ("rust programming", "fast", "café", "nothing here matches"):
        assert ix.search(q, 10) == ref.search(q, 10)


# ---------------------------------------------------------------- the library (host-only entry points)
@pytest.fixture(scope="module")
def ffi():
    lib = C.CDLL(FFI_SO)
    lib.kjarni_last_error_message.restype = C.c_char_p
    lib.kjarni_search_keywords.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(SearchResults)]
    lib.kjarni_search_results_free.argtypes = [C.POINTER(SearchResults)]
    lib.kjarni_index_info.argtypes = [C.c_char_p, C.POINTER(IndexInfo)]
    lib.kjarni_index_info_free.argtypes = [IndexInfo]
    lib.kjarni_index_delete.argtypes = [C.c_char_p]
    lib.kjarni_cancel_token_new.restype = C.c_void_p
    lib.kjarni_cancel_token_is_cancelled.restype = C.c_bool
    for fn in ("kjarni_cancel_token_cancel", "kjarni_cancel_token_is_cancelled", "kjarni_cancel_token_reset", "kjarni_cancel_token_free"):
        getattr(lib, fn).argtypes = [C.c_void_p]
    return lib


DOCS = [["Rust is a fast and safe systems programming language", "Python programming is popular for data science",
         "The quick brown fox jumps over the lazy dog", "GPU kernels are written in CUDA"],
        ["Rust and CUDA can be combined through an FFI layer", "Cosine similarity ranks documents by angle", "fast fast fast retrieval with BM25",
         "programming programming programming languages: rust, python, go"]]


def test_keyword_search_matches_oracle(ffi, tmp_path):
    rng = np.random.default_rng(0)
    root = synth.write_index_dir(str(tmp_path / "idx"), [rng.standard_normal((len(d), 8)).astype(np.float32) for d in DOCS], docs=DOCS,
                                 metadata=[[{"source": "a.txt"}] * len(DOCS[0]), [{"source": "b.md"}] * len(DOCS[1])])
    for query, k in (("rust programming", 10), ("fast", 3), ("cuda ffi layer", 2), ("zzz unknown", 5), ("", 5)):
        want = ko.index_search_keywords(DOCS, query, k)
        res = SearchResults()
        assert ffi.kjarni_search_keywords(root.encode(), query.encode(), k, C.byref(res)) == 0, ffi.kjarni_last_error_message()
        got = [(res.results[i].document_id, res.results[i].score, res.results[i].text.decode(), json.loads(res.results[i].metadata_json)) for i in range(res.len)]
        ffi.kjarni_search_results_free(C.byref(res))
        assert [g[0] for g in got] == [w[0] for w in want], (query, got, want)
        assert np.allclose([g[1] for g in got], [w[1] for w in want], rtol=1e-6, atol=1e-7)
        flat = DOCS[0] + DOCS[1]
        for g in got:
            assert g[2] == flat[g[0]] and g[3]["source"] == ("a.txt" if g[0] < 4 else "b.md")
    res = SearchResults()
    assert ffi.kjarni_search_keywords(str(tmp_path / "missing").encode(), b"x", 3, C.byref(res)) != 0 and res.len == 0
    assert ffi.kjarni_search_keywords(None, b"x", 3, C.byref(res)) == 1


def test_index_info_and_delete(ffi, tmp_path):
    rng = np.random.default_rng(1)
    root = synth.write_index_dir(str(tmp_path / "idx"), [rng.standard_normal((5, 16)).astype(np.float32), rng.standard_normal((3, 16)).astype(np.float32)])
    info = IndexInfo()
    assert ffi.kjarni_index_info(root.encode(), C.byref(info)) == 0
    assert (info.document_count, info.segment_count, info.dimension) == (8, 2, 16) and info.path.decode() == root and info.size_bytes > 8 * 16 * 4
    assert info.embedding_model is None  # config.json: "embedding_model": null
    assert ffi.kjarni_index_info(str(tmp_path / "nope").encode(), C.byref(info)) == 3  # ModelNotFound (IndexNotFound)
    assert ffi.kjarni_index_delete(root.encode()) == 0 and not os.path.exists(root)
    assert ffi.kjarni_index_delete(root.encode()) == 3


def test_cancel_token(ffi):
    t = ffi.kjarni_cancel_token_new()
    assert t and not ffi.kjarni_cancel_token_is_cancelled(t)
    ffi.kjarni_cancel_token_cancel(t)
    assert ffi.kjarni_cancel_token_is_cancelled(t)
    ffi.kjarni_cancel_token_reset(t)
    assert not ffi.kjarni_cancel_token_is_cancelled(t)
    ffi.kjarni_cancel_token_free(t)
    assert not ffi.kjarni_cancel_token_is_cancelled(None)


def test_library_exports_every_reference_ffi_symbol():
    """Every export of the reference's kjarni-ffi crate, and every symbol its Python / Go / C# bindings resolve, exists in
    libkjarni_ffi.so (names from tests/golden/ffi_symbols.json, written by scripts/gen_ffi_symbols.py)."""
    sym = json.load(open(os.path.join(HERE, "golden", "ffi_symbols.json")))
    lib = C.CDLL(FFI_SO)
    assert len(sym["rust_exports"]) >= 79
    for group in ("rust_exports", "python_binding", "go_binding", "csharp_binding"):
        missing = [n for n in sym[group] if not hasattr(lib, n)]
        assert not missing, (group, missing)


REF_FFI_PY = "/root/reference/crates/kjarni-ffi/bindings/python/kjarni/_ffi.py"


@pytest.mark.skipif(not os.path.exists(REF_FFI_PY), reason="the reference tree is only present in the build container")
def test_reference_python_binding_imports_against_this_library():
    """The reference's own, unmodified ctypes module resolves all of its symbols against libkjarni_ffi.so."""
    code = ("import importlib.util as u; s = u.spec_from_file_location('_ffi', %r); m = u.module_from_spec(s); s.loader.exec_module(m); "
            "print(m._lib.kjarni_version)" % REF_FFI_PY)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(FFI_SO))
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd="/tmp", capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
