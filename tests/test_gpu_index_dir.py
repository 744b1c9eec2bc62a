"""On-disk index directory -> GPU shards through the C ABI: rows byte-identical to vectors.bin,
global ids = IndexReader::local_to_global (kjarni-rag/src/index_reader.rs:313-319), top-k identical to the oracle's
IndexReader::search_semantic (index_reader.rs:207-228) for one shard and for several contiguous parts merged."""
import numpy as np
import pytest

from kjarni_b200 import api, synth
from oracle import kjarni_oracle as ko

pytestmark = pytest.mark.gpu

LENS = [700, 1, 2300, 999, 4000]


def make_dir(tmp_path, lens=LENS, dim=384, **kw):
    segs = [ko.synth_rows(7, sum(lens[:i]), n, dim) for i, n in enumerate(lens)]
    return synth.write_index_dir(str(tmp_path / "idx"), segs, dimension=dim, **kw), segs


def same_topk(got, want):
    assert len(got) == len(want)
    for (gi, gs), (wi, ws) in zip(got, want):
        assert abs(gs - ws) < 2e-6
        assert gi == wi or abs(gs - ws) <= 1e-6  # ties within 1e-6 exempt (north_star)


def test_open_dir_loads_every_segment_in_order(tmp_path):
    root, segs = make_dir(tmp_path)
    sh = api.IndexShard.open_dir(root)
    allrows = np.concatenate(segs)
    assert len(sh) == allrows.shape[0] and sh.dim == 384 and sh.id_base == 0
    for r in (0, 699, 700, 701, 3000, len(sh) - 1):  # Segment::get_embedding across segment boundaries
        assert np.array_equal(sh.get_embedding(r), allrows[r])
    q = ko.synth_rows(11, 0, 6, 384)
    for qi in range(q.shape[0]):
        same_topk(sh.search_vectors(q[qi], 10), ko.index_search_semantic(segs, q[qi], 10))
    sh.close()


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_parts_merge_to_the_reference_result(tmp_path, parts):
    root, segs = make_dir(tmp_path)
    rd = api.IndexReader.open(root, devices=[0] * parts)  # the row-sharding of SURVEY 8e, all parts on one GPU here
    assert len(rd) == sum(LENS)
    bases = [s.id_base for s in rd.shards]
    assert bases == [api.index_part_range(sum(LENS), p, parts)[0] for p in range(parts)]
    q = ko.synth_rows(11, 100, 5, 384)
    for qi in range(q.shape[0]):
        same_topk(rd.search_semantic(q[qi], 10), ko.index_search_semantic(segs, q[qi], 10))
    # batched path: per-part candidates merged by the device merge kernel's host twin (order: score desc, id asc)
    allrows = np.concatenate(segs)
    wi, ws = ko.batched_topk(allrows, q, 10)
    cand_i = np.concatenate([s.search_batch(q, 10)[0] for s in rd.shards], axis=1)
    cand_s = np.concatenate([s.search_batch(q, 10)[1] for s in rd.shards], axis=1)
    for qi in range(q.shape[0]):
        order = np.lexsort((cand_i[qi], -cand_s[qi]))[:10]
        assert np.abs(cand_s[qi][order] - ws[qi]).max() < 2e-6
        assert (cand_i[qi][order] == wi[qi].astype(np.uint64)).all()
    for s in rd.shards:
        s.close()


def test_skipped_segment_shifts_global_ids_like_the_reference(tmp_path):
    root, segs = make_dir(tmp_path, lens=[300, 200, 500], broken=(1,))
    sh = api.IndexShard.open_dir(root)
    kept = [segs[0], segs[2]]
    assert len(sh) == 800
    q = ko.synth_rows(11, 7, 3, 384)
    for qi in range(3):
        same_topk(sh.search_vectors(q[qi], 5), ko.index_search_semantic(kept, q[qi], 5))
    sh.close()


def test_truncated_vectors_bin_is_a_load_error(tmp_path):
    import os

    from kjarni_b200 import _native as N

    root, _ = make_dir(tmp_path, lens=[64, 64], dim=64)
    p = os.path.join(root, "segments", "seg_000001", "vectors.bin")
    os.truncate(p, 64 * 64 * 4 - 256)
    with pytest.raises(N.KjarniCudaError) as e:
        api.IndexShard.open_dir(root)
    assert e.value.status == N.KJC_LOAD_FAILED


def test_reference_rag_full_lifecycle(tmp_path):
    """`test_rag_full_lifecycle` of the reference (kjarni-rag/src/tests.rs:8-79), transcribed: three documents with dimension 4 and
    max_docs_per_segment 2 (=> two segments), exact vectors and texts as in the reference test.  Semantic search runs on the GPU
    shard, keyword search through kjarni_search_keywords (the segments' bm25.bin), hybrid = reciprocal-rank fusion of the two
    (index_reader.rs:248-289, hybrid.rs:3-31)."""
    import ctypes as C
    import json
    import os

    from kjarni_b200 import _native as N

    docs = [["Apple is a fruit", "Car is a vehicle"], ["Banana is yellow"]]
    vecs = [np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0]], np.float32), np.array([[0.9, 0.1, 0.0, 0.0]], np.float32)]
    meta = [[{"category": "fruit"}, {}], [{}]]
    root = synth.write_index_dir(str(tmp_path / "my_index"), vecs, dimension=4, max_docs_per_segment=2, docs=docs, metadata=meta)
    flat = docs[0] + docs[1]
    info = N.KjcIndexDirInfo()
    N.check(N.lib().kjc_index_dir_info(root.encode(), C.byref(info)))
    assert (info.total_rows, info.dimension, info.n_segments) == (3, 4, 2)  # reader.len() / dimension() / segment_count()
    sh = api.IndexShard.open_dir(root)
    sem = sh.search_vectors(np.array([1.0, 0.0, 0.0, 0.0], np.float32), 10)
    assert len(sem) == 3
    assert [flat[i] for i, _ in sem] == ["Apple is a fruit", "Banana is yellow", "Car is a vehicle"]
    assert sem[0][1] > 0.99
    sh.close()

    class SearchResult(C.Structure):
        _fields_ = [("score", C.c_float), ("document_id", C.c_size_t), ("text", C.c_char_p), ("metadata_json", C.c_char_p)]

    class SearchResults(C.Structure):
        _fields_ = [("results", C.POINTER(SearchResult)), ("len", C.c_size_t)]

    ffi = C.CDLL(os.path.join(os.path.dirname(N.LIB_PATH), "libkjarni_ffi.so"))
    ffi.kjarni_search_keywords.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(SearchResults)]
    ffi.kjarni_search_results_free.argtypes = [C.POINTER(SearchResults)]

    def keywords(q, k):
        res = SearchResults()
        assert ffi.kjarni_search_keywords(root.encode(), q.encode(), k, C.byref(res)) == 0
        out = [(res.results[i].document_id, res.results[i].score, res.results[i].text.decode(), json.loads(res.results[i].metadata_json)) for i in range(res.len)]
        ffi.kjarni_search_results_free(C.byref(res))
        return out

    kw = keywords("yellow", 10)
    assert len(kw) == 1 and kw[0][2] == "Banana is yellow"
    # search_hybrid("vehicle", [1,0,0,0], 10): keyword and semantic lists of 2 x limit each, fused
    hyb = ko.rrf_hybrid([(d, s) for d, s, _, _ in keywords("vehicle", 20)], sem, 10)
    top2 = [flat[i] for i, _ in hyb[:2]]
    assert "Car is a vehicle" in top2 and "Apple is a fruit" in top2
    apple = [r for r in keywords("apple", 10) if "Apple" in r[2]][0]
    assert apple[3].get("category") == "fruit"
