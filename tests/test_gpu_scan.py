"""Cosine top-k scan parity through the C ABI: ids bit-exact against the fp32 oracle
(ties within 1e-6 exempt), scores to fp32 summation-order accuracy."""
import os

import numpy as np
import pytest

from kjarni_b200 import _native as N
from kjarni_b200 import api
from oracle import kjarni_oracle as ko

pytestmark = pytest.mark.gpu


def check_topk(ids, sc, want_ids, want_sc, all_scores=None):
    assert np.abs(sc - want_sc).max() < 2e-6
    for qi in range(ids.shape[0]):
        if (ids[qi] == want_ids[qi].astype(np.uint64)).all():
            continue
        # any disagreement must be a tie within 1e-6 (north_star exemption)
        for j in range(ids.shape[1]):
            if ids[qi, j] != np.uint64(want_ids[qi, j]):
                assert abs(float(sc[qi, j]) - float(want_sc[qi, j])) <= 1e-6


@pytest.mark.parametrize("n,dim,nq,k", [(1000, 384, 1, 10), (5000, 384, 8, 10), (3000, 768, 3, 5), (257, 64, 13, 32),
                                        (20000, 384, 5, 100), (64, 128, 2, 10), (7, 16, 1, 10)])
def test_search_matches_oracle(n, dim, nq, k):
    rows = ko.synth_rows(7, 0, n, dim)
    q = ko.synth_rows(11, 0, nq, dim)
    sh = api.IndexShard(dim, n, id_base=1000)
    sh.add_rows(rows)
    assert len(sh) == n
    wi, ws = ko.batched_topk(rows, q, k, row_offset=1000)
    kk = min(k, n)
    for min_q in (1, 1 << 30):  # default dispatch (tensor-core filter where it applies), then the exact scan kernel alone
        sh.set_filter(min_queries=min_q)
        ids, sc, cnt = sh.search_batch(q, k)
        assert (cnt == kk).all()
        check_topk(ids[:, :kk], sc[:, :kk], wi[:, :kk], ws[:, :kk])
        assert (ids[:, kk:] == np.uint64(N.NO_ID)).all() and np.isneginf(sc[:, kk:]).all()
    sh.close()


def test_device_generator_matches_oracle_rows():
    sh = api.IndexShard(384, 5000)
    sh.append_synthetic(7, 123456, 5000)
    got = sh.get_rows(0, 5000)
    assert np.array_equal(got, ko.synth_rows(7, 123456, 5000, 384))
    assert np.array_equal(sh.get_embedding(17), got[17])
    sh.close()


def test_reference_vector_store_kats(kats):
    """The reference's own VectorStore tests (kjarni-search/src/vector.rs:201-309) through the CUDA path."""
    vs = api.VectorStore(4, 16)
    vs.add_rows(np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0.9, 0.1, 0, 0], [0, 0, 0, 0]], np.float32))
    r = vs.search([1, 0, 0, 0], 2)
    assert [i for i, _ in r] == [0, 2] and abs(r[0][1] - 1.0) < 1e-6
    assert vs.search([1, 0, 0], 2) == []  # dimension mismatch
    r = vs.search([0, 0, 0, 0], 4)  # zero query: all scores 0 (denominator clamps to 1e-9), stable order
    assert [i for i, _ in r] == [0, 1, 2, 3] and all(s == 0.0 for _, s in r)
    assert api.VectorStore(4, 4).search([1, 0, 0, 0], 2) == []  # empty store
    vs.close()


def test_segment_semantics_and_merge():
    rng = np.random.default_rng(5)
    dim = 32
    segs = [rng.standard_normal((n, dim)).astype(np.float32) for n in (50, 1, 120)]
    segs[0][3] = 0.0  # zero-norm row scores 0
    segs[2][7] = segs[0][5]  # exact duplicate across segments: tie broken by the lower global id
    shards, off = [], 0
    for s in segs:
        sh = api.IndexShard(dim, len(s), id_base=off)
        sh.add_rows(s)
        shards.append(sh)
        off += len(s)
    rd = api.IndexReader(shards)
    for qi in range(5):
        q = rng.standard_normal(dim).astype(np.float32) if qi else segs[0][5].copy()
        want = ko.index_search_semantic(segs, q, 10)
        got = rd.search_semantic(q, 10)
        assert [g[0] for g in got] == [w[0] for w in want]
        assert np.allclose([g[1] for g in got], [w[1] for w in want], atol=2e-6)
    assert shards[0].search_vectors(np.zeros(dim, np.float32), 5) == []  # |q| < 1e-9 -> []
    assert shards[0].search_vectors(np.zeros(dim + 1, np.float32), 5) == []
    for sh in shards:
        sh.close()


def test_vectors_bin_segment_file(tmp_path):
    rows = ko.synth_rows(3, 0, 999, 384)
    p = tmp_path / "vectors.bin"
    rows.astype("<f4").tofile(p)
    sh = api.IndexShard(384, 2000)
    sh.load_vectors_bin(str(p))
    sh.load_vectors_bin(str(p))  # appending a second segment file
    assert len(sh) == 1998
    q = ko.synth_rows(11, 5, 1, 384)[0]
    got = sh.search_vectors(q, 4)
    want = ko.segment_search(np.concatenate([rows, rows]), q, 4)
    assert [g[0] for g in got] == [w[0] for w in want]  # duplicates: lower id first
    with pytest.raises(N.KjarniCudaError):
        sh.load_vectors_bin(str(tmp_path / "missing.bin"))
    sh.close()


def test_full_size_shard_properties():
    """BASELINE config-4 shard shape at reduced row count that still exceeds L2 (1M x 384 fp32 = 1.5 GB):
    planted near-duplicates must come back first, and the result must be invariant to how rows are split
    into shards (per-shard top-k + merge == single-shard top-k)."""
    n, dim, k = 1_000_000, 384, 10
    big = api.IndexShard(dim, n + 16)
    big.append_synthetic(7, 0, n)
    q = ko.synth_rows(11, 0, 4, dim)
    planted = np.stack([q[i] * (1.0 + 0.01 * j) + 1e-3 * j for i in range(2) for j in range(3)]).astype(np.float32)
    big.add_rows(planted)  # local ids n .. n+5
    ids, sc, cnt = big.search_batch(q, k)
    assert (cnt == k).all()
    assert set(ids[0, :3].tolist()) == {n, n + 1, n + 2} and set(ids[1, :3].tolist()) == {n + 3, n + 4, n + 5}
    assert (np.diff(sc, axis=1) <= 0).all()
    # oracle on the candidate set: returned scores are the true cosines of the returned rows
    for qi in range(4):
        rows = np.stack([big.get_embedding(int(i)) for i in ids[qi]])
        s = ko.segment_scores(rows, q[qi])
        assert np.abs(s - sc[qi]).max() < 2e-6
    # sampled block: nothing in it beats the k-th score unless it was returned
    blk = big.get_rows(123_000, 20_000)
    for qi in range(4):
        s = ko.segment_scores(blk, q[qi])
        better = np.nonzero(s > sc[qi, -1] + 1e-6)[0] + 123_000
        assert set(better.tolist()) <= set(ids[qi].tolist())
    # shard invariance
    halves = [api.IndexShard(dim, n // 2 + 16, id_base=0), api.IndexShard(dim, n // 2 + 16, id_base=n // 2)]
    halves[0].append_synthetic(7, 0, n // 2)
    halves[1].append_synthetic(7, n // 2, n // 2)
    halves[1].add_rows(planted)
    for qi in range(4):
        merged = api.IndexReader(halves).search_semantic(q[qi], k)
        assert [m[0] for m in merged] == ids[qi].tolist()
    for h in halves + [big]:
        h.close()


# ------------------------------------------------------------------ tensor-core filter path (scan_gemm.cuh)
@pytest.mark.parametrize("n,dim,nq,k", [(5000, 384, 9, 10), (70000, 384, 130, 10), (40000, 128, 300, 16), (300, 64, 17, 1),
                                        (20, 384, 12, 10), (257, 256, 128, 5), (100001, 320, 64, 10),
                                        # round 2: 16 < k <= 64 keeps 128 approximate candidates per query (a reranking Searcher fetches top_k * 5 = 50),
                                        # dim > 384 streams the query tile k-block by k-block (768-dim embedders)
                                        (70000, 384, 130, 50), (5000, 384, 9, 64), (12000, 64, 33, 20), (60000, 768, 40, 10), (30000, 1024, 9, 50),
                                        (100000, 768, 200, 50), (3000, 448, 5, 17)])
def test_gemm_filter_matches_oracle_and_exact_path(n, dim, nq, k):
    rows = ko.synth_rows(7, 0, n, dim)
    q = ko.synth_rows(11, 0, nq, dim)
    q[nq // 2] = rows[n // 3] * 1.5  # a query with an exact-duplicate direction in the index
    sh = api.IndexShard(dim, n + 3, id_base=500)
    sh.add_rows(rows)
    sh.add_rows(np.stack([rows[n // 3], np.zeros(dim, np.float32), rows[n // 3] * 0.5]))  # duplicates (ties -> lower id) + a zero row
    allrows = np.concatenate([rows, rows[n // 3][None], np.zeros((1, dim), np.float32), rows[n // 3][None] * 0.5])
    ids, sc, cnt = sh.search_batch(q, k)  # filter path
    assert sh.last_launch_count >= 5  # prep, (seed,) seed select, filter GEMM, candidate select, exact rescoring -- not the 3-launch exact scan
    wi, ws = ko.batched_topk(allrows, q, k, row_offset=500)
    kk = min(k, n + 3)
    assert (cnt == kk).all()
    check_topk(ids[:, :kk], sc[:, :kk], wi[:, :kk], ws[:, :kk])
    # the filter path must return exactly what the exact scan returns (same fp32 arithmetic for the final scores)
    sh.set_filter(min_queries=1 << 30)
    ids2, sc2, cnt2 = sh.search_batch(q, k)
    assert np.array_equal(ids, ids2) and np.array_equal(sc, sc2) and np.array_equal(cnt, cnt2)
    # forcing the proof to fail sends every query through the exact re-run: same answer again
    sh.set_filter(eps=10.0, min_queries=1)
    ids3, sc3, cnt3 = sh.search_batch(q, k)
    assert np.array_equal(ids, ids3) and np.array_equal(sc, sc3) and np.array_equal(cnt, cnt3)
    sh.close()


def test_gemm_filter_zero_query_and_modes():
    n, dim, nq, k = 3000, 384, 20, 10
    rows = ko.synth_rows(7, 0, n, dim)
    q = ko.synth_rows(11, 0, nq, dim)
    q[3] = 0.0
    sh = api.IndexShard(dim, n)
    sh.add_rows(rows)
    ids, sc, cnt = sh.search_batch(q, k, N.SCAN_SEGMENT)
    assert cnt[3] == 0 and (ids[3] == np.uint64(N.NO_ID)).all() and np.isneginf(sc[3]).all()  # |q| < 1e-9 -> [] (segment.rs:312-314)
    assert (np.delete(cnt, 3) == k).all()
    ids_v, sc_v, cnt_v = sh.search_batch(q, k, N.SCAN_VECTORSTORE)
    assert cnt_v[3] == k and ids_v[3].tolist() == list(range(k)) and (sc_v[3] == 0).all()  # all scores 0 -> stable order = id asc
    keep = np.arange(nq) != 3
    assert np.array_equal(ids[keep], ids_v[keep])
    sh.close()


def test_gemm_filter_async_device_api_large_batch():
    """BASELINE config 4 query-batch shape (4096 queries, k = 10) on a 1M-row shard through the device-pointer entry point:
    every result proven exact by the filter's bound, spot-checked against the exact scan and the oracle."""
    import ctypes as C

    import torch

    n, dim, nq, k = 1_000_000, 384, 4096, 10
    sh = api.IndexShard(dim, n, id_base=0)
    sh.append_synthetic(7, 0, n)
    q = ko.synth_rows(11, 0, nq, dim)
    dq = torch.from_numpy(q).cuda()
    d_ids = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    d_sc = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    d_cnt = torch.empty((nq,), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    N.check(N.lib().kjc_index_search_device_async(sh._h, dq.data_ptr(), nq, k, N.SCAN_SEGMENT, d_ids.data_ptr(), d_sc.data_ptr(),
                                                  d_cnt.data_ptr(), C.c_void_p(st)))
    torch.cuda.synchronize()
    assert sh.unverified_count == 0
    ids = d_ids.cpu().numpy().astype(np.uint64)
    sc = d_sc.cpu().numpy()
    assert (d_cnt.cpu().numpy() == k).all() and (np.diff(sc, axis=1) <= 0).all()
    sel = [0, 1, 127, 128, 2047, 4095]
    sh.set_filter(min_queries=1 << 30)
    ids_e, sc_e, _ = sh.search_batch(q[sel], k)
    assert np.array_equal(ids[sel], ids_e) and np.array_equal(sc[sel], sc_e)
    blk = sh.get_rows(500_000, 50_000)
    for qi in sel[:3]:
        s = ko.segment_scores(blk, q[qi])
        better = np.nonzero(s > sc[qi, -1] + 1e-6)[0] + 500_000
        assert set(better.tolist()) <= set(ids[qi].tolist())
    sh.close()


def test_unproven_queries_escalate_before_the_exact_scan(monkeypatch):
    """A query whose 32-candidate proof fails is first re-selected with 128 candidates from the SAME filter buffer (it holds every row
    above the seed bound) and re-scored; only what is still unproven goes to the exact scan.  With a deliberately wide error bound
    many proofs fail: the answers must equal the exact scan's bit for bit, with far fewer launches than going straight to it."""
    n, dim, nq, k = 400_000, 384, 512, 10
    q = ko.synth_rows(11, 0, nq, dim)
    res = {}
    for esc in (True, False):
        if esc:
            monkeypatch.delenv("KJC_SCAN_NO_ESCALATE", raising=False)
        else:
            monkeypatch.setenv("KJC_SCAN_NO_ESCALATE", "1")
        sh = api.IndexShard(dim, n)
        sh.append_synthetic(7, 0, n)
        sh.set_filter(eps=0.012, min_queries=1)
        ids, sc, cnt = sh.search_batch(q, k)
        res[esc] = (ids, sc, cnt, sh.last_launch_count)
        if esc:
            sh.set_filter(min_queries=1 << 30)
            ids_e, sc_e, cnt_e = sh.search_batch(q[:64], k)
            assert np.array_equal(ids[:64], ids_e) and np.array_equal(sc[:64], sc_e) and np.array_equal(cnt[:64], cnt_e)
        sh.close()
    assert np.array_equal(res[True][0], res[False][0]) and np.array_equal(res[True][1], res[False][1])
    assert res[False][3] > 6 + 3 * 4, res[False][3]        # the wide bound really left several exact passes to do ...
    assert res[True][3] < res[False][3] // 2, (res[True][3], res[False][3])  # ... and the escalation resolved most of them
