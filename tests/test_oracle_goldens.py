"""Pins the numpy oracle against the reference's own known-answer tests
(tests/golden/reference_kats.json; each block cites its reference file:line)."""
import numpy as np

from oracle import kjarni_oracle as ko

F32 = np.float32


def _det_layer(k):
    """create_deterministic_layer, encoder_layer.rs:244-307: weights count up
    in 0.001 steps across q,k,v,o,fc1,fc2 in that order."""
    h, i = k["hidden"], k["intermediate"]
    count = [1]

    def w(rows, cols):
        n = rows * cols
        a = (np.arange(count[0], count[0] + n, dtype=F32) * F32(k["weight_step"])).reshape(rows, cols)
        count[0] += n
        return a

    b = lambda n: np.full(n, k["bias"], dtype=F32)
    wq, wk, wv, wo = w(h, h), w(h, h), w(h, h), w(h, h)
    w1, w2 = w(i, h), w(h, i)
    g = np.full(h, k["ln_gamma"], dtype=F32)
    be = np.full(h, k["ln_beta"], dtype=F32)
    return ko.LayerWeights(wq, b(h), wk, b(h), wv, b(h), wo, b(h), g, be, w1, b(i), w2, b(h), g, be)


def test_encoder_layer_goldens(kats):
    k = kats["encoder_layer"]
    lw = _det_layer(k)
    x = np.array(k["input"], dtype=F32).reshape(k["batch"], k["seq"], k["hidden"])
    mask = np.array(k["mask"], dtype=F32).reshape(k["batch"], k["seq"])
    pb = np.array(k["pos_bias"], dtype=F32).reshape(1, k["heads"], k["seq"], k["seq"])
    for noalloc in (False, True):
        post = ko.encoder_layer(x, mask, lw, k["heads"], k["eps"], noalloc=noalloc, position_bias=pb)
        pre = ko.encoder_layer(x, mask, lw, k["heads"], k["eps"], noalloc=noalloc, prenorm=True, position_bias=pb)
        assert np.abs(post.ravel() - np.array(k["golden_postnorm"], dtype=F32)).max() < k["tol"]
        assert np.abs(pre.ravel() - np.array(k["golden_prenorm"], dtype=F32)).max() < k["tol"]


def test_softmax_goldens(kats):
    for c in kats["softmax"]["cases"]:
        out = ko.softmax_rows(np.array(c["input"], dtype=F32))
        assert np.abs(out - np.array(c["golden"], dtype=F32)).max() < c["tol"]
        assert abs(float(out.sum()) - 1.0) < 1e-6
    c = kats["softmax"]["case_4d"]
    out = ko.softmax_rows(np.array(c["input"], dtype=F32).reshape(c["shape"]))
    assert np.abs(out.ravel() - np.array(c["golden"], dtype=F32)).max() < c["tol"]
    assert ko.softmax_rows(np.zeros((0,), dtype=F32)).shape == (0,)


def test_softmax_fully_masked_rows():
    """all -1e9 -> uniform; all -inf -> NaN (sum>0 guard), activations.rs:223-242."""
    u = ko.softmax_rows(np.full((4,), -1e9, dtype=F32))
    assert np.allclose(u, 0.25)
    n = ko.softmax_rows(np.full((4,), -np.inf, dtype=F32))
    assert np.isnan(n).all()


def test_gelu_scalars(kats):
    k = kats["gelu_scalars"]
    for x, y in k["erf"]:
        assert abs(float(ko.gelu_erf(np.array([x], dtype=F32))[0]) - y) < k["tol"]
    for x, y in k["tanh"]:
        assert abs(float(ko.gelu_tanh(np.array([x], dtype=F32))[0]) - y) < k["tol"]


def test_ffn_gelu_golden(kats):
    k = kats["ffn_gelu"]
    x = np.array(k["input"], dtype=F32).reshape(-1, 3)
    w1 = np.array(k["w1"], dtype=F32).reshape(k["w1_shape"])
    w2 = np.array(k["w2"], dtype=F32).reshape(k["w2_shape"])
    z = np.zeros
    lw = ko.LayerWeights(None, None, None, None, None, None, None, None, None, None, w1, z(4, F32), w2, z(3, F32), None, None)
    out = ko.feed_forward(x, lw, "gelu")
    assert np.abs(out.ravel() - np.array(k["golden"], dtype=F32)).max() < k["tol"]


def test_layer_norm_kats(kats):
    for c in kats["layer_norm"]["cases"]:
        out = ko.layer_norm(np.array(c["input"], dtype=F32)[None, :], np.array(c["gamma"], dtype=F32),
                            np.array(c["beta"], dtype=F32), c["eps"])
        assert np.abs(out.ravel() - np.array(c["golden"], dtype=F32)).max() < c["tol"]


def test_pooling_goldens(kats):
    k = kats["pooling"]
    h = np.array(k["hidden"], dtype=F32).reshape(k["shape"])
    m = np.array(k["mask"], dtype=F32).reshape(k["shape"][0], k["shape"][1])
    tol = k["tol"]
    assert np.abs(ko.mean_pool(h, m).ravel() - np.array(k["mean"], dtype=F32)).max() < tol
    assert np.abs(ko.cls_pool(h).ravel() - np.array(k["cls"], dtype=F32)).max() < tol
    assert np.abs(ko.max_pool(h, m).ravel() - np.array(k["max"], dtype=F32)).max() < tol
    assert np.abs(ko.last_token_pool(h, m).ravel() - np.array(k["last"], dtype=F32)).max() < tol
    assert np.abs(ko.l2_normalize(ko.mean_pool(h, m)).ravel() - np.array(k["mean_l2"], dtype=F32)).max() < tol
    c = k["l2_case"]
    out = ko.l2_normalize(np.array(c["input"], dtype=F32).reshape(c["shape"]))
    assert np.abs(out.ravel() - np.array(c["golden"], dtype=F32)).max() < c["tol"]


def test_mean_pool_zero_mask_row_takes_token0():
    """pooling/mod.rs:24-31."""
    h = np.arange(12, dtype=F32).reshape(1, 3, 4)
    out = ko.mean_pool(h, np.zeros((1, 3), dtype=F32))
    assert np.array_equal(out[0], h[0, 0])
    z = ko.l2_normalize(np.zeros((1, 4), dtype=F32))
    assert np.array_equal(z, np.zeros((1, 4), dtype=F32))


def test_head_goldens(kats):
    k = kats["heads"]
    for name in ("bert", "distilbert"):
        c = k[name]
        out = ko.classification_head(
            np.array(c["input"], dtype=F32).reshape(c["input_shape"]), c["kind"],
            np.array(c["w_pre"], dtype=F32).reshape(c["w_pre_shape"]), np.array(c["b_pre"], dtype=F32),
            np.array(c["w_cls"], dtype=F32).reshape(c["w_cls_shape"]), np.array(c["b_cls"], dtype=F32))
        assert np.abs(out.ravel() - np.array(c["golden"], dtype=F32)).max() < k["tol"], name


def test_vector_store_kats(kats):
    k = kats["vector_store"]
    for c in k["cosine"]:
        assert abs(ko.cosine_similarity(c["a"], c["b"]) - c["golden"]) < k["tol"]
    c = k["search_sorted"]
    r = ko.vector_store_search(np.array(c["rows"], dtype=F32), c["query"], c["limit"])
    assert [i for i, _ in r] == c["order"]
    assert r[0][1] >= r[1][1] >= r[2][1]
    for name in ("search_limit", "search_limit_exceeds"):
        c = k[name]
        assert len(ko.vector_store_search(np.array(c["rows"], dtype=F32), c["query"], c["limit"])) == c["count"]
    assert ko.vector_store_search(np.zeros((0, 3), dtype=F32), [1, 2, 3], 10) == []
    assert ko.vector_store_search(np.ones((1, 3), dtype=F32), [1, 2], 10) == []


def test_segment_and_index_semantics():
    """KR/segment.rs:307-370 and KR/index_reader.rs:207-228,313-319."""
    rng = np.random.default_rng(0)
    segs = [rng.standard_normal((7, 8)).astype(F32), rng.standard_normal((5, 8)).astype(F32)]
    segs[0][3] = 0  # zero row scores 0
    q = rng.standard_normal(8).astype(F32)
    res = ko.index_search_semantic(segs, q, 4)
    allrows = np.concatenate(segs)
    ids, sc = ko.batched_topk(allrows, q[None], 4)
    assert [i for i, _ in res] == ids[0].tolist()
    assert np.allclose([s for _, s in res], sc[0], atol=1e-6)
    assert ko.segment_search(segs[0], np.zeros(8, dtype=F32), 3) == []
    assert ko.segment_search(segs[0], np.ones(5, dtype=F32), 3) == []
    s = ko.segment_scores(segs[0], q)
    assert s[3] == 0.0
    # ties resolve to the lowest id (stable sort)
    dup = np.stack([q, q, q * 2]).astype(F32)
    assert [i for i, _ in ko.segment_search(dup, q, 2)][0] in (0, 1, 2)
    assert ko.stable_argsort_desc(np.array([1.0, 1.0, 0.5], dtype=F32)).tolist() == [0, 1, 2]


def test_embeddings_forward_reference_vectors():
    """Embeddings::forward known answers transcribed from the reference's own unit tests
    (kjarni-transformers/src/cpu/embeddings/tests.rs:59-220): word + position (+ offset, clamped to the table) + token type."""
    ones = lambda r, c: np.full((r, c), 1.0, np.float32)
    # test_cpu_embeddings_math / test_position_embedding_broadcasting_batch / test_batch_sequence_broadcasting
    for (b, s, h) in ((1, 2, 4), (2, 3, 4), (3, 4, 2)):
        out = ko.embeddings_forward(np.zeros((b, s), np.uint32), ones(10, h), np.full((10, h), 0.5, np.float32), None, None, 0)
        assert out.shape == (b, s, h) and (out == 1.5).all()
    # test_position_embedding_with_offset: pos[i, j] = 0.1 i + j, offset 2
    pos = np.fromfunction(lambda i, j: i * np.float32(0.1) + j, (10, 2), dtype=np.float32).astype(np.float32)
    out = ko.embeddings_forward(np.zeros((1, 4), np.uint32), ones(10, 2), pos, None, None, 2)
    assert out[0, 0, 0] == np.float32(1.0) + np.float32(2 * np.float32(0.1)) and out[0, 0, 1] == np.float32(1.0) + (np.float32(2 * np.float32(0.1)) + 1)
    assert out[0, 3, 0] == np.float32(1.0) + np.float32(0.5) and out[0, 3, 1] == np.float32(1.0) + np.float32(1.5)
    # test_token_type_embeddings: type[i, j] = i + 0.1 j, all type ids 1
    typ = np.fromfunction(lambda i, j: i + np.float32(0.1) * j, (2, 3), dtype=np.float32).astype(np.float32)
    out = ko.embeddings_forward(np.zeros((1, 2), np.uint32), ones(5, 3), None, typ, np.ones((1, 2), np.uint32), 0)
    for h in range(3):
        assert (out[:, :, h] == np.float32(1.0) + typ[1, h]).all()
    # test_position_offset_clamping: a 3-row position table, offset 1, sequence 4 -> only two positions receive a row
    out = ko.embeddings_forward(np.zeros((1, 4), np.uint32), ones(5, 2), np.full((3, 2), 0.5, np.float32), None, None, 1)
    assert out[0, 0, 0] == 1.5 and out[0, 1, 0] == 1.5 and out[0, 2, 0] == 1.0 and out[0, 3, 0] == 1.0
