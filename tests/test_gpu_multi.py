"""Multi-GPU behind the C ABI (kjc_encoder_create_multi, kjc_sharded_index_*): one process, one host thread per GPU, no torch.
On a box with one GPU the device list names it several times, which exercises the same split / peer-copy gather / merge code;
with >= 2 GPUs the replicas and shards sit on distinct devices."""
import ctypes as C
import os

import numpy as np
import pytest

from kjarni_b200 import _native as N
from kjarni_b200 import api, synth
from oracle import kjarni_oracle as ko

pytestmark = pytest.mark.gpu


def device_lists():
    n = N.lib().kjc_device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists.append([0, 1])
    if n >= 4:
        lists.append([0, 1, 2, 3])
    return lists


@pytest.fixture(scope="module")
def model_dirs(tmp_path_factory):
    root = tmp_path_factory.mktemp("models")
    return {a: synth.write_model_dir(str(root / a), a) for a in ("tiny-bert", "tiny-distilbert", "minilm-l6")}


@pytest.mark.parametrize("arch,B,S", [("tiny-bert", 37, 16), ("tiny-bert", 1, 16), ("minilm-l6", 301, 128), ("tiny-distilbert", 11, 24)])
def test_encoder_group_rows_equal_single_gpu(model_dirs, arch, B, S):
    """A batch split over replicas gives bit-identical rows to the one-GPU forward (sequences are independent; the mask convention is
    resolved for the whole batch), for pooled embeddings, hidden states and logits."""
    vocab = synth.ARCHS[arch][5]
    ids, mask, _ = synth.synth_tokens(B, S, vocab, regime="P", seed=5)
    one = api.EncoderModel(model_dirs[arch])
    want = one.encode_batch_from_ids(ids, mask)
    want_h = one.get_hidden_states_batch_from_ids(ids, mask) if B <= 64 else None
    want_l = one.predict_logits(ids, mask) if one.num_labels else None
    one.close()
    for devs in device_lists():
        grp = api.EncoderModel(model_dirs[arch], devices=devs)
        assert grp.n_devices == len(devs)
        got = grp.encode_batch_from_ids(ids, mask)
        assert np.array_equal(got, want), devs
        if want_h is not None:
            assert np.array_equal(grp.get_hidden_states_batch_from_ids(ids, mask), want_h)
        if want_l is not None:
            assert np.array_equal(grp.predict_logits(ids, mask), want_l)
        grp.close()
    # and against the oracle, through the multi-replica handle
    if arch == "tiny-bert":
        grp = api.EncoderModel(model_dirs[arch], devices=device_lists()[-1])
        ref = ko.embed(ko.load_model_dir(model_dirs[arch]), ids, mask)
        got = grp.encode_batch_from_ids(ids, mask)
        cos = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
        assert cos.min() >= 0.9995 and np.abs(got - ref).max() <= 2e-2
        grp.close()


@pytest.mark.parametrize("n,dim,nq,k", [(5000, 384, 7, 10), (20001, 384, 130, 10), (3000, 128, 3, 50), (40, 64, 2, 10)])
def test_sharded_index_matches_oracle_and_single_shard(n, dim, nq, k):
    rows = ko.synth_rows(3, 0, n, dim)
    rows[17] = rows[5]  # duplicate rows in different shards at large n: tie -> lower global id first
    rows[n - 1] = rows[5]
    q = ko.synth_rows(9, 0, nq, dim)
    q[0] = rows[5]
    wi, ws = ko.batched_topk(rows, q, k)
    single = api.IndexShard(dim, n)
    single.add_rows(rows)
    si, ss, sc = single.search_batch(q, k)
    single.close()
    for devs in device_lists():
        sh = api.ShardedIndex(dim, n, devs)
        half = n // 3
        sh.add_rows(rows[:half])  # appended in two calls: rows are routed to the shard that owns their global id
        sh.add_rows(rows[half:])
        assert len(sh) == n and sum(sh.shard_lens) == n and max(sh.shard_lens) - min(sh.shard_lens) <= 1
        gi, gs, gc = sh.search_batch(q, k)
        kk = min(k, n)
        assert (gi[:, :kk] == wi[:, :kk].astype(np.uint64)).all(), devs
        assert np.array_equal(gi, si) and np.array_equal(gs, ss) and np.array_equal(gc, sc)
        assert np.abs(gs[:, :kk] - ws[:, :kk]).max() < 2e-6
        sh.close()


def test_sharded_index_zero_norm_query_and_synthetic_rows():
    dim, n = 384, 9000
    devs = device_lists()[-1]
    sh = api.ShardedIndex(dim, n, devs)
    sh.append_synthetic(7, 4000)
    sh.append_synthetic(7, 5000)
    rows = ko.synth_rows(7, 0, n, dim)
    q = np.concatenate([rows[[123, 8999]], np.zeros((1, dim), np.float32)])
    ids, sc, cnt = sh.search_batch(q, 5)
    assert ids[0, 0] == 123 and ids[1, 0] == 8999 and abs(sc[0, 0] - 1) < 1e-5
    assert cnt[2] == 0 and (ids[2] == np.iinfo(np.uint64).max).all()  # Segment::search_vectors: zero-norm query -> no results
    wi, _ = ko.batched_topk(rows, q[:2], 5)
    assert (ids[:2] == wi.astype(np.uint64)).all()
    sh.close()


def test_sharded_index_open_dir(tmp_path):
    dim = 64
    segs = [ko.synth_rows(1, 0, 700, dim), ko.synth_rows(1, 700, 45, dim), ko.synth_rows(1, 745, 1300, dim)]
    root = synth.write_index_dir(str(tmp_path / "idx"), segs)
    allrows = np.concatenate(segs)
    q = ko.synth_rows(2, 0, 6, dim)
    want = [ko.index_search_semantic(segs, q[i], 10) for i in range(q.shape[0])]
    for devs in device_lists():
        sh = api.ShardedIndex.open_dir(root, devs)
        assert len(sh) == allrows.shape[0] and sh.dim == dim
        ids, sc, cnt = sh.search_batch(q, 10)
        for i in range(q.shape[0]):
            assert [int(x) for x in ids[i]] == [d for d, _ in want[i]]
        sh.close()


def test_public_abi_uses_every_listed_gpu(tmp_path, monkeypatch):
    """KJARNI_GPU_DEVICES routes the public kjarni_* ABI through the replica group / sharded index; results equal the default."""
    import shutil

    ffi_so = os.path.join(os.path.dirname(N.LIB_PATH), "libkjarni_ffi.so")
    lib = C.CDLL(ffi_so)
    tok = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tokenizers", "bert_uncased.tokenizer.json")
    cache = tmp_path / "cache"
    d = synth.write_model_dir(str(cache / "sentence-transformers_all-MiniLM-L6-v2"), "tiny-bert")
    shutil.copy(tok, os.path.join(d, "tokenizer.json"))

    class EmbedderConfig(C.Structure):
        _fields_ = [("device", C.c_int), ("cache_dir", C.c_char_p), ("model_name", C.c_char_p), ("model_path", C.c_char_p),
                    ("normalize", C.c_int32), ("quiet", C.c_int32)]

    class Float2DArray(C.Structure):
        _fields_ = [("data", C.POINTER(C.c_float)), ("rows", C.c_size_t), ("cols", C.c_size_t)]

    lib.kjarni_embedder_config_default.restype = EmbedderConfig
    lib.kjarni_embedder_encode_batch.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_size_t, C.POINTER(Float2DArray)]
    lib.kjarni_embedder_free.argtypes = [C.c_void_p]
    lib.kjarni_float_2d_array_free.argtypes = [C.POINTER(Float2DArray)]
    texts = ["text number %d about gpus and search" % i for i in range(23)]
    arr = (C.c_char_p * len(texts))(*[t.encode() for t in texts])

    def embed():
        cfg = lib.kjarni_embedder_config_default()
        cfg.device, cfg.cache_dir = 1, str(cache).encode()
        h = C.c_void_p()
        assert lib.kjarni_embedder_new(C.byref(cfg), C.byref(h)) == 0
        out = Float2DArray()
        assert lib.kjarni_embedder_encode_batch(h, arr, len(texts), C.byref(out)) == 0
        a = np.ctypeslib.as_array(out.data, shape=(out.rows, out.cols)).copy()
        lib.kjarni_float_2d_array_free(C.byref(out))
        lib.kjarni_embedder_free(h)
        return a

    monkeypatch.delenv("KJARNI_GPU_DEVICES", raising=False)
    base = embed()
    monkeypatch.setenv("KJARNI_GPU_DEVICES", ",".join(str(x) for x in device_lists()[-1]))
    assert np.array_equal(embed(), base)
    monkeypatch.setenv("KJARNI_GPU_DEVICES", "all")
    assert np.array_equal(embed(), base)
