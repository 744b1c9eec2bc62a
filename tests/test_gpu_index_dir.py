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
