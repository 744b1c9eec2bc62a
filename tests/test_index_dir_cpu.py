"""Host side of the on-disk index reader (no GPU): config.json + segment table parsing, segment
ordering and skipping as IndexReader::open does (kjarni-rag/src/index_reader.rs:161-204), global-id bookkeeping
(index_reader.rs:313-331) and the contiguous part ranges used to row-shard the index over GPUs."""
import json
import os

import numpy as np
import pytest

from kjarni_b200 import _native as N
from kjarni_b200 import api, synth
from oracle import kjarni_oracle as ko


def make_dir(tmp_path, lens, dim=32, **kw):
    segs = [ko.synth_rows(7, sum(lens[:i]), n, dim) for i, n in enumerate(lens)]
    return synth.write_index_dir(str(tmp_path / "idx"), segs, dimension=dim, **kw), segs


def test_dir_info_orders_segments_and_sums_lengths(tmp_path):
    root, _ = make_dir(tmp_path, [5, 3, 9, 1])
    info = api.index_dir_info(root)
    assert info["dimension"] == 32 and info["n_segments"] == 4 and info["n_skipped"] == 0
    assert info["segment_lens"] == [5, 3, 9, 1] and info["total_rows"] == 18
    assert info["max_docs_per_segment"] == 10_000  # IndexConfig::default, kjarni-rag/src/config.rs:19


def test_broken_segment_is_skipped_like_indexreader_open(tmp_path):
    # Segment::open fails without bm25.bin -> IndexReader::open logs a warning and carries on; later segments shift down
    root, _ = make_dir(tmp_path, [4, 6, 2], broken=(1,))
    info = api.index_dir_info(root)
    assert info["n_segments"] == 2 and info["n_skipped"] == 1
    assert info["segment_lens"] == [4, 2] and info["total_rows"] == 6
    os.remove(os.path.join(root, "segments", "seg_000002", "segment.json"))
    info = api.index_dir_info(root)
    assert info["segment_lens"] == [4] and info["n_skipped"] == 2


def test_empty_index_and_missing_segments_dir(tmp_path):
    root = str(tmp_path / "empty")
    os.makedirs(root)
    with open(os.path.join(root, "config.json"), "w") as f:
        json.dump({"dimension": 384, "max_docs_per_segment": 10000, "max_segment_memory": 1, "embedding_model": None,
                   "model_name": None, "created_at": None, "version": 1}, f)
    info = api.index_dir_info(root)  # `if segments_dir.exists()`: no segments is an empty index, not an error
    assert info["n_segments"] == 0 and info["total_rows"] == 0 and info["dimension"] == 384


def test_errors(tmp_path):
    with pytest.raises(N.KjarniCudaError) as e:
        api.index_dir_info(str(tmp_path / "nope"))
    assert e.value.status == N.KJC_MODEL_NOT_FOUND
    root = str(tmp_path / "bad")
    os.makedirs(root)
    open(os.path.join(root, "config.json"), "w").write("{not json")
    with pytest.raises(N.KjarniCudaError) as e:
        api.index_dir_info(root)
    assert e.value.status == N.KJC_LOAD_FAILED
    open(os.path.join(root, "config.json"), "w").write('{"version": 1}')  # serde: missing field `dimension`
    with pytest.raises(N.KjarniCudaError) as e:
        api.index_dir_info(root)
    assert e.value.status == N.KJC_LOAD_FAILED
    root2, _ = make_dir(tmp_path, [3, 3])
    meta_p = os.path.join(root2, "segments", "seg_000001", "segment.json")
    meta = json.load(open(meta_p))
    meta["dimension"] = 16
    json.dump(meta, open(meta_p, "w"))
    with pytest.raises(N.KjarniCudaError) as e:
        api.index_dir_info(root2)
    assert e.value.status == N.KJC_LOAD_FAILED and "dimension" in e.value.message
    assert N.lib().kjc_index_dir_info(None, None) == N.KJC_NULL_POINTER


@pytest.mark.parametrize("total,parts", [(18, 1), (18, 4), (7, 8), (50_000_000, 8), (0, 3), (1, 2)])
def test_part_ranges_are_contiguous_and_balanced(total, parts):
    ranges = [api.index_part_range(total, p, parts) for p in range(parts)]
    assert ranges[0][0] == 0 and ranges[-1][1] == total
    for (a, b), (c, d) in zip(ranges, ranges[1:]):
        assert b == c
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1
    from kjarni_b200 import distributed

    assert [(a, b - a) for a, b in ranges] == [distributed.shard_rows(total, parts, p) for p in range(parts)]  # same split as the NCCL path
    with pytest.raises(N.KjarniCudaError):
        api.index_part_range(total, parts, parts)
