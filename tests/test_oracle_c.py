"""The C restatement (oracle/kjarni_oracle.c, the timed CPU baseline) against the pinned numpy oracle."""
import numpy as np
import pytest

from kjarni_b200 import synth
from oracle import kjarni_oracle as ko

koc = pytest.importorskip("oracle.kjarni_oracle_c")


@pytest.mark.parametrize("arch,b,s", [("tiny-bert", 5, 16), ("tiny-distilbert", 3, 9), ("tiny-bert", 70, 16)])
def test_c_encoder_matches_numpy_oracle(tmp_path, arch, b, s):
    d = synth.write_model_dir(str(tmp_path / arch), arch)
    m = ko.load_model_dir(d)
    cm = koc.CModel(m)
    ids, mask, _ = synth.synth_tokens(b, s, synth.ARCHS[arch][5], regime="P", seed=3)
    maskf = mask.astype(np.float32)
    for noalloc in (False, True):
        want = ko.encoder_forward(m, ids, maskf, None, noalloc=noalloc)
        got = cm.hidden_states(ids, maskf, None, noalloc=noalloc)
        assert np.abs(got - want).max() < 2e-5
    want = ko.embed(m, ids, mask)
    got = cm.embed(ids, maskf)
    assert np.abs(got - want).max() < 1e-5


def test_c_scan_matches_numpy_oracle():
    rows = ko.synth_rows(7, 0, 3000, 384)
    rows[17] = 0.0
    rows[40] = rows[12]  # exact duplicate: the lower id wins the tie
    q = ko.synth_rows(11, 0, 5, 384)
    q[2] = rows[12] * 2.0
    ids, sc = koc.scan_topk(rows, q, 10)
    wi, ws = ko.batched_topk(rows, q, 10)
    assert np.abs(sc - ws).max() < 2e-6
    for qi in range(5):
        for j in range(10):
            assert ids[qi, j] == np.uint64(wi[qi, j]) or abs(float(sc[qi, j]) - float(ws[qi, j])) <= 1e-6
    assert ids[2, 0] == 12 and ids[2, 1] == 40
