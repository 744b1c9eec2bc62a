"""Oracle (loaded through the model-directory path) vs committed HuggingFace
outputs (tests/golden/hf_goldens.npz, produced by tests/golden/make_hf_goldens.py)."""
import os

import numpy as np
import pytest

from kjarni_b200 import synth
from oracle import kjarni_oracle as ko

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hf_goldens.npz"))


@pytest.fixture(scope="module")
def model_dirs(tmp_path_factory):
    root = tmp_path_factory.mktemp("models")
    return {a: synth.write_model_dir(str(root / a), a) for a in
            ("tiny-bert", "tiny-cross-encoder", "tiny-distilbert", "minilm-l6")}


@pytest.mark.parametrize("arch", ["tiny-bert", "minilm-l6"])
def test_embedding_path_matches_hf(model_dirs, arch):
    B, S = G[arch + "/shape"]
    vocab = synth.ARCHS[arch][5]
    ids, mask, _ = synth.synth_tokens(int(B), int(S), vocab, regime="P", seed=7)
    m = ko.load_model_dir(model_dirs[arch])
    assert m.arch == "bert" and m.head_kind is None
    for noalloc in (False, True):
        h = ko.encoder_forward(m, ids, mask, None, noalloc=noalloc)
        hv = h * mask.astype(np.float32)[:, :, None]
        assert np.abs(hv - G[arch + "/hidden_valid"]).max() < 2e-5
    e = ko.embed(m, ids, mask)
    assert np.abs(e - G[arch + "/embedding"]).max() < 2e-6
    assert np.abs(np.linalg.norm(e, axis=1) - 1).max() < 1e-6


def test_cross_encoder_matches_hf(model_dirs):
    arch = "tiny-cross-encoder"
    B, S = G[arch + "/shape"]
    ids, mask, types = synth.synth_tokens(int(B), int(S), synth.ARCHS[arch][5], regime="P", seed=7, pair=True)
    m = ko.load_model_dir(model_dirs[arch])
    assert m.arch == "bert_prefixed" and m.head_kind == "pooler_tanh"
    lg = ko.predict_logits(m, ids, mask, types)
    assert lg.shape == (B, 1)
    assert np.abs(lg - G[arch + "/logits"]).max() < 2e-5


def test_distilbert_classifier_matches_hf(model_dirs):
    arch = "tiny-distilbert"
    B, S = G[arch + "/shape"]
    ids, mask, _ = synth.synth_tokens(int(B), int(S), synth.ARCHS[arch][5], regime="P", seed=7)
    m = ko.load_model_dir(model_dirs[arch])
    assert m.arch == "distilbert" and m.head_kind == "pre_relu" and m.typ is None
    assert m.labels == ["NEGATIVE", "POSITIVE"]
    lg = ko.predict_logits(m, ids, mask)
    assert np.abs(lg - G[arch + "/logits"]).max() < 2e-5
    p = ko.classify_probs(lg)
    assert np.allclose(p.sum(1), 1, atol=1e-6)
