"""The C++ tokenizer (kjarni_b200/csrc/tokenizer.hpp; host-only, no GPU: WordPiece, WordLevel and the byte-level
BPE of the RoBERTa family) against goldens produced by
the HuggingFace `tokenizers` crate's own Python binding (tests/golden/make_tokenizer_goldens.py) with the reference's
settings: truncation to max_length (LongestFirst), BatchLongest padding, add_special_tokens = true, texts and pairs."""
import json
import os

import numpy as np
import pytest

from kjarni_b200 import _native as N
from kjarni_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "tokenizer_goldens.json")))


def tok_path(name):
    return os.path.join(HERE, "golden", "tokenizers", name + ".tokenizer.json")


@pytest.mark.parametrize("ci", range(len(GOLD["cases"])))
def test_matches_hf_tokenizers(ci):
    c = GOLD["cases"][ci]
    t = api.Tokenizer(tok_path(c["tokenizer"]), c["max_length"])
    ids, mask, types = t.encode_batch(c["texts"], c["pairs"], add_special_tokens=not c.get("no_special", False))
    want_ids = np.array(c["ids"], np.uint32).reshape(len(c["texts"]), -1)
    assert ids.shape == want_ids.shape, (c["tokenizer"], c["max_length"], ids.shape, want_ids.shape)
    for i, text in enumerate(c["texts"]):
        assert ids[i].tolist() == c["ids"][i], (c["tokenizer"], c["max_length"], text, None if c["pairs"] is None else c["pairs"][i])
        assert types[i].tolist() == c["types"][i], (c["tokenizer"], text)
        assert mask[i].astype(int).tolist() == c["mask"][i], (c["tokenizer"], text)
    t.close()


def test_token_to_id_and_errors(tmp_path):
    t = api.Tokenizer(tok_path("bert_uncased"), 512)
    assert t.token_to_id("[CLS]") == 101 and t.token_to_id("[SEP]") == 102 and t.token_to_id("[PAD]") == 0
    assert t.token_to_id("definitely-not-a-token") is None
    with pytest.raises(N.KjarniCudaError) as e:
        api.Tokenizer(str(tmp_path / "missing.json"))
    assert e.value.status == N.KJC_MODEL_NOT_FOUND
    p = tmp_path / "unigram.json"
    p.write_text(json.dumps({"model": {"type": "Unigram", "vocab": []}}))
    with pytest.raises(N.KjarniCudaError) as e:
        api.Tokenizer(str(p))
    assert e.value.status == N.KJC_INVALID_CONFIG  # WordPiece / WordLevel / byte-level BPE are restated; SentencePiece models are not
    p.write_text(json.dumps({"model": {"type": "BPE", "vocab": {"a": 0, "b": 1}, "merges": ["a b"]}}))
    with pytest.raises(N.KjarniCudaError) as e:  # the merged token "ab" is not in the vocabulary
        api.Tokenizer(str(p))
    assert e.value.status == N.KJC_LOAD_FAILED
    p.write_text("{broken")
    with pytest.raises(N.KjarniCudaError) as e:
        api.Tokenizer(str(p))
    assert e.value.status == N.KJC_LOAD_FAILED
    # invalid UTF-8 is rejected at the boundary like CStr::to_str in the reference's FFI
    import ctypes as C
    bad = (C.c_char_p * 1)(b"\xff\xfe")
    S = C.c_int()
    assert N.lib().kjc_tokenizer_encode_batch(t._h, bad, None, 1, 1, None, None, None, 0, C.byref(S)) == N.KJC_INVALID_UTF8
    t.close()
