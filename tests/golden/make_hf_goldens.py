"""Generates tests/golden/hf_goldens.npz in the BUILD container (needs `transformers`
+ torch CPU; not needed at test time).  Independent cross-check of the oracle:
the random-init model directories written by kjarni_b200.synth are loaded into
HuggingFace BertModel / BertForSequenceClassification /
DistilBertForSequenceClassification (whose state_dict key names are exactly the
tensor names Kjarni's layouts expect) and run on the synthetic token ids.

    python tests/golden/make_hf_goldens.py
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kjarni_b200 import synth  # noqa: E402


def load_hf(arch, d):
    from transformers import BertConfig, BertForSequenceClassification, BertModel
    from transformers import DistilBertConfig, DistilBertForSequenceClassification

    t, cfg = synth.make_weights(arch)
    family = synth.ARCHS[arch][0]
    sd = {k: torch.from_numpy(v.copy()) for k, v in t.items()}
    if family == "roberta":
        from transformers import RobertaConfig, RobertaForSequenceClassification

        c = RobertaConfig(vocab_size=cfg["vocab_size"], hidden_size=cfg["hidden_size"], num_hidden_layers=cfg["num_hidden_layers"],
                          num_attention_heads=cfg["num_attention_heads"], intermediate_size=cfg["intermediate_size"],
                          hidden_act="gelu", layer_norm_eps=cfg["layer_norm_eps"], type_vocab_size=cfg["type_vocab_size"],
                          max_position_embeddings=cfg["max_position_embeddings"], hidden_dropout_prob=0.0,
                          attention_probs_dropout_prob=0.0, classifier_dropout=0.0, num_labels=len(cfg["id2label"]), pad_token_id=1)
        m = RobertaForSequenceClassification(c)
    elif family == "distilbert":
        c = DistilBertConfig(vocab_size=cfg["vocab_size"], dim=cfg["dim"], hidden_dim=cfg["hidden_dim"],
                             n_layers=cfg["n_layers"], n_heads=cfg["n_heads"], activation="gelu",
                             max_position_embeddings=cfg["max_position_embeddings"], num_labels=2,
                             dropout=0.0, attention_dropout=0.0, seq_classif_dropout=0.0)
        m = DistilBertForSequenceClassification(c)
    else:
        c = BertConfig(vocab_size=cfg["vocab_size"], hidden_size=cfg["hidden_size"],
                       num_hidden_layers=cfg["num_hidden_layers"], num_attention_heads=cfg["num_attention_heads"],
                       intermediate_size=cfg["intermediate_size"], hidden_act="gelu", layer_norm_eps=1e-12,
                       type_vocab_size=cfg["type_vocab_size"], max_position_embeddings=cfg["max_position_embeddings"],
                       hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                       num_labels=cfg.get("num_labels", 2))
        m = BertForSequenceClassification(c) if family == "bert_prefixed" else BertModel(c, add_pooling_layer=False)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    missing = [k for k in missing if "position_ids" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    return m.eval()


def main():
    out = {}
    cases = [("tiny-bert", 6, 16), ("tiny-cross-encoder", 6, 16), ("tiny-distilbert", 6, 16), ("minilm-l6", 4, 32), ("tiny-roberta", 6, 16)]
    for arch, B, S in cases:
        family, H, L, heads, I, vocab = synth.ARCHS[arch][:6]
        ids, mask, types = synth.synth_tokens(B, S, vocab, regime="P", seed=7, pair=(family == "bert_prefixed"))
        m = load_hf(arch, None)
        with torch.no_grad():
            ti = torch.from_numpy(ids.astype(np.int64))
            tm = torch.from_numpy(mask.astype(np.int64))
            if family == "bert":
                # Embedder path: no token-type ids -> row 0 everywhere == HF default zeros
                h = m(input_ids=ti, attention_mask=tm).last_hidden_state.numpy()
                mf = mask.astype(np.float32)
                e = (h * mf[:, :, None]).sum(1) / np.maximum(mf.sum(1, keepdims=True), 1)
                e = e / np.linalg.norm(e, axis=1, keepdims=True)
                out[arch + "/hidden_valid"] = (h * mf[:, :, None]).astype(np.float32)
                out[arch + "/embedding"] = e.astype(np.float32)
            elif family == "roberta":
                # Kjarni adds position rows offset+s = 2+s to token s (extra_pos_embeddings 2,
                # KM/models/sequence_classifier/configs.rs:223 -> KT/cpu/embeddings/mod.rs:199-214); HF derives the same rows
                # from padding_idx for left-aligned text, given explicitly here so pad ids play no role
                pos = torch.arange(2, 2 + S).unsqueeze(0).expand(B, S)
                lg = m(input_ids=ti, attention_mask=tm, position_ids=pos, token_type_ids=torch.zeros_like(ti)).logits
                out[arch + "/logits"] = lg.numpy().astype(np.float32)
            elif family == "bert_prefixed":
                lg = m(input_ids=ti, attention_mask=tm, token_type_ids=torch.from_numpy(types.astype(np.int64))).logits
                out[arch + "/logits"] = lg.numpy().astype(np.float32)
            else:
                out[arch + "/logits"] = m(input_ids=ti, attention_mask=tm).logits.numpy().astype(np.float32)
        out[arch + "/shape"] = np.array([B, S])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hf_goldens.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
