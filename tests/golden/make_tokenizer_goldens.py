"""Generates the tokenizer fixtures and goldens in the BUILD container with the HuggingFace `tokenizers` Python binding
(0.22.x: the same Rust crate, tokenizers 0.22.1, that the reference links -- Cargo.toml) configured the way
EncoderLoader::load_from_pretrained does (kjarni-transformers/src/pipeline/encoder/loader.rs:99-115): truncation to
max_length (defaults: LongestFirst, right, stride 0) + BatchLongest padding (pad id 0), encode_batch(..., add_special_tokens=True).

    python tests/golden/make_tokenizer_goldens.py

Writes tests/golden/tokenizers/<name>.tokenizer.json (small synthetic vocabularies) and tests/golden/tokenizer_goldens.json."""
import json
import os

from tokenizers import Tokenizer, decoders, models, normalizers, pre_tokenizers, processors, trainers

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "tokenizers")

WORDS = """the a an of and to in is it that was for on with as by at from be this have or not are but his they her she he you we
all one two three there been their said each which do how if will up other about out many then them these so some would make like
him into time has look more write go see number no way could people my than first water call who oil its now find long down day did
get come made may part over new sound take only little work know place year live me back give most very after thing our just name
good sentence man think say great where help through much before line right too mean old any same tell boy follow came want show
also around form small set put end does another well large must big even such because turn here why ask went men read need land
different home us move try kind hand picture again change off play spell air away animal house point page letter mother answer found
study still learn should america world high every near add food between own below country plant last school father keep tree never
start city earth eye light thought head under story saw left don few while along might close something seem next hard open example
begin life always those both paper together got group often run important until children side feet car mile night walk white sea
began grow took river four carry state once book hear stop without second later miss idea enough eat face watch far indian real
almost let above girl sometimes mountain cut young talk soon list song being leave family it's hello world embedding search index
quick brown fox jumps lazy dog cafe naive resume uber strasse tokyo beijing unbelievable running runner token tokens tokenizer
gpu kernel tensor blackwell query document passage relevant ranking score cosine similarity vector database retrieval""".split()
PIECES = ["##s", "##ed", "##ing", "##er", "##est", "##ly", "##ion", "##able", "##un", "##izer", "##ization", "##ize", "##ness", "##ment",
          "##e", "##d", "##n", "##t", "##y", "##al", "##ic", "##ous", "##ful", "##less", "##1", "##2", "##0", "##9", "##th", "##es"]


def build_vocab():
    vocab = {"[PAD]": 0, "[unused0]": 1}
    nxt = 2
    for t in ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]:
        vocab[t] = 100 + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"].index(t)
    nxt = 104
    singles = list("abcdefghijklmnopqrstuvwxyz0123456789") + list("!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~") + ["é", "ü", "ß", "ñ", "日", "本", "語", "中", "文", "€", "—", "“", "”", "¿", "ı", "σ", "ς", "α"]
    for t in singles + ["##" + c for c in "abcdefghijklmnopqrstuvwxyz0123456789"] + sorted(set(WORDS)) + PIECES + ["Hello", "World", "GPU", "Tokyo", "Café", "##É"]:
        if t not in vocab:
            vocab[t] = nxt
            nxt += 1
    return vocab


def make_tokenizers():
    vocab = build_vocab()
    toks = {}
    t = Tokenizer(models.WordPiece(vocab, unk_token="[UNK]"))
    t.normalizer = normalizers.BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None, lowercase=True)
    t.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    t.post_processor = processors.TemplateProcessing(single="[CLS] $A [SEP]", pair="[CLS] $A [SEP] $B:1 [SEP]:1",
                                                     special_tokens=[("[CLS]", 101), ("[SEP]", 102)])
    t.add_special_tokens(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"])
    toks["bert_uncased"] = t
    t = Tokenizer(models.WordPiece(vocab, unk_token="[UNK]", max_input_chars_per_word=20))
    t.normalizer = normalizers.BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=False, lowercase=False)
    t.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    t.post_processor = processors.BertProcessing(("[SEP]", 102), ("[CLS]", 101))
    t.add_special_tokens(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"])
    toks["bert_cased"] = t
    # the reference's own test fixture shape: WordLevel + Whitespace, no normalizer / post-processor
    # (kjarni-models/src/models/sentence_encoder/tests.rs:16-176)
    wl = {"[PAD]": 0, "[UNK]": 1, "[CLS]": 2, "[SEP]": 3}
    for w in sorted(set(WORDS))[:200] + [",", ".", "!", "?", "'"]:
        wl.setdefault(w, len(wl))
    t = Tokenizer(models.WordLevel(wl, unk_token="[UNK]"))
    t.pre_tokenizer = pre_tokenizers.Whitespace()
    toks["wordlevel"] = t
    t = Tokenizer(models.WordPiece(vocab, unk_token="[UNK]"))
    t.normalizer = normalizers.Sequence([normalizers.NFD(), normalizers.Lowercase(), normalizers.StripAccents()])
    t.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.WhitespaceSplit(), pre_tokenizers.BertPreTokenizer()])
    t.post_processor = processors.TemplateProcessing(single="[CLS] $A [SEP]", pair="[CLS] $A [SEP] $B:1 [SEP]:1",
                                                     special_tokens=[("[CLS]", 101), ("[SEP]", 102)])
    toks["sequence"] = t
    # byte-level BPE in the RoBERTa layout (roberta-base / distilroberta tokenizer.json): ByteLevel pre-tokenizer with the GPT-2
    # pattern, RobertaProcessing (<s> A </s>, <s> A </s></s> B </s>), merges learned here from a small corpus
    t = Tokenizer(models.BPE(unk_token=None))
    t.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    t.decoder = decoders.ByteLevel()
    corpus = [" ".join(WORDS[i:i + 9]) + "." for i in range(0, len(WORDS), 7)] + SINGLES + [a + " " + b for a, b in PAIRS] + \
             ["It's 2019, isn't it? They've said we'll go; I'm sure he'd agree.", "Numbers 123 4567 and symbols #!$ %^& mix\twith\n\nnewlines   and   spaces "]
    trainer = trainers.BpeTrainer(vocab_size=700, min_frequency=1, special_tokens=["<s>", "<pad>", "</s>", "<unk>", "<mask>"],
                                  initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False)
    t.train_from_iterator(corpus, trainer)
    t.post_processor = processors.RobertaProcessing(sep=("</s>", 2), cls=("<s>", 0))
    toks["roberta_bpe"] = t
    return toks


SINGLES = [
    "Hello world", "The quick brown fox jumps over the lazy dog.", "", "   ", "Café naïve résumé ÜBER Straße", "日本語 and 中文 text", "unbelievable tokenization of tokenizers!",
    "it's 2019, isn't it? (yes) [maybe] {no}", "email@example.com costs €9.99 — “quoted”", "tab\tnewline\ncarriage\rreturn\x01soh\x07bell​zero�repl\x7fdel",
    "a" * 101 + " short", "x" * 25 + " y", "[CLS] literal [SEP] inside [MASK] text [UNK]", "İstanbul DİYARBAKIR ΑΣ Σίσυφος", "emoji 😀 mixed 🚀rocket",
    "hello " * 40, "¿Qué tal? ¡Muy bien!", "under_score snake_case __dunder__", "3.14159 1,000,000 10-20 1/2", "Ａ full width Ｂ",
    "한국어 텍스트", "é combining and é precomposed", "word" + " " + "nbsp" + "　" + "ideographic",
]
PAIRS = [
    ("what is the capital", "tokyo is the capital city of the country"), ("query", ""), ("", "document only"), ("hello " * 30, "world " * 30),
    ("short", "document " * 50), ("question " * 50, "short"), ("a b c", "d e f"), ("Café?", "naïve — résumé!"),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    gold = {"tokenizers_version": __import__("tokenizers").__version__, "cases": []}
    for name, tok in make_tokenizers().items():
        path = os.path.join(OUT, name + ".tokenizer.json")
        tok.save(path, pretty=False)
        for max_len in (512, 32, 16, 5):
            t = Tokenizer.from_file(path)
            t.enable_truncation(max_length=max_len)
            t.enable_padding()
            enc = t.encode_batch(SINGLES, add_special_tokens=True)
            gold["cases"].append({"tokenizer": name, "max_length": max_len, "texts": SINGLES, "pairs": None,
                                  "ids": [e.ids for e in enc], "types": [e.type_ids for e in enc], "mask": [e.attention_mask for e in enc]})
            enc = t.encode_batch([(a, b) for a, b in PAIRS], add_special_tokens=True)
            gold["cases"].append({"tokenizer": name, "max_length": max_len, "texts": [a for a, _ in PAIRS], "pairs": [b for _, b in PAIRS],
                                  "ids": [e.ids for e in enc], "types": [e.type_ids for e in enc], "mask": [e.attention_mask for e in enc]})
        # each text alone (no batch padding), without special tokens
        t = Tokenizer.from_file(path)
        for s in SINGLES[:8]:
            e = t.encode(s, add_special_tokens=False)
            gold["cases"].append({"tokenizer": name, "max_length": 0, "texts": [s], "pairs": None, "no_special": True,
                                  "ids": [e.ids], "types": [e.type_ids], "mask": [e.attention_mask]})
    with open(os.path.join(HERE, "tokenizer_goldens.json"), "w") as f:
        json.dump(gold, f, ensure_ascii=True, separators=(",", ":"))
    print(len(gold["cases"]), "cases", os.path.getsize(os.path.join(HERE, "tokenizer_goldens.json")), "bytes")


if __name__ == "__main__":
    main()
