// Text -> token ids for the BERT-family checkpoints on the hot path.
//
// The reference tokenises with the HuggingFace `tokenizers` crate (0.22.1, Cargo.toml) configured in
// EncoderLoader::load_from_pretrained (kjarni-transformers/src/pipeline/encoder/loader.rs:99-115): `tokenizer.json` from the
// model directory, truncation to max_position_embeddings (strategy LongestFirst, right, stride 0) and BatchLongest padding
// (pad id 0, type id 0); texts go through `encode_batch(texts, add_special_tokens = true)`
// (kjarni-transformers/src/cpu/encoder/traits.rs:140-145), query/document pairs through the same call with tuples
// (kjarni-models/src/models/cross_encoder/model.rs:176-194).  That crate is a third-party dependency absent from the
// reference tree; this file restates the published algorithm of the components BERT-family tokenizer.json files use:
//   normalizer      BertNormalizer | Lowercase | NFD | StripAccents | Sequence | null
//   pre_tokenizer   BertPreTokenizer | Whitespace | WhitespaceSplit | Sequence | null
//   model           WordPiece | WordLevel
//   post_processor  BertProcessing | TemplateProcessing | null
//   added_tokens    matched verbatim in the raw text (special tokens such as [CLS] / [SEP] / [MASK])
// Parity is pinned by tests/golden/tokenizer_goldens.json, produced here by the crate's own Python binding
// (tokenizers 0.22.2; tests/golden/make_tokenizer_goldens.py).  Byte-level BPE (RoBERTa) and Unigram are not restated: the loader
// reports KJC_INVALID_CONFIG for them and such models are driven with token ids.
#pragma once
#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "host_util.hpp"

namespace kj {

namespace uni {
#include "unicode_tables.inc"

inline bool in_ranges(const uint32_t (*r)[2], int n, uint32_t c) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        if (c < r[mid][0]) hi = mid - 1;
        else if (c > r[mid][1]) lo = mid + 1;
        else return true;
    }
    return false;
}
inline bool is_white(uint32_t c) { return in_ranges(kUniWhite, kUniWhite_n, c); }
inline bool is_other(uint32_t c) { return in_ranges(kUniOther, kUniOther_n, c); }
inline bool is_mn(uint32_t c) { return in_ranges(kUniMn, kUniMn_n, c); }
inline bool is_word(uint32_t c) { return in_ranges(kUniWord, kUniWord_n, c); }
inline bool is_alnum(uint32_t c) { return in_ranges(kUniAlnum, kUniAlnum_n, c); }  // ~ Rust char::is_alphanumeric (see gen_unicode_tables.py)
inline bool is_letter(uint32_t c) { return in_ranges(kUniLetter, kUniLetter_n, c); }  // regex \p{L}
inline bool is_number(uint32_t c) { return in_ranges(kUniNumber, kUniNumber_n, c); }  // regex \p{N}
inline bool is_ascii_punct(uint32_t c) { return (c >= 33 && c <= 47) || (c >= 58 && c <= 64) || (c >= 91 && c <= 96) || (c >= 123 && c <= 126); }
inline bool is_punct(uint32_t c) { return is_ascii_punct(c) || in_ranges(kUniPunct, kUniPunct_n, c); }
// BertNormalizer::is_chinese_char ranges
inline bool is_cjk(uint32_t c) {
    return (c >= 0x4E00 && c <= 0x9FFF) || (c >= 0x3400 && c <= 0x4DBF) || (c >= 0x20000 && c <= 0x2A6DF) || (c >= 0x2A700 && c <= 0x2B73F) ||
           (c >= 0x2B740 && c <= 0x2B81F) || (c >= 0x2B920 && c <= 0x2CEAF) || (c >= 0xF900 && c <= 0xFAFF) || (c >= 0x2F800 && c <= 0x2FA1F);
}
template <int W>
inline const uint32_t* find_row(const uint32_t (*t)[W], int n, uint32_t c) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        if (c < t[mid][0]) hi = mid - 1;
        else if (c > t[mid][0]) lo = mid + 1;
        else return t[mid];
    }
    return nullptr;
}
inline void nfd_append(uint32_t c, std::vector<uint32_t>& out) {
    if (c >= 0xAC00 && c <= 0xD7A3) {  // Hangul syllable: algorithmic decomposition
        const uint32_t s = c - 0xAC00, l = 0x1100 + s / 588, v = 0x1161 + (s % 588) / 28, t = 0x11A7 + s % 28;
        out.push_back(l);
        out.push_back(v);
        if (t != 0x11A7) out.push_back(t);
        return;
    }
    if (c < 0xC0) { out.push_back(c); return; }
    if (const uint32_t* r = find_row<6>(kUniNfd, kUniNfd_n, c)) {
        for (uint32_t i = 0; i < r[1]; ++i) out.push_back(r[2 + i]);
    } else {
        out.push_back(c);
    }
}
inline void lower_append(uint32_t c, std::vector<uint32_t>& out) {
    if (c < 0x80) { out.push_back((c >= 'A' && c <= 'Z') ? c + 32 : c); return; }
    if (const uint32_t* r = find_row<5>(kUniLower, kUniLower_n, c)) {
        for (uint32_t i = 0; i < r[1]; ++i) out.push_back(r[2 + i]);
    } else {
        out.push_back(c);
    }
}
// UTF-8 <-> code points (invalid bytes decode to U+FFFD; the C ABI validates UTF-8 before it gets here)
inline std::vector<uint32_t> decode(const std::string& s) {
    std::vector<uint32_t> out;
    out.reserve(s.size());
    const unsigned char* p = reinterpret_cast<const unsigned char*>(s.data());
    const size_t n = s.size();
    for (size_t i = 0; i < n;) {
        const unsigned char b = p[i];
        uint32_t c;
        int len;
        if (b < 0x80) { c = b; len = 1; }
        else if ((b >> 5) == 6) { c = b & 0x1F; len = 2; }
        else if ((b >> 4) == 14) { c = b & 0x0F; len = 3; }
        else if ((b >> 3) == 30) { c = b & 0x07; len = 4; }
        else { out.push_back(0xFFFD); ++i; continue; }
        if (i + len > n) { out.push_back(0xFFFD); break; }
        bool ok = true;
        for (int k = 1; k < len; ++k) {
            if ((p[i + k] >> 6) != 2) { ok = false; break; }
            c = (c << 6) | (p[i + k] & 0x3F);
        }
        if (!ok) { out.push_back(0xFFFD); ++i; continue; }
        out.push_back(c);
        i += len;
    }
    return out;
}
inline void encode_append(uint32_t c, std::string& s) {
    if (c < 0x80) s.push_back(static_cast<char>(c));
    else if (c < 0x800) { s.push_back(static_cast<char>(0xC0 | (c >> 6))); s.push_back(static_cast<char>(0x80 | (c & 0x3F))); }
    else if (c < 0x10000) {
        s.push_back(static_cast<char>(0xE0 | (c >> 12)));
        s.push_back(static_cast<char>(0x80 | ((c >> 6) & 0x3F)));
        s.push_back(static_cast<char>(0x80 | (c & 0x3F)));
    } else {
        s.push_back(static_cast<char>(0xF0 | (c >> 18)));
        s.push_back(static_cast<char>(0x80 | ((c >> 12) & 0x3F)));
        s.push_back(static_cast<char>(0x80 | ((c >> 6) & 0x3F)));
        s.push_back(static_cast<char>(0x80 | (c & 0x3F)));
    }
}
inline bool valid_utf8(const char* s) {
    const unsigned char* p = reinterpret_cast<const unsigned char*>(s);
    while (*p) {
        int len;
        uint32_t c;
        if (*p < 0x80) { ++p; continue; }
        else if ((*p >> 5) == 6) { len = 2; c = *p & 0x1F; }
        else if ((*p >> 4) == 14) { len = 3; c = *p & 0x0F; }
        else if ((*p >> 3) == 30) { len = 4; c = *p & 0x07; }
        else return false;
        for (int k = 1; k < len; ++k) {
            if ((p[k] >> 6) != 2) return false;
            c = (c << 6) | (p[k] & 0x3F);
        }
        if ((len == 2 && c < 0x80) || (len == 3 && c < 0x800) || (len == 4 && (c < 0x10000 || c > 0x10FFFF)) || (c >= 0xD800 && c <= 0xDFFF)) return false;
        p += len;
    }
    return true;
}
}  // namespace uni

class Tokenizer {
  public:
    struct Encoded {
        std::vector<uint32_t> ids, types;
    };

    // max_length <= 0: no truncation
    Tokenizer(const std::string& json_path, int max_length) : max_length_(max_length) {
        const std::string text = read_text_file(json_path, KJC_MODEL_NOT_FOUND);
        Json j;
        try {
            j = JsonParser(text.data(), text.size()).parse();
        } catch (const Error& e) {
            throw Error(KJC_LOAD_FAILED, "Failed to load tokenizer: " + std::string(e.what()));
        }
        const Json* model = j.get("model");
        if (!model || model->type != Json::Obj) throw Error(KJC_LOAD_FAILED, "Failed to load tokenizer: missing `model`");
        std::string mtype = model->string("type", "");
        const Json* vocab = model->get("vocab");
        if (mtype.empty()) mtype = model->has("continuing_subword_prefix") ? "WordPiece" : "WordLevel";  // older files omit the tag
        if (mtype == "WordPiece") {
            wordpiece_ = true;
            prefix_ = model->string("continuing_subword_prefix", "##");
            max_chars_ = static_cast<int>(model->number("max_input_chars_per_word", 100));
        } else if (mtype == "WordLevel") {
            wordpiece_ = false;
        } else if (mtype == "BPE") {  // byte-level BPE of the RoBERTa family (tokenizers crate: models/bpe/model.rs, word.rs)
            wordpiece_ = false;
            bpe_ = true;
            for (const char* k : {"dropout", "continuing_subword_prefix", "end_of_word_suffix"})
                if (const Json* v = model->get(k))
                    if (v->type != Json::Null) throw Error(KJC_INVALID_CONFIG, std::string("BPE option '") + k + "' is not supported");
            if (const Json* v = model->get("byte_fallback"))
                if (v->type == Json::Bool && v->b) throw Error(KJC_INVALID_CONFIG, "BPE byte_fallback is not supported");
        } else {
            throw Error(KJC_INVALID_CONFIG, "tokenizer model '" + mtype + "' is not supported by the CUDA backend's tokenizer (WordPiece / WordLevel / BPE): "
                                            "pass token ids instead");
        }
        if (!vocab || vocab->type != Json::Obj) throw Error(KJC_LOAD_FAILED, "Failed to load tokenizer: model.vocab is not an object");
        vocab_.reserve(vocab->obj.size() * 2);
        for (auto& kv : vocab->obj)
            if (kv.second.type == Json::Num) vocab_[kv.first] = static_cast<uint32_t>(kv.second.num);
        unk_ = model->string("unk_token", bpe_ ? "" : "[UNK]");
        has_unk_ = !unk_.empty() && lookup(unk_, unk_id_);
        if (bpe_) {
            // merges: ["a b", ...] (older files) or [["a", "b"], ...]; rank = position; the merged token must be in the vocabulary
            const Json* mg = model->get("merges");
            if (!mg || mg->type != Json::Arr) throw Error(KJC_LOAD_FAILED, "Failed to load tokenizer: model.merges is not an array");
            uint32_t rank = 0;
            for (const Json& m : mg->arr) {
                std::string a, b;
                if (m.type == Json::Str) {
                    const size_t sp = m.str.find(' ');
                    if (sp == std::string::npos) throw Error(KJC_LOAD_FAILED, "Failed to load tokenizer: bad merge entry");
                    a = m.str.substr(0, sp);
                    b = m.str.substr(sp + 1);
                } else if (m.type == Json::Arr && m.arr.size() == 2 && m.arr[0].type == Json::Str && m.arr[1].type == Json::Str) {
                    a = m.arr[0].str;
                    b = m.arr[1].str;
                } else {
                    throw Error(KJC_LOAD_FAILED, "Failed to load tokenizer: bad merge entry");
                }
                uint32_t ia, ib, iab;
                if (!lookup(a, ia) || !lookup(b, ib) || !lookup(a + b, iab)) throw Error(KJC_LOAD_FAILED, "Failed to load tokenizer: merge token out of vocabulary");
                merges_.emplace((static_cast<uint64_t>(ia) << 32) | ib, std::make_pair(rank, iab));
                ++rank;
            }
            // bytes_to_unicode of the ByteLevel pre-tokenizer: printable bytes map to themselves, the rest to U+0100...
            int n = 0;
            for (int b = 0; b < 256; ++b) {
                const bool keep = (b >= 33 && b <= 126) || (b >= 161 && b <= 172) || (b >= 174 && b <= 255);
                byte_char_[b] = keep ? static_cast<uint32_t>(b) : static_cast<uint32_t>(256 + n++);
            }
        }
        if (const Json* n = j.get("normalizer")) parse_normalizer(*n);
        if (const Json* p = j.get("pre_tokenizer")) parse_pre(*p);
        if (const Json* at = j.get("added_tokens"))
            if (at->type == Json::Arr)
                for (const Json& t : at->arr) {
                    const std::string c = t.string("content", "");
                    if (c.empty()) continue;
                    const uint32_t id = static_cast<uint32_t>(t.number("id", 0));
                    added_.push_back({c, id});
                    vocab_.emplace(c, id);  // token_to_id also sees added tokens
                }
        std::sort(added_.begin(), added_.end(), [](const Added& a, const Added& b) { return a.content.size() > b.content.size(); });
        if (const Json* pp = j.get("post_processor")) parse_post(*pp);
        if (single_.empty()) single_.push_back({false, 0, 0, 0});                                     // $A
        if (pair_.empty()) { pair_.push_back({false, 0, 0, 0}); pair_.push_back({false, 1, 1, 0}); }  // $A $B:1
    }

    bool token_to_id(const std::string& tok, uint32_t& id) const { return lookup(tok, id); }
    int max_length() const { return max_length_; }

    // One text (b == nullptr) or a pair: tokenise, truncate, add the post-processor's special tokens.
    Encoded encode(const std::string& a, const std::string* b, bool add_special) const {
        Encoded ea, eb;
        tokenize_into(a, ea.ids);
        if (b) tokenize_into(*b, eb.ids);
        const std::vector<Piece>& tpl = b ? pair_ : single_;
        int n_added = 0;
        if (add_special)
            for (const Piece& p : tpl) n_added += p.special ? 1 : 0;
        if (max_length_ > 0) {
            // TruncationParams {max_length - n_added_tokens}, strategy LongestFirst, direction Right, stride 0
            const size_t budget = static_cast<size_t>(std::max(0, max_length_ - n_added));
            size_t n1 = ea.ids.size(), n2 = eb.ids.size();
            if (!b) {
                if (n1 > budget) ea.ids.resize(budget);
            } else if (n1 + n2 > budget) {
                bool swap = false;
                if (n1 > n2) { swap = true; std::swap(n1, n2); }
                if (n1 > budget) n2 = n1;
                else n2 = std::max(n1, budget - n1);
                if (n1 + n2 > budget) { n1 = budget / 2; n2 = n1 + budget % 2; }
                if (swap) std::swap(n1, n2);
                if (ea.ids.size() > n1) ea.ids.resize(n1);
                if (eb.ids.size() > n2) eb.ids.resize(n2);
            }
        }
        Encoded out;
        for (const Piece& p : tpl) {
            if (p.special) {
                if (!add_special) continue;
                out.ids.push_back(p.token_id);
                out.types.push_back(p.type_id);
            } else {
                const std::vector<uint32_t>& src = p.seq == 0 ? ea.ids : eb.ids;
                out.ids.insert(out.ids.end(), src.begin(), src.end());
                out.types.insert(out.types.end(), src.size(), p.type_id);
            }
        }
        return out;
    }

    // encode_batch(texts, add_special_tokens) + BatchLongest padding (pad id 0, type id 0, mask 0).  `b` empty or one per text.
    void encode_batch(const std::vector<std::string>& a, const std::vector<std::string>& b, bool add_special, std::vector<uint32_t>& ids,
                      std::vector<float>& mask, std::vector<uint32_t>& types, int& seq_len) const {
        std::vector<Encoded> enc(a.size());
        size_t S = 0;
        for (size_t i = 0; i < a.size(); ++i) {
            enc[i] = encode(a[i], b.empty() ? nullptr : &b[i], add_special);
            S = std::max(S, enc[i].ids.size());
        }
        seq_len = static_cast<int>(S);
        ids.assign(a.size() * S, 0u);
        types.assign(a.size() * S, 0u);
        mask.assign(a.size() * S, 0.0f);
        for (size_t i = 0; i < a.size(); ++i)
            for (size_t k = 0; k < enc[i].ids.size(); ++k) {
                ids[i * S + k] = enc[i].ids[k];
                types[i * S + k] = enc[i].types[k];
                mask[i * S + k] = 1.0f;
            }
    }

  private:
    enum NormOp { N_CLEAN, N_CJK, N_NFD, N_STRIP, N_LOWER };
    enum PreOp { P_BERT, P_WHITESPACE, P_SPLIT, P_BYTELEVEL, P_BYTELEVEL_NOSPLIT };
    struct Added { std::string content; uint32_t id; };
    struct Piece { bool special; int seq; uint32_t type_id; uint32_t token_id; };

    bool lookup(const std::string& t, uint32_t& id) const {
        auto it = vocab_.find(t);
        if (it == vocab_.end()) return false;
        id = it->second;
        return true;
    }

    void parse_normalizer(const Json& n) {
        if (n.type != Json::Obj) return;
        const std::string t = n.string("type", "");
        if (t == "BertNormalizer") {
            auto flag = [&](const char* k, bool d) { const Json* v = n.get(k); return (v && v->type == Json::Bool) ? v->b : d; };
            const bool lower = flag("lowercase", true);
            const Json* sa = n.get("strip_accents");
            const bool strip = (sa && sa->type == Json::Bool) ? sa->b : lower;  // strip_accents.unwrap_or(lowercase)
            if (flag("clean_text", true)) norm_.push_back(N_CLEAN);
            if (flag("handle_chinese_chars", true)) norm_.push_back(N_CJK);
            if (strip) { norm_.push_back(N_NFD); norm_.push_back(N_STRIP); }
            if (lower) norm_.push_back(N_LOWER);
        } else if (t == "Lowercase") norm_.push_back(N_LOWER);
        else if (t == "NFD") norm_.push_back(N_NFD);
        else if (t == "StripAccents") norm_.push_back(N_STRIP);
        else if (t == "Sequence") {
            if (const Json* l = n.get("normalizers"))
                for (const Json& x : l->arr) parse_normalizer(x);
        } else throw Error(KJC_INVALID_CONFIG, "tokenizer normalizer '" + t + "' is not supported");
    }
    void parse_pre(const Json& p) {
        if (p.type != Json::Obj) return;
        const std::string t = p.string("type", "");
        if (t == "BertPreTokenizer") pre_.push_back(P_BERT);
        else if (t == "Whitespace") pre_.push_back(P_WHITESPACE);
        else if (t == "WhitespaceSplit") pre_.push_back(P_SPLIT);
        else if (t == "ByteLevel") {
            auto flag = [&](const char* k, bool d) { const Json* v = p.get(k); return (v && v->type == Json::Bool) ? v->b : d; };
            byte_prefix_space_ = flag("add_prefix_space", true);
            pre_.push_back(flag("use_regex", true) ? P_BYTELEVEL : P_BYTELEVEL_NOSPLIT);
        }
        else if (t == "Sequence") {
            if (const Json* l = p.get("pretokenizers"))
                for (const Json& x : l->arr) parse_pre(x);
        } else throw Error(KJC_INVALID_CONFIG, "tokenizer pre_tokenizer '" + t + "' is not supported");
    }
    void parse_post(const Json& pp) {
        if (pp.type != Json::Obj) return;
        const std::string t = pp.string("type", "");
        auto tok_pair = [&](const char* key, uint32_t& id) {
            const Json* v = pp.get(key);
            if (!v || v->type != Json::Arr || v->arr.size() != 2) throw Error(KJC_LOAD_FAILED, "Failed to load tokenizer: bad BertProcessing");
            id = static_cast<uint32_t>(v->arr[1].num);
        };
        if (t == "BertProcessing") {
            uint32_t cls = 0, sep = 0;
            tok_pair("cls", cls);
            tok_pair("sep", sep);
            single_ = {{true, 0, 0, cls}, {false, 0, 0, 0}, {true, 0, 0, sep}};
            pair_ = {{true, 0, 0, cls}, {false, 0, 0, 0}, {true, 0, 0, sep}, {false, 1, 1, 0}, {true, 0, 1, sep}};
        } else if (t == "RobertaProcessing") {  // <s> A </s>   and   <s> A </s> </s> B </s>, every type id 0 (processors/roberta.rs)
            uint32_t cls = 0, sep = 0;
            tok_pair("cls", cls);
            tok_pair("sep", sep);
            single_ = {{true, 0, 0, cls}, {false, 0, 0, 0}, {true, 0, 0, sep}};
            pair_ = {{true, 0, 0, cls}, {false, 0, 0, 0}, {true, 0, 0, sep}, {true, 0, 0, sep}, {false, 1, 0, 0}, {true, 0, 0, sep}};
        } else if (t == "TemplateProcessing") {
            const Json* st = pp.get("special_tokens");
            auto parse_tpl = [&](const char* key, std::vector<Piece>& out) {
                const Json* arr = pp.get(key);
                if (!arr || arr->type != Json::Arr) return;
                for (const Json& item : arr->arr) {
                    if (const Json* s = item.get("SpecialToken")) {
                        const std::string name = s->string("id", "");
                        uint32_t id = 0;
                        bool found = false;
                        if (st)
                            if (const Json* e = st->get(name))
                                if (const Json* idsj = e->get("ids"))
                                    if (idsj->type == Json::Arr && !idsj->arr.empty()) { id = static_cast<uint32_t>(idsj->arr[0].num); found = true; }
                        if (!found && !lookup(name, id)) throw Error(KJC_LOAD_FAILED, "Failed to load tokenizer: special token '" + name + "' has no id");
                        out.push_back({true, 0, static_cast<uint32_t>(s->number("type_id", 0)), id});
                    } else if (const Json* q = item.get("Sequence")) {
                        out.push_back({false, q->string("id", "A") == "B" ? 1 : 0, static_cast<uint32_t>(q->number("type_id", 0)), 0});
                    }
                }
            };
            parse_tpl("single", single_);
            parse_tpl("pair", pair_);
        } else if (t == "Sequence") {
            if (const Json* l = pp.get("processors"))
                for (const Json& x : l->arr) parse_post(x);
        } else if (t == "ByteLevel") {
            // no special tokens
        } else {
            throw Error(KJC_INVALID_CONFIG, "tokenizer post_processor '" + t + "' is not supported");
        }
    }

    std::vector<uint32_t> normalize(const std::vector<uint32_t>& in) const {
        std::vector<uint32_t> cur = in, nxt;
        for (NormOp op : norm_) {
            nxt.clear();
            nxt.reserve(cur.size() + 8);
            switch (op) {
                case N_CLEAN:
                    for (uint32_t c : cur) {
                        const bool ws = c == '\t' || c == '\n' || c == '\r' || uni::is_white(c);
                        const bool ctrl = !(c == '\t' || c == '\n' || c == '\r') && uni::is_other(c);
                        if (c == 0 || c == 0xFFFD || ctrl) continue;
                        nxt.push_back(ws ? ' ' : c);
                    }
                    break;
                case N_CJK:
                    for (uint32_t c : cur) {
                        if (uni::is_cjk(c)) { nxt.push_back(' '); nxt.push_back(c); nxt.push_back(' '); }
                        else nxt.push_back(c);
                    }
                    break;
                case N_NFD:
                    for (uint32_t c : cur) uni::nfd_append(c, nxt);
                    break;
                case N_STRIP:
                    for (uint32_t c : cur)
                        if (!uni::is_mn(c)) nxt.push_back(c);
                    break;
                case N_LOWER:
                    for (uint32_t c : cur) uni::lower_append(c, nxt);
                    break;
            }
            cur.swap(nxt);
        }
        return cur;
    }

    // GPT-2 / RoBERTa pattern of the ByteLevel pre-tokenizer (pre_tokenizers/byte_level.rs):
    //   's|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+        (first matching alternative, greedy)
    static size_t gpt2_match(const std::vector<uint32_t>& w, size_t i) {
        const size_t n = w.size();
        if (w[i] == '\'' && i + 1 < n) {
            const uint32_t a = w[i + 1], b = i + 2 < n ? w[i + 2] : 0;
            if (a == 's' || a == 't') return i + 2;
            if ((a == 'r' && b == 'e') || (a == 'v' && b == 'e')) return i + 3;
            if (a == 'm') return i + 2;
            if (a == 'l' && b == 'l') return i + 3;
            if (a == 'd') return i + 2;
        }
        auto cls = [&](size_t j) { return uni::is_letter(w[j]) ? 1 : (uni::is_number(w[j]) ? 2 : (uni::is_white(w[j]) ? 0 : 3)); };
        const size_t j = (w[i] == ' ' && i + 1 < n) ? i + 1 : i;  // " ?"
        const int c = cls(j);
        if (c != 0) {  // alternatives 2-4: an optional space, then a run of one class
            size_t e = j;
            while (e < n && cls(e) == c) ++e;
            return e;
        }
        size_t e = i;  // whitespace run
        while (e < n && uni::is_white(w[e])) ++e;
        if (e < n && e - i > 1) --e;  // \s+(?!\S): leave the last blank to the next token unless the run ends the text
        return e;
    }

    // splits [words] further according to one pre-tokenizer
    void pre_split(PreOp op, const std::vector<std::vector<uint32_t>>& in, std::vector<std::vector<uint32_t>>& out) const {
        out.clear();
        for (const auto& w : in) {
            std::vector<uint32_t> cur;
            auto flush = [&] { if (!cur.empty()) { out.push_back(cur); cur.clear(); } };
            if (op == P_BYTELEVEL || op == P_BYTELEVEL_NOSPLIT) {
                std::vector<uint32_t> t;
                if (byte_prefix_space_ && !w.empty() && w[0] != ' ') t.push_back(' ');
                t.insert(t.end(), w.begin(), w.end());
                auto emit = [&](size_t a, size_t b) {  // UTF-8 bytes of the piece, each mapped to its stand-in character
                    std::string u8;
                    for (size_t i = a; i < b; ++i) uni::encode_append(t[i], u8);
                    std::vector<uint32_t> piece;
                    for (unsigned char ch : u8) piece.push_back(byte_char_[ch]);
                    if (!piece.empty()) out.push_back(std::move(piece));
                };
                if (op == P_BYTELEVEL_NOSPLIT) emit(0, t.size());
                else
                    for (size_t i = 0; i < t.size();) {
                        const size_t e = gpt2_match(t, i);
                        emit(i, e);
                        i = e;
                    }
            } else if (op == P_BERT) {
                for (uint32_t c : w) {
                    if (uni::is_white(c)) flush();
                    else if (uni::is_punct(c)) { flush(); out.push_back({c}); }
                    else cur.push_back(c);
                }
                flush();
            } else if (op == P_SPLIT) {
                for (uint32_t c : w) {
                    if (uni::is_white(c)) flush();
                    else cur.push_back(c);
                }
                flush();
            } else {  // Whitespace: \w+|[^\w\s]+
                int kind = 0;  // 1 word run, 2 symbol run
                for (uint32_t c : w) {
                    const int k = uni::is_word(c) ? 1 : (uni::is_white(c) ? 0 : 2);
                    if (k != kind) flush();
                    kind = k;
                    if (k) cur.push_back(c);
                }
                flush();
            }
        }
    }

    // BPE of one pre-token (models/bpe/word.rs merge_all): symbols = characters; repeatedly apply the merge of lowest rank, leftmost
    // first, one occurrence at a time; a character outside the vocabulary becomes unk (or is dropped when there is no unk token)
    void bpe_tokens(const std::vector<uint32_t>& word, std::vector<uint32_t>& ids) const {
        std::vector<uint32_t> sym;
        std::string s;
        for (uint32_t c : word) {
            s.clear();
            uni::encode_append(c, s);
            uint32_t id;
            if (lookup(s, id)) sym.push_back(id);
            else if (has_unk_) sym.push_back(unk_id_);
        }
        while (sym.size() > 1) {
            uint32_t best_rank = 0xFFFFFFFFu, best_id = 0;
            size_t best_pos = 0;
            for (size_t i = 0; i + 1 < sym.size(); ++i) {
                auto it = merges_.find((static_cast<uint64_t>(sym[i]) << 32) | sym[i + 1]);
                if (it != merges_.end() && it->second.first < best_rank) { best_rank = it->second.first; best_id = it->second.second; best_pos = i; }
            }
            if (best_rank == 0xFFFFFFFFu) break;
            sym[best_pos] = best_id;
            sym.erase(sym.begin() + static_cast<long>(best_pos) + 1);
        }
        ids.insert(ids.end(), sym.begin(), sym.end());
    }

    void model_tokens(const std::vector<uint32_t>& word, std::vector<uint32_t>& ids) const {
        std::string s;
        if (bpe_) {
            bpe_tokens(word, ids);
            return;
        }
        if (!wordpiece_) {  // WordLevel: whole word or unk
            for (uint32_t c : word) uni::encode_append(c, s);
            uint32_t id;
            if (lookup(s, id)) ids.push_back(id);
            else if (has_unk_) ids.push_back(unk_id_);
            return;
        }
        // WordPiece: greedy longest-match-first; any unmatched remainder turns the whole word into unk
        if (static_cast<int>(word.size()) > max_chars_) {
            if (has_unk_) ids.push_back(unk_id_);
            return;
        }
        std::vector<uint32_t> sub;
        size_t start = 0;
        while (start < word.size()) {
            size_t end = word.size();
            bool found = false;
            uint32_t id = 0;
            while (start < end) {
                s.clear();
                if (start > 0) s = prefix_;
                for (size_t i = start; i < end; ++i) uni::encode_append(word[i], s);
                if (lookup(s, id)) { found = true; break; }
                --end;
            }
            if (!found) {
                if (has_unk_) ids.push_back(unk_id_);
                return;
            }
            sub.push_back(id);
            start = end;
        }
        ids.insert(ids.end(), sub.begin(), sub.end());
    }

    void tokenize_segment(const std::vector<uint32_t>& cps, std::vector<uint32_t>& ids) const {
        std::vector<std::vector<uint32_t>> words{normalize(cps)}, tmp;
        for (PreOp op : pre_) {
            pre_split(op, words, tmp);
            words.swap(tmp);
        }
        for (const auto& w : words)
            if (!w.empty()) model_tokens(w, ids);
    }

    void tokenize_into(const std::string& text, std::vector<uint32_t>& ids) const {
        // added tokens are cut out of the raw text first (leftmost, longest content first), the rest goes through the pipeline
        size_t pos = 0;
        while (pos < text.size()) {
            size_t best = std::string::npos;
            const Added* hit = nullptr;
            for (const Added& a : added_) {
                const size_t f = text.find(a.content, pos);
                if (f != std::string::npos && (f < best)) { best = f; hit = &a; }
            }
            const size_t seg_end = hit ? best : text.size();
            if (seg_end > pos) tokenize_segment(uni::decode(text.substr(pos, seg_end - pos)), ids);
            if (!hit) break;
            ids.push_back(hit->id);
            pos = best + hit->content.size();
        }
    }

    int max_length_;
    bool wordpiece_ = true, has_unk_ = false, bpe_ = false, byte_prefix_space_ = true;
    std::unordered_map<uint64_t, std::pair<uint32_t, uint32_t>> merges_;  // (left id, right id) -> (rank, merged id)
    uint32_t byte_char_[256] = {};
    std::string prefix_ = "##", unk_ = "[UNK]";
    uint32_t unk_id_ = 0;
    int max_chars_ = 100;
    std::unordered_map<std::string, uint32_t> vocab_;
    std::vector<NormOp> norm_;
    std::vector<PreOp> pre_;
    std::vector<Added> added_;
    std::vector<Piece> single_, pair_;
};

}  // namespace kj
