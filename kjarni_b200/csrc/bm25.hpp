// BM25 keyword index of a segment and reciprocal-rank fusion: the CPU half of the Searcher's Keyword / Hybrid modes.
// The reference keeps both on the host as well (kjarni-search/src/bm25.rs, kjarni-search/src/hybrid.rs); the semantic half of
// a hybrid query is the GPU scan (index.cu).  This file restates
//   Bm25Index::{add_document, search, calculate_score}, tokenize      kjarni-search/src/bm25.rs:85-198
//   hybrid_search (RRF, k = 60)                                       kjarni-search/src/hybrid.rs:3-31
//   bm25.bin = bincode 1.3 (default options) of Bm25Index             kjarni-rag/src/segment.rs:163-165,224-227
// bincode default: little-endian fixed-width integers, usize as u64, sequence/map/string lengths as u64, f32 as 4 bytes, struct
// fields in declaration order:
//   doc_frequencies HashMap<String,usize> | doc_lengths Vec<usize> | avg_doc_length f32 | total_docs usize |
//   inverted_index HashMap<String,Vec<(usize,usize)>> | params {k1,b,epsilon: f32} | token_to_docs HashMap<String,HashSet<usize>> |
//   total_length usize
// Ties: the reference collects scores into a HashMap before its (stable) sort, so the order of equal scores is the map's iteration
// order, i.e. unspecified; here equal scores come out in ascending document id.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "host_util.hpp"
#include "tokenizer.hpp"

namespace kj {

struct Bm25Index {
    std::unordered_map<std::string, uint64_t> doc_frequencies;
    std::vector<uint64_t> doc_lengths;
    float avg_doc_length = 0.0f;
    uint64_t total_docs = 0;
    std::unordered_map<std::string, std::vector<std::pair<uint64_t, uint64_t>>> inverted_index;  // term -> (doc, tf), in insertion order
    float k1 = 1.2f, b = 0.75f, epsilon = 0.25f;
    uint64_t total_length = 0;

    // text.to_lowercase().split(|c| !c.is_alphanumeric()).filter(|s| s.len() >= 2)   (bm25.rs:192-198; len() is in bytes)
    static std::vector<std::string> tokenize(const std::string& text) {
        std::vector<std::string> out;
        std::vector<uint32_t> low;
        for (uint32_t c : uni::decode(text)) uni::lower_append(c, low);
        std::string cur;
        for (uint32_t c : low) {
            if (uni::is_alnum(c)) uni::encode_append(c, cur);
            else {
                if (cur.size() >= 2) out.push_back(cur);
                cur.clear();
            }
        }
        if (cur.size() >= 2) out.push_back(cur);
        return out;
    }

    // Bm25Index::add_document, bm25.rs:115-147 (term order inside one document is irrelevant: one posting per term)
    void add_document(uint64_t doc_id, const std::string& text) {
        const std::vector<std::string> tokens = tokenize(text);
        if (doc_id >= doc_lengths.size()) doc_lengths.resize(doc_id + 1, 0);
        doc_lengths[doc_id] = tokens.size();
        std::unordered_map<std::string, uint64_t> counts;
        for (const std::string& t : tokens) ++counts[t];
        for (auto& kv : counts) {
            inverted_index[kv.first].emplace_back(doc_id, kv.second);
            ++doc_frequencies[kv.first];
        }
        total_docs = std::max<uint64_t>(total_docs, doc_id + 1);
        total_length += tokens.size();
        avg_doc_length = static_cast<float>(total_length) / static_cast<float>(total_docs);
    }

    // Bm25Index::search, bm25.rs:85-113: every document's score is the sum over the query tokens IN QUERY ORDER (duplicates count
    // again) of idf * tf * (k1 + 1) / (tf + k1 * length_norm), all in f32; documents with score > 0, best first, at most `limit`.
    std::vector<std::pair<uint64_t, float>> search(const std::string& query, size_t limit) const {
        std::vector<std::pair<uint64_t, float>> res;
        if (total_docs == 0) return res;
        const std::vector<std::string> q = tokenize(query);
        if (q.empty()) return res;
        std::unordered_map<uint64_t, float> score;
        for (const std::string& term : q) {
            auto it = inverted_index.find(term);
            if (it == inverted_index.end()) continue;
            auto dfi = doc_frequencies.find(term);
            const float df = dfi == doc_frequencies.end() ? 0.0f : static_cast<float>(dfi->second);
            if (df == 0.0f) continue;
            const float idf = logf((static_cast<float>(total_docs) - df + 0.5f) / (df + 0.5f) + 1.0f);
            std::vector<bool> seen;  // get_term_frequency takes the FIRST posting of a document (find), bm25.rs:178-189
            for (const auto& post : it->second) {
                const uint64_t doc = post.first;
                if (doc >= total_docs || doc >= doc_lengths.size()) continue;
                if (doc < seen.size() && seen[doc]) continue;
                if (doc >= seen.size()) seen.resize(doc + 1, false);
                seen[doc] = true;
                const float tf = static_cast<float>(post.second);
                if (tf == 0.0f) continue;
                const float length_norm = 1.0f - b + b * (static_cast<float>(doc_lengths[doc]) / avg_doc_length);
                const float ntf = (tf * (k1 + 1.0f)) / (tf + k1 * length_norm);
                score[doc] += idf * ntf;
            }
        }
        for (auto& kv : score)
            if (kv.second > 0.0f) res.emplace_back(kv.first, kv.second);
        std::sort(res.begin(), res.end(), [](const auto& x, const auto& y) { return x.second != y.second ? x.second > y.second : x.first < y.first; });
        if (res.size() > limit) res.resize(limit);
        return res;
    }

    // ---- bincode
    static Bm25Index load(const std::string& path) {
        const std::string buf = read_text_file(path, KJC_LOAD_FAILED);
        size_t p = 0;
        auto need = [&](size_t n) {
            if (p + n > buf.size()) throw Error(KJC_LOAD_FAILED, "bm25.bin: truncated (" + path + ")");
        };
        auto u64 = [&]() {
            need(8);
            uint64_t v;
            memcpy(&v, buf.data() + p, 8);
            p += 8;
            return v;
        };
        auto f32 = [&]() {
            need(4);
            float v;
            memcpy(&v, buf.data() + p, 4);
            p += 4;
            return v;
        };
        auto str = [&]() {
            const uint64_t n = u64();
            need(n);
            std::string s(buf.data() + p, n);
            p += n;
            return s;
        };
        Bm25Index ix;
        for (uint64_t n = u64(); n > 0; --n) {
            std::string k = str();
            ix.doc_frequencies[std::move(k)] = u64();
        }
        const uint64_t nl = u64();
        need(nl * 8);
        ix.doc_lengths.resize(nl);
        for (uint64_t i = 0; i < nl; ++i) ix.doc_lengths[i] = u64();
        ix.avg_doc_length = f32();
        ix.total_docs = u64();
        for (uint64_t n = u64(); n > 0; --n) {
            std::string k = str();
            const uint64_t m = u64();
            need(m * 16);
            auto& v = ix.inverted_index[std::move(k)];
            v.reserve(m);
            for (uint64_t i = 0; i < m; ++i) {
                const uint64_t d = u64();
                v.emplace_back(d, u64());
            }
        }
        ix.k1 = f32();
        ix.b = f32();
        ix.epsilon = f32();
        for (uint64_t n = u64(); n > 0; --n) {  // token_to_docs: never filled by add_document, skipped
            str();
            const uint64_t m = u64();
            need(m * 8);
            p += m * 8;
        }
        ix.total_length = p + 8 <= buf.size() ? u64() : 0;  // #[serde(default)]: older files end before it (bincode would fail; tolerate)
        return ix;
    }
    std::string to_bincode() const {
        std::string o;
        auto u64 = [&](uint64_t v) { o.append(reinterpret_cast<const char*>(&v), 8); };
        auto f32 = [&](float v) { o.append(reinterpret_cast<const char*>(&v), 4); };
        auto str = [&](const std::string& s) { u64(s.size()); o += s; };
        u64(doc_frequencies.size());
        for (auto& kv : doc_frequencies) { str(kv.first); u64(kv.second); }
        u64(doc_lengths.size());
        for (uint64_t v : doc_lengths) u64(v);
        f32(avg_doc_length);
        u64(total_docs);
        u64(inverted_index.size());
        for (auto& kv : inverted_index) {
            str(kv.first);
            u64(kv.second.size());
            for (auto& pr : kv.second) { u64(pr.first); u64(pr.second); }
        }
        f32(k1); f32(b); f32(epsilon);
        u64(0);  // token_to_docs
        u64(total_length);
        return o;
    }
};

// hybrid_search, kjarni-search/src/hybrid.rs:3-31: score(doc) = sum over the two ranked lists of 1 / (60 + rank), rank from 1;
// best first (equal scores: ascending id -- the reference's order there is its HashMap's), truncated to `limit`.
inline std::vector<std::pair<uint64_t, float>> rrf_fuse(const std::vector<std::pair<uint64_t, float>>& keyword,
                                                        const std::vector<std::pair<uint64_t, float>>& semantic, size_t limit) {
    std::unordered_map<uint64_t, float> s;
    const float k = 60.0f;
    for (size_t r = 0; r < keyword.size(); ++r) s[keyword[r].first] += 1.0f / (k + static_cast<float>(r + 1));
    for (size_t r = 0; r < semantic.size(); ++r) s[semantic[r].first] += 1.0f / (k + static_cast<float>(r + 1));
    std::vector<std::pair<uint64_t, float>> out(s.begin(), s.end());
    std::sort(out.begin(), out.end(), [](const auto& x, const auto& y) { return x.second != y.second ? x.second > y.second : x.first < y.first; });
    if (out.size() > limit) out.resize(limit);
    return out;
}

}  // namespace kj
