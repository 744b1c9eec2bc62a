// Public kjarni-ffi C ABI (include/kjarni_ffi.h) over the B200 backend: the text-level handles the C#/Go/Python bindings
// use -- Embedder, Classifier, Reranker, Searcher -- built from Tokenizer (tokenizer.hpp) + Encoder (encoder.cu) + Index
// (index.cu).  Each function restates the control flow of its reference counterpart in kjarni-ffi/src/*.rs (cited in the
// header) with the model forward, the pooling/head and the cosine scan running on the GPU.
#include <fnmatch.h>

#include <cmath>
#include <map>
#include <mutex>

#include "../../include/kjarni_ffi.h"
#include <atomic>
#include <chrono>

#include "multi.hpp"
#include "indexer.hpp"
#include "tokenizer.hpp"

namespace kj {
namespace {

struct Fail {
    int code;
    std::string msg;
};

bool file_ok(const std::string& p) {
    struct stat sb;
    return stat(p.c_str(), &sb) == 0 && S_ISREG(sb.st_mode);
}
std::string lower(std::string s) {
    for (char& c : s) c = static_cast<char>(tolower(static_cast<unsigned char>(c)));
    return s;
}

// ModelType::from_cli_name + repo_id().replace('/', "_") for the encoder-family entries of the registry
// (kjarni-transformers/src/models/registry.rs:226-262,754-800,809-811,851-860; weights_url lines :318-478)
const char* registry_dir(const std::string& name_in) {
    static const std::pair<const char*, const char*> table[] = {
        {"minilm-l6-v2", "sentence-transformers_all-MiniLM-L6-v2"},
        {"all-minilm-l6-v2", "sentence-transformers_all-MiniLM-L6-v2"},
        {"sentence-transformers/all-minilm-l6-v2", "sentence-transformers_all-MiniLM-L6-v2"},
        {"mpnet-base-v2", "sentence-transformers_all-mpnet-base-v2"},
        {"all-mpnet-base-v2", "sentence-transformers_all-mpnet-base-v2"},
        {"sentence-transformers/all-mpnet-base-v2", "sentence-transformers_all-mpnet-base-v2"},
        {"distilbert-base", "distilbert-base-cased-distilled-squad_resolve"},  // the URL heuristic takes parts[3]/parts[4] (:851-860)
        {"minilm-l6-v2-cross-encoder", "cross-encoder_ms-marco-MiniLM-L-6-v2"},
        {"ms-marco-minilm-l-6-v2", "cross-encoder_ms-marco-MiniLM-L-6-v2"},
        {"cross-encoder/ms-marco-minilm-l-6-v2", "cross-encoder_ms-marco-MiniLM-L-6-v2"},
        {"sentiment", "distilbert_distilbert-base-uncased-finetuned-sst-2-english"},  // Classifier::builder("sentiment") default preset
        {"distilbert-sentiment", "distilbert_distilbert-base-uncased-finetuned-sst-2-english"},
        {"distilbert-base-uncased-finetuned-sst-2-english", "distilbert_distilbert-base-uncased-finetuned-sst-2-english"},
        {"roberta-sentiment", "olafuraron_twitter-roberta-base-sentiment-latest-safetensors"},
        {"twitter-roberta-base-sentiment-latest", "olafuraron_twitter-roberta-base-sentiment-latest-safetensors"},
        {"bert-sentiment-multilingual", "olafuraron_bert-base-multilingual-uncased-sentiment-safetensors"},
        {"bert-base-multilingual-uncased-sentiment", "olafuraron_bert-base-multilingual-uncased-sentiment-safetensors"},
        {"roberta-emotions", "SamLowe_roberta-base-go_emotions"},
        {"roberta-base-go_emotions", "SamLowe_roberta-base-go_emotions"},
        {"distilroberta-emotion", "olafuraron_emotion-english-distilroberta-base-safetensors"},
        {"emotion-english-distilroberta-base", "olafuraron_emotion-english-distilroberta-base-safetensors"},
        {"toxic-bert", "olafuraron_toxic-bert-safetensors"},
    };
    const std::string name = lower(name_in);
    for (auto& e : table)
        if (name == e.first) return e.second;
    return nullptr;
}

std::string default_cache_dir() {
    // get_default_cache_dir, registry.rs:958-966: $KJARNI_CACHE_DIR, else dirs::cache_dir()/kjarni
    if (const char* e = getenv("KJARNI_CACHE_DIR")) return e;
    if (const char* x = getenv("XDG_CACHE_HOME"))
        if (*x) return std::string(x) + "/kjarni";
    const char* home = getenv("HOME");
    return std::string(home ? home : ".") + "/.cache/kjarni";
}

// Resolves the model directory the way the builders do; throws Fail{ModelNotFound} when it is not on disk (no download here).
std::string resolve_model_dir(const char* cache_dir, const char* model_name, const char* model_path, const char* default_name) {
    if (model_path && *model_path) {
        struct stat sb;
        if (stat(model_path, &sb) != 0 || !S_ISDIR(sb.st_mode)) throw Fail{KJARNI_ERROR_MODEL_NOT_FOUND, std::string("Model path does not exist: ") + model_path};
        return model_path;
    }
    const std::string name = (model_name && *model_name) ? model_name : default_name;
    const char* dir = registry_dir(name);
    if (!dir) throw Fail{KJARNI_ERROR_MODEL_NOT_FOUND, "Unknown model: '" + name + "'"};
    const std::string root = (cache_dir && *cache_dir) ? cache_dir : default_cache_dir();
    const std::string d = root + "/" + dir;
    // ModelType::is_downloaded, registry.rs:814-827
    if (!file_ok(d + "/config.json") || !file_ok(d + "/tokenizer.json") || !file_ok(d + "/model.safetensors"))
        throw Fail{KJARNI_ERROR_MODEL_NOT_FOUND, "Model '" + name + "' not downloaded (expected config.json, tokenizer.json, model.safetensors in " + d +
                                                     "); this backend never downloads"};
    return d;
}

// The reference's configs only say KJARNI_DEVICE_GPU; which GPUs that means is deployment configuration:
// KJARNI_GPU_DEVICES = "all" | comma-separated device indices (default "0").  With more than one device every model holds one weight
// replica per GPU (batches are split by rows, no collective) and every opened index is row-sharded over the same GPUs (multi.hpp).
std::vector<int> shim_devices() {
    std::vector<int> d;
    const char* e = getenv("KJARNI_GPU_DEVICES");
    if (e && !strcmp(e, "all")) {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
        for (int i = 0; i < n; ++i) d.push_back(i);
    } else if (e && *e) {
        for (const char* p = e; *p;) {
            char* end = nullptr;
            const long v = strtol(p, &end, 10);
            if (end == p) throw Fail{KJARNI_ERROR_INVALID_CONFIG, "KJARNI_GPU_DEVICES must be \"all\" or a comma-separated list of device indices"};
            d.push_back(static_cast<int>(v));
            p = (*end == ',') ? end + 1 : end;
            if (*end && *end != ',') throw Fail{KJARNI_ERROR_INVALID_CONFIG, "KJARNI_GPU_DEVICES must be \"all\" or a comma-separated list of device indices"};
        }
    }
    if (d.empty()) d.push_back(0);
    return d;
}

// One text model: tokenizer + encoder replica(s) (EncoderLoader::load_from_pretrained, pipeline/encoder/loader.rs:82-141)
struct TextModel {
    std::unique_ptr<EncoderGroup> grp;
    Encoder* enc = nullptr;  // replica 0: info / labels
    std::unique_ptr<Tokenizer> tok;
    std::string name;

    TextModel(const std::string& dir, const std::string& nm) : name(nm) {
        if (!file_ok(dir + "/tokenizer.json")) throw Fail{KJARNI_ERROR_LOAD_FAILED, "Tokenizer not found at \"" + dir + "/tokenizer.json\""};
        const std::vector<int> devs = shim_devices();
        grp.reset(new EncoderGroup(dir, devs.data(), static_cast<int>(devs.size())));
        enc = &grp->replica(0);
        tok.reset(new Tokenizer(dir + "/tokenizer.json", enc->config_max_seq_len()));  // truncation max_length = meta.max_seq_len (loader.rs:108)
    }

    // Result rows handed over while they still sit in the encoder's pinned staging buffer: rows[i] belongs to text index[i]
    using TextSink = std::function<void(const float* rows, size_t n, const size_t* index)>;

    // texts (+ optional second segments) -> [n, out_cols] through one encoder forward (into `out`, or into `sink` when given)
    void run(const std::vector<std::string>& a, const std::vector<std::string>& b, const KjcForwardOptions& o, bool with_types, std::vector<float>& out,
             size_t out_cols, const TextSink* sink = nullptr) {
        std::vector<uint32_t> ids, types;
        std::vector<float> mask;
        int S = 0;
        tok->encode_batch(a, b, true, ids, mask, types, S);
        if (S == 0) throw Error(KJC_INFERENCE_FAILED, "Tokenizer produced an empty batch");
        if (!sink) out.assign(a.size() * out_cols, 0.f);
        const bool types_ok = with_types && enc->info().type_vocab_size > 0;
        const size_t n = a.size();
        // Length-bucketed batching: BatchLongest pads every text of the call to the longest one, and padded
        // positions are pure waste on the GPU (masked keys contribute exactly 0 to every valid token, so pooled rows and logits
        // do not depend on the padding).  Texts are grouped by token count rounded up to 16 and each group runs at its own
        // length; results land in the caller's order.  A call whose texts share one bucket is a single forward as before.
        std::vector<int> len(n, 0);
        bool one_bucket = true;
        for (size_t i = 0; i < n; ++i) {
            int l = 0;
            for (int k = 0; k < S; ++k) l += mask[i * S + k] != 0.0f;
            len[i] = std::max(l, 1);
            if ((len[i] + 15) / 16 != (len[0] + 15) / 16) one_bucket = false;
        }
        std::vector<size_t> order(n);
        for (size_t i = 0; i < n; ++i) order[i] = i;
        static const bool no_buckets = getenv("KJC_NO_LENGTH_BUCKETS") != nullptr;
        if (one_bucket || no_buckets) {
            const Encoder::RowSink rs = [&](const float* rows, size_t first, size_t cnt) { (*sink)(rows, cnt, order.data() + first); };
            grp->forward_host(ids.data(), mask.data(), types_ok ? types.data() : nullptr, static_cast<int>(n), S, o, sink ? nullptr : out.data(), sink ? &rs : nullptr);
            return;
        }
        std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return len[x] < len[y]; });
        std::vector<uint32_t> gi, gt;
        std::vector<float> gm, go;
        for (size_t g0 = 0; g0 < n;) {
            const int bucket = (len[order[g0]] + 15) / 16;
            size_t g1 = g0;
            while (g1 < n && (len[order[g1]] + 15) / 16 == bucket) ++g1;
            const int Sg = std::min(S, len[order[g1 - 1]]);  // longest text of the group: tokens are left-aligned, the rest is padding
            const size_t nb = g1 - g0;
            gi.assign(nb * Sg, 0u);
            gt.assign(nb * Sg, 0u);
            gm.assign(nb * Sg, 0.f);
            go.assign(nb * out_cols, 0.f);
            for (size_t r = 0; r < nb; ++r) {
                const size_t src = order[g0 + r];
                memcpy(&gi[r * Sg], &ids[src * S], Sg * sizeof(uint32_t));
                memcpy(&gt[r * Sg], &types[src * S], Sg * sizeof(uint32_t));
                memcpy(&gm[r * Sg], &mask[src * S], Sg * sizeof(float));
            }
            if (sink) {
                const Encoder::RowSink rs = [&](const float* rows, size_t first, size_t cnt) { (*sink)(rows, cnt, order.data() + g0 + first); };
                grp->forward_host(gi.data(), gm.data(), types_ok ? gt.data() : nullptr, static_cast<int>(nb), Sg, o, nullptr, &rs);
            } else {
                grp->forward_host(gi.data(), gm.data(), types_ok ? gt.data() : nullptr, static_cast<int>(nb), Sg, o, go.data());
                for (size_t r = 0; r < nb; ++r) memcpy(&out[order[g0 + r] * out_cols], &go[r * out_cols], out_cols * sizeof(float));
            }
            g0 = g1;
        }
    }
};

int device_check(KjarniDevice d) {
    if (d != KJARNI_DEVICE_GPU) throw Fail{KJARNI_ERROR_INVALID_CONFIG, "libkjarni_cuda serves KJARNI_DEVICE_GPU only: there is no CPU path in this library"};
    return 0;
}

template <typename F>
KjarniErrorCode guarded(KjarniErrorCode on_error, F&& f, bool passthrough = false) {
    try {
        f();
        return KJARNI_OK;
    } catch (const Fail& e) {
        set_last_error(e.msg);
        return static_cast<KjarniErrorCode>(e.code);
    } catch (const Error& e) {
        set_last_error(e.what());
        // construction paths surface the backend's own status; inference paths collapse to InferenceFailed as the reference does
        return (passthrough || on_error == KJARNI_ERROR_LOAD_FAILED) ? static_cast<KjarniErrorCode>(e.status) : on_error;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return on_error;
    } catch (...) {
        set_last_error("unknown error");
        return KJARNI_ERROR_UNKNOWN;
    }
}

char* dup_cstr(const std::string& s) {
    char* p = static_cast<char*>(malloc(s.size() + 1));
    if (!p) throw std::bad_alloc();
    memcpy(p, s.c_str(), s.size() + 1);
    return p;
}
float* dup_floats(const float* src, size_t n) {
    float* p = static_cast<float*>(malloc(std::max<size_t>(n, 1) * sizeof(float)));
    if (!p) throw std::bad_alloc();
    memcpy(p, src, n * sizeof(float));
    return p;
}
std::string json_escape(const std::string& s) {
    std::string o;
    for (unsigned char c : s) {
        if (c == '"') o += "\\\"";
        else if (c == '\\') o += "\\\\";
        else if (c == '\n') o += "\\n";
        else if (c == '\r') o += "\\r";
        else if (c == '\t') o += "\\t";
        else if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
        else o += static_cast<char>(c);
    }
    return o;
}

}  // namespace
}  // namespace kj

using namespace kj;

struct KjarniEmbedder {
    std::unique_ptr<TextModel> m;
    bool normalize;
    std::mutex mu;
    void embed(const std::vector<std::string>& texts, std::vector<float>& out) {
        std::lock_guard<std::mutex> lock(mu);
        const KjcForwardOptions o{KJC_OUT_POOLED, KJC_POOL_MEAN, normalize ? 1 : 0, KJC_MASK_AUTO};
        // Embedder path: no token-type ids, row 0 of the type table for every token (cpu/encoder/traits.rs:80)
        m->run(texts, {}, o, false, out, static_cast<size_t>(m->enc->info().hidden_size));
    }
};
struct KjarniClassifier {
    std::unique_ptr<TextModel> m;
    std::vector<std::string> labels;
    bool multi_label;
    std::mutex mu;
};
struct KjarniReranker {
    std::unique_ptr<TextModel> m;
    std::mutex mu;
    void score(const std::string& query, const std::vector<std::string>& docs, std::vector<float>& out) {
        std::lock_guard<std::mutex> lock(mu);
        const KjcForwardOptions o{KJC_OUT_LOGITS, KJC_POOL_CLS, 0, KJC_MASK_ALLOC};
        std::vector<std::string> q(docs.size(), query);
        std::vector<float> logits;
        const size_t C = static_cast<size_t>(m->enc->info().num_labels);
        m->run(q, docs, o, true, logits, C);
        out.resize(docs.size());
        for (size_t i = 0; i < docs.size(); ++i) out[i] = logits[i * C];  // logits.column(0), cross_encoder/model.rs:239
    }
};
struct KjarniSearcher {
    std::unique_ptr<KjarniEmbedder> embedder;
    std::unique_ptr<KjarniReranker> reranker;
    KjarniSearchMode default_mode;
    size_t default_top_k;
    std::mutex mu;
    // GPU shards of the index directories searched so far (the reference re-opens the index per call: IndexReader::open is an
    // mmap there; here it is an upload, so it is kept)
    struct Opened {
        std::unique_ptr<ShardedIndex> idx;
        IndexDir dir;
        std::string fingerprint;  // segment list + file sizes / mtimes at open: a rebuilt or appended index is re-opened
        std::vector<std::unique_ptr<Bm25Index>> bm25;  // per segment, loaded on the first keyword / hybrid query
        uint64_t last_use = 0;
    };
    std::map<std::string, Opened> opened;
    uint64_t use_counter = 0;
    std::string model_name, reranker_name;
};
struct KjarniCancelToken {
    std::atomic<bool> flag{false};
};
struct KjarniIndexer {
    std::unique_ptr<KjarniEmbedder> embedder;
    std::string model_name;
    TextSplitter splitter;
    LoaderOptions loader;
    size_t batch_size = 32, max_docs_per_segment = 10000;
    bool quiet = false;
    std::mutex mu;
};

namespace {

// Segment::get_document / get_metadata (kjarni-rag/src/segment.rs:264-304) for one global id
void fetch_doc(const IndexDir& d, uint64_t gid, std::string& text, std::vector<std::pair<std::string, std::string>>& meta) {
    const IndexDirSegment* seg = nullptr;
    for (const IndexDirSegment& s : d.segments)
        if (gid >= s.global_base && gid < s.global_base + s.doc_count) { seg = &s; break; }
    if (!seg) throw Error(KJC_INFERENCE_FAILED, "Document ID out of range");
    const uint64_t local = gid - seg->global_base;
    // docs.idx: bincode Vec<u64> = u64 length + offsets
    FILE* f = fopen((seg->dir + "/docs.idx").c_str(), "rb");
    if (!f) throw Error(KJC_INFERENCE_FAILED, "cannot open docs.idx");
    uint64_t n = 0, off[2] = {0, 0};
    bool ok = fread(&n, 8, 1, f) == 1 && local < n && fseek(f, static_cast<long>(8 + 8 * local), SEEK_SET) == 0 && fread(&off[0], 8, 1, f) == 1;
    const bool has_next = ok && local + 1 < n && fread(&off[1], 8, 1, f) == 1;
    fclose(f);
    if (!ok) throw Error(KJC_INFERENCE_FAILED, "Document ID out of range");
    FILE* g = fopen((seg->dir + "/docs.bin").c_str(), "rb");
    if (!g) throw Error(KJC_INFERENCE_FAILED, "cannot open docs.bin");
    uint64_t end;
    if (has_next) end = off[1] - 1;  // -1 for the newline
    else {
        fseek(g, 0, SEEK_END);
        end = static_cast<uint64_t>(ftell(g)) - 1;
    }
    text.assign(end > off[0] ? end - off[0] : 0, '\0');
    fseek(g, static_cast<long>(off[0]), SEEK_SET);
    if (!text.empty() && fread(&text[0], 1, text.size(), g) != text.size()) { fclose(g); throw Error(KJC_INFERENCE_FAILED, "short read in docs.bin"); }
    fclose(g);
    // metadata.jsonl: line `local` is a JSON object of strings
    meta.clear();
    FILE* h = fopen((seg->dir + "/metadata.jsonl").c_str(), "rb");
    if (!h) throw Error(KJC_INFERENCE_FAILED, "cannot open metadata.jsonl");
    std::string line;
    uint64_t ln = 0;
    int c;
    bool found = false;
    while ((c = fgetc(h)) != EOF) {
        if (c == '\n') {
            if (ln == local) { found = true; break; }
            ++ln;
            line.clear();
        } else if (ln == local) line.push_back(static_cast<char>(c));
    }
    if (!found && ln == local && !line.empty()) found = true;
    fclose(h);
    if (!found) throw Error(KJC_INFERENCE_FAILED, "Document ID out of range");
    const Json j = JsonParser(line.data(), line.size()).parse();
    if (j.type == Json::Obj)
        for (auto& kv : j.obj)
            if (kv.second.type == Json::Str) meta.emplace_back(kv.first, kv.second.str);
}

// What IndexReader::open would see: every segment directory with the size and mtime of its vectors.bin and segment.json
std::string index_fingerprint(const IndexDir& d) {
    std::string f = std::to_string(d.dimension) + ":" + std::to_string(d.total_rows);
    for (const IndexDirSegment& s : d.segments) {
        f += "|" + s.dir + ":" + std::to_string(s.doc_count);
        for (const char* name : {"/vectors.bin", "/segment.json"}) {
            struct stat sb;
            if (stat((s.dir + name).c_str(), &sb) == 0)
                f += ":" + std::to_string(static_cast<long long>(sb.st_size)) + "@" + std::to_string(static_cast<long long>(sb.st_mtim.tv_sec)) + "." +
                     std::to_string(static_cast<long long>(sb.st_mtim.tv_nsec));
        }
    }
    return f;
}

// IndexReader::search_keywords (kjarni-rag/src/index_reader.rs:230-245): per-segment BM25 top-`limit`, concatenated in segment order,
// stable sort by score descending, truncated; ids are global.  `bm25` caches the segments' indexes (nullptr entries are loaded).
std::vector<std::pair<uint64_t, float>> keyword_search(const IndexDir& d, std::vector<std::unique_ptr<Bm25Index>>& bm25, const std::string& query, size_t limit) {
    if (bm25.size() != d.segments.size()) {
        bm25.clear();
        bm25.resize(d.segments.size());
    }
    std::vector<std::pair<uint64_t, float>> all;
    for (size_t i = 0; i < d.segments.size(); ++i) {
        if (!bm25[i]) bm25[i].reset(new Bm25Index(Bm25Index::load(d.segments[i].dir + "/bm25.bin")));
        for (const auto& hit : bm25[i]->search(query, limit)) all.emplace_back(d.segments[i].global_base + hit.first, hit.second);
    }
    std::stable_sort(all.begin(), all.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
    if (all.size() > limit) all.resize(limit);
    return all;
}

struct SearchHit {
    float score;
    uint64_t id;
    std::string text;
    std::vector<std::pair<std::string, std::string>> meta;
};

// MetadataFilter::matches for the two filters the C ABI can express (kjarni-rag/src/index_reader.rs:27-80)
bool filter_matches(const KjarniSearchOptions& o, const SearchHit& h);

void fill_results(std::vector<SearchHit>& hits, KjarniSearchResults* out) {
    if (hits.empty()) return;
    KjarniSearchResult* r = static_cast<KjarniSearchResult*>(calloc(hits.size(), sizeof(KjarniSearchResult)));
    if (!r) throw std::bad_alloc();
    for (size_t i = 0; i < hits.size(); ++i) {
        std::string mj = "{";
        for (size_t k = 0; k < hits[i].meta.size(); ++k)
            mj += std::string(k ? "," : "") + "\"" + json_escape(hits[i].meta[k].first) + "\":\"" + json_escape(hits[i].meta[k].second) + "\"";
        mj += "}";
        r[i] = KjarniSearchResult{hits[i].score, static_cast<size_t>(hits[i].id), dup_cstr(hits[i].text), dup_cstr(mj)};
    }
    out->results = r;
    out->len = hits.size();
}

size_t copy_name(const std::string& name, char* buf, size_t buf_len) {  // the buffer protocol of kjarni_*_model_name
    const size_t required = name.size() + 1;
    if (!buf || buf_len == 0) return required;
    const size_t n = std::min(name.size(), buf_len - 1);
    memcpy(buf, name.data(), n);
    buf[n] = 0;
    return required;
}

std::vector<std::string> split_csv(const char* s, bool strip_dot) {
    std::vector<std::string> out;
    if (!s) return out;
    std::string cur;
    auto push = [&] {
        size_t a = 0, b = cur.size();
        while (a < b && isspace(static_cast<unsigned char>(cur[a]))) ++a;
        while (b > a && isspace(static_cast<unsigned char>(cur[b - 1]))) --b;
        std::string t = cur.substr(a, b - a);
        if (strip_dot) {
            t = lower(t);
            while (!t.empty() && t[0] == '.') t.erase(0, 1);
        }
        out.push_back(t);
        cur.clear();
    };
    for (const char* p = s; *p; ++p) {
        if (*p == ',') push();
        else cur.push_back(*p);
    }
    push();
    return out;
}

const std::string* meta_get(const std::vector<std::pair<std::string, std::string>>& m, const std::string& k) {
    for (auto& kv : m)
        if (kv.first == k) return &kv.second;
    return nullptr;
}

bool filter_matches(const KjarniSearchOptions& o, const SearchHit& h) {
    if (o.filter_key && o.filter_value) {
        const std::string* v = meta_get(h.meta, o.filter_key);
        if (!v || *v != o.filter_value) return false;
    }
    if (o.source_pattern && uni::valid_utf8(o.source_pattern)) {
        const std::string* src = meta_get(h.meta, "source");
        if (!src) return false;
        const std::string pat = o.source_pattern;
        const std::string name = pat.find('/') != std::string::npos ? *src : src->substr(src->find_last_of('/') + 1);
        if (fnmatch(pat.c_str(), name.c_str(), 0) != 0) return false;
    }
    return true;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------ runtime / errors / frees
KjarniErrorCode kjarni_init(void) { return KJARNI_OK; }
void kjarni_shutdown(void) {}
const char* kjarni_version(void) { return "0.1.0+b200"; }
void kjarni_float_array_free(const KjarniFloatArray* arr) {
    if (arr && arr->data && arr->len > 0) free(arr->data);
}
void kjarni_float_2d_array_free(const KjarniFloat2DArray* arr) {
    if (arr && arr->data && arr->rows > 0 && arr->cols > 0) free(arr->data);
}
void kjarni_string_free(char* s) { free(s); }
void kjarni_string_array_free(const KjarniStringArray* arr) {
    if (!arr || !arr->strings || arr->len == 0) return;
    for (size_t i = 0; i < arr->len; ++i) free(arr->strings[i]);
    free(arr->strings);
}
float kjarni_cosine_similarity(const float* a, const float* b, size_t len) { return kjc_cosine_similarity(a, b, len); }
const char* kjarni_error_name(KjarniErrorCode err) { return kjc_error_name(static_cast<int>(err)); }
const char* kjarni_error_code_to_string(KjarniErrorCode err) { return kjc_error_name(static_cast<int>(err)); }
const char* kjarni_last_error_message(void) { return kjc_last_error_message(); }
void kjarni_clear_error(void) { kjc_clear_error(); }

#define KJ_NULLCHECK(cond)                       \
    do {                                         \
        if (cond) return KJARNI_ERROR_NULL_POINTER; \
    } while (0)
#define KJ_UTF8(p)                                                  \
    do {                                                            \
        if ((p) && !uni::valid_utf8(p)) return KJARNI_ERROR_INVALID_UTF8; \
    } while (0)

// ------------------------------------------------------------------ Embedder
KjarniEmbedderConfig kjarni_embedder_config_default(void) { return KjarniEmbedderConfig{KJARNI_DEVICE_CPU, nullptr, nullptr, nullptr, 1, 0}; }

KjarniErrorCode kjarni_embedder_new(const KjarniEmbedderConfig* config, KjarniEmbedder** out) {
    KJ_NULLCHECK(!out);
    *out = nullptr;
    const KjarniEmbedderConfig dflt = kjarni_embedder_config_default();
    const KjarniEmbedderConfig& c = config ? *config : dflt;
    KJ_UTF8(c.cache_dir);
    KJ_UTF8(c.model_name);
    KJ_UTF8(c.model_path);
    return guarded(KJARNI_ERROR_LOAD_FAILED, [&] {
        device_check(c.device);
        const std::string dir = resolve_model_dir(c.cache_dir, c.model_name, c.model_path, "minilm-l6-v2");
        std::unique_ptr<KjarniEmbedder> e(new KjarniEmbedder);
        e->m.reset(new TextModel(dir, c.model_name ? c.model_name : "minilm-l6-v2"));
        e->normalize = c.normalize != 0;
        *out = e.release();
    });
}
void kjarni_embedder_free(KjarniEmbedder* e) { delete e; }

KjarniErrorCode kjarni_embedder_encode(KjarniEmbedder* e, const char* text, KjarniFloatArray* out) {
    KJ_NULLCHECK(!e || !text || !out);
    *out = KjarniFloatArray{nullptr, 0};
    KJ_UTF8(text);
    return guarded(KJARNI_ERROR_INFERENCE_FAILED, [&] {
        std::vector<float> v;
        e->embed({std::string(text)}, v);
        out->data = dup_floats(v.data(), v.size());
        out->len = v.size();
    });
}

KjarniErrorCode kjarni_embedder_encode_batch(KjarniEmbedder* e, const char* const* texts, size_t n, KjarniFloat2DArray* out) {
    KJ_NULLCHECK(!e || !texts || !out);
    *out = KjarniFloat2DArray{nullptr, 0, 0};
    if (n == 0) return KJARNI_OK;
    std::vector<std::string> tv;
    tv.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        KJ_NULLCHECK(!texts[i]);
        KJ_UTF8(texts[i]);
        tv.emplace_back(texts[i]);
    }
    return guarded(KJARNI_ERROR_INFERENCE_FAILED, [&] {
        std::vector<float> v;
        e->embed(tv, v);
        out->data = dup_floats(v.data(), v.size());
        out->rows = n;
        out->cols = v.size() / n;
    });
}

KjarniErrorCode kjarni_embedder_similarity(KjarniEmbedder* e, const char* t1, const char* t2, float* out) {
    KJ_NULLCHECK(!e || !t1 || !t2 || !out);
    *out = 0.0f;
    KJ_UTF8(t1);
    KJ_UTF8(t2);
    return guarded(KJARNI_ERROR_INFERENCE_FAILED, [&] {
        // Embedder::similarity: embed both (one batch of two = BatchLongest padding over the pair), cosine of the two rows
        std::vector<float> v;
        e->embed({std::string(t1), std::string(t2)}, v);
        const size_t H = v.size() / 2;
        *out = kjc_cosine_similarity(v.data(), v.data() + H, H);
    });
}
size_t kjarni_embedder_dim(const KjarniEmbedder* e) { return e ? static_cast<size_t>(e->m->enc->info().hidden_size) : 0; }

// ------------------------------------------------------------------ Classifier
void kjarni_class_results_free(const KjarniClassResults* r) {
    if (!r || !r->results || r->len == 0) return;
    for (size_t i = 0; i < r->len; ++i) free(r->results[i].label);
    free(r->results);
}
KjarniClassifierConfig kjarni_classifier_config_default(void) {
    return KjarniClassifierConfig{KJARNI_DEVICE_CPU, nullptr, nullptr, nullptr, nullptr, 0, 0, 0};
}
KjarniErrorCode kjarni_classifier_new(const KjarniClassifierConfig* config, KjarniClassifier** out) {
    KJ_NULLCHECK(!out);
    *out = nullptr;
    const KjarniClassifierConfig dflt = kjarni_classifier_config_default();
    const KjarniClassifierConfig& c = config ? *config : dflt;
    KJ_UTF8(c.cache_dir);
    KJ_UTF8(c.model_name);
    KJ_UTF8(c.model_path);
    std::vector<std::string> custom;
    if (c.labels && c.num_labels > 0)
        for (size_t i = 0; i < c.num_labels; ++i) {
            KJ_NULLCHECK(!c.labels[i]);
            KJ_UTF8(c.labels[i]);
            custom.emplace_back(c.labels[i]);
        }
    return guarded(KJARNI_ERROR_LOAD_FAILED, [&] {
        device_check(c.device);
        const std::string dir = resolve_model_dir(c.cache_dir, c.model_name, c.model_path, "sentiment");
        std::unique_ptr<KjarniClassifier> k(new KjarniClassifier);
        k->m.reset(new TextModel(dir, c.model_name ? c.model_name : "sentiment"));
        const KjcEncoderInfo& info = k->m->enc->info();
        if (info.head_kind == KJC_HEAD_ABSENT) throw Fail{KJARNI_ERROR_INVALID_CONFIG, "model has no classification head"};
        if (!custom.empty()) {
            if (static_cast<int>(custom.size()) != info.num_labels)
                throw Fail{KJARNI_ERROR_INVALID_CONFIG, "Label count mismatch: model has " + std::to_string(info.num_labels) + " outputs, " +
                                                            std::to_string(custom.size()) + " labels given"};
            k->labels = custom;
        } else {
            k->labels = k->m->enc->labels();
            for (int i = static_cast<int>(k->labels.size()); i < info.num_labels; ++i) k->labels.push_back("LABEL_" + std::to_string(i));
        }
        k->multi_label = c.multi_label != 0;
        *out = k.release();
    });
}
void kjarni_classifier_free(KjarniClassifier* k) { delete k; }

KjarniErrorCode kjarni_classifier_classify(KjarniClassifier* k, const char* text, KjarniClassResults* out) {
    KJ_NULLCHECK(!k || !text || !out);
    *out = KjarniClassResults{nullptr, 0};
    KJ_UTF8(text);
    return guarded(KJARNI_ERROR_INFERENCE_FAILED, [&] {
        std::lock_guard<std::mutex> lock(k->mu);
        const KjcForwardOptions o{KJC_OUT_LOGITS, KJC_POOL_CLS, 0, KJC_MASK_ALLOC};
        const size_t C = static_cast<size_t>(k->m->enc->info().num_labels);
        std::vector<float> s;
        k->m->run({std::string(text)}, {}, o, true, s, C);  // predict_logits passes the tokenizer's type ids (sequence_classifier/mod.rs:281-296)
        if (k->multi_label) kjc_sigmoid_rows(s.data(), 1, static_cast<int>(C));
        else kjc_softmax_rows(s.data(), 1, static_cast<int>(C));
        std::vector<size_t> order(C);
        for (size_t i = 0; i < C; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return s[a] > s[b]; });  // types.rs:106-131
        KjarniClassResult* r = static_cast<KjarniClassResult*>(calloc(C, sizeof(KjarniClassResult)));
        if (!r) throw std::bad_alloc();
        for (size_t i = 0; i < C; ++i) {
            r[i].label = dup_cstr(k->labels[order[i]]);
            r[i].score = s[order[i]];
        }
        out->results = r;
        out->len = C;
    });
}
KjarniErrorCode kjarni_classifier_labels(const KjarniClassifier* k, KjarniStringArray* out) {
    KJ_NULLCHECK(!k || !out);
    *out = KjarniStringArray{nullptr, 0};
    return guarded(KJARNI_ERROR_INFERENCE_FAILED, [&] {
        if (k->labels.empty()) return;
        char** p = static_cast<char**>(calloc(k->labels.size(), sizeof(char*)));
        if (!p) throw std::bad_alloc();
        for (size_t i = 0; i < k->labels.size(); ++i) p[i] = dup_cstr(k->labels[i]);
        out->strings = p;
        out->len = k->labels.size();
    });
}
size_t kjarni_classifier_num_labels(const KjarniClassifier* k) { return k ? k->labels.size() : 0; }

// ------------------------------------------------------------------ Reranker
void kjarni_rerank_results_free(const KjarniRerankResults* r) {
    if (r && r->results && r->len > 0) free(r->results);
}
KjarniRerankerConfig kjarni_reranker_config_default(void) { return KjarniRerankerConfig{KJARNI_DEVICE_CPU, nullptr, nullptr, nullptr, 0}; }
KjarniErrorCode kjarni_reranker_new(const KjarniRerankerConfig* config, KjarniReranker** out) {
    KJ_NULLCHECK(!out);
    *out = nullptr;
    const KjarniRerankerConfig dflt = kjarni_reranker_config_default();
    const KjarniRerankerConfig& c = config ? *config : dflt;
    KJ_UTF8(c.cache_dir);
    KJ_UTF8(c.model_name);
    KJ_UTF8(c.model_path);
    return guarded(KJARNI_ERROR_LOAD_FAILED, [&] {
        device_check(c.device);
        const std::string dir = resolve_model_dir(c.cache_dir, c.model_name, c.model_path, "minilm-l6-v2-cross-encoder");
        std::unique_ptr<KjarniReranker> r(new KjarniReranker);
        r->m.reset(new TextModel(dir, c.model_name ? c.model_name : "minilm-l6-v2-cross-encoder"));
        if (r->m->enc->info().head_kind == KJC_HEAD_ABSENT) throw Fail{KJARNI_ERROR_INVALID_CONFIG, "model has no classification head"};
        *out = r.release();
    });
}
void kjarni_reranker_free(KjarniReranker* r) { delete r; }
KjarniErrorCode kjarni_reranker_score(KjarniReranker* r, const char* query, const char* document, float* out) {
    KJ_NULLCHECK(!r || !query || !document || !out);
    *out = 0.0f;
    KJ_UTF8(query);
    KJ_UTF8(document);
    return guarded(KJARNI_ERROR_INFERENCE_FAILED, [&] {
        std::vector<float> s;
        r->score(query, {std::string(document)}, s);
        *out = s[0];
    });
}
KjarniErrorCode kjarni_reranker_rerank_top_k(KjarniReranker* r, const char* query, const char* const* documents, size_t n, size_t top_k,
                                             KjarniRerankResults* out) {
    KJ_NULLCHECK(!r || !query || !documents || !out);
    *out = KjarniRerankResults{nullptr, 0};
    KJ_UTF8(query);
    if (n == 0) return KJARNI_OK;
    std::vector<std::string> docs;
    for (size_t i = 0; i < n; ++i) {
        KJ_NULLCHECK(!documents[i]);
        KJ_UTF8(documents[i]);
        docs.emplace_back(documents[i]);
    }
    return guarded(KJARNI_ERROR_INFERENCE_FAILED, [&] {
        std::vector<float> s;
        r->score(query, docs, s);
        std::vector<size_t> order(n);
        for (size_t i = 0; i < n; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return s[a] > s[b]; });  // cross_encoder/model.rs:243-255
        const size_t m = std::min(n, top_k);
        if (m == 0) return;
        KjarniRerankResult* res = static_cast<KjarniRerankResult*>(calloc(m, sizeof(KjarniRerankResult)));
        if (!res) throw std::bad_alloc();
        for (size_t i = 0; i < m; ++i) res[i] = KjarniRerankResult{order[i], s[order[i]]};
        out->results = res;
        out->len = m;
    });
}
KjarniErrorCode kjarni_reranker_rerank(KjarniReranker* r, const char* query, const char* const* documents, size_t n, KjarniRerankResults* out) {
    return kjarni_reranker_rerank_top_k(r, query, documents, n, static_cast<size_t>(-1), out);
}

// ------------------------------------------------------------------ Searcher
void kjarni_search_results_free(const KjarniSearchResults* r) {
    if (!r || !r->results || r->len == 0) return;
    for (size_t i = 0; i < r->len; ++i) {
        free(r->results[i].text);
        free(r->results[i].metadata_json);
    }
    free(r->results);
}
KjarniSearchOptions kjarni_search_options_default(void) { return KjarniSearchOptions{-1, 0, -1, 0.0f, nullptr, nullptr, nullptr}; }
KjarniSearcherConfig kjarni_searcher_config_default(void) {
    return KjarniSearcherConfig{KJARNI_DEVICE_CPU, nullptr, nullptr, nullptr, KJARNI_SEARCH_MODE_HYBRID, 10, 0};
}
KjarniErrorCode kjarni_searcher_new(const KjarniSearcherConfig* config, KjarniSearcher** out) {
    KJ_NULLCHECK(!out);
    *out = nullptr;
    const KjarniSearcherConfig dflt = kjarni_searcher_config_default();
    const KjarniSearcherConfig& c = config ? *config : dflt;
    KJ_UTF8(c.cache_dir);
    KJ_UTF8(c.model_name);
    KJ_UTF8(c.rerank_model);
    return guarded(KJARNI_ERROR_LOAD_FAILED, [&] {
        device_check(c.device);
        std::unique_ptr<KjarniSearcher> s(new KjarniSearcher);
        s->embedder.reset(new KjarniEmbedder);
        s->embedder->m.reset(new TextModel(resolve_model_dir(c.cache_dir, c.model_name, nullptr, "minilm-l6-v2"), c.model_name ? c.model_name : "minilm-l6-v2"));
        s->embedder->normalize = true;
        if (c.rerank_model && *c.rerank_model) {
            s->reranker.reset(new KjarniReranker);
            s->reranker->m.reset(new TextModel(resolve_model_dir(c.cache_dir, c.rerank_model, nullptr, "minilm-l6-v2-cross-encoder"), c.rerank_model));
        }
        s->default_mode = c.default_mode;
        s->default_top_k = c.default_top_k;
        s->model_name = (c.model_name && *c.model_name) ? c.model_name : "minilm-l6-v2";
        if (s->reranker) s->reranker_name = c.rerank_model;
        *out = s.release();
    });
}
void kjarni_searcher_free(KjarniSearcher* s) { delete s; }
bool kjarni_searcher_has_reranker(const KjarniSearcher* s) { return s && s->reranker; }
KjarniSearchMode kjarni_searcher_default_mode(const KjarniSearcher* s) { return s ? s->default_mode : KJARNI_SEARCH_MODE_HYBRID; }
size_t kjarni_searcher_default_top_k(const KjarniSearcher* s) { return s ? s->default_top_k : 0; }

KjarniErrorCode kjarni_searcher_search_with_options(KjarniSearcher* s, const char* index_path, const char* query, const KjarniSearchOptions* options,
                                                    KjarniSearchResults* out) {
    KJ_NULLCHECK(!s || !index_path || !query || !out);
    *out = KjarniSearchResults{nullptr, 0};
    KJ_UTF8(index_path);
    KJ_UTF8(query);
    const KjarniSearchOptions dflt = kjarni_search_options_default();
    const KjarniSearchOptions& o = options ? *options : dflt;
    return guarded(KJARNI_ERROR_INFERENCE_FAILED, [&] {
        // Searcher::search_with_options, kjarni/src/searcher/model.rs:96-189
        std::lock_guard<std::mutex> lock(s->mu);
        if (!file_ok(std::string(index_path) + "/config.json")) throw Fail{KJARNI_ERROR_MODEL_NOT_FOUND, std::string("Index not found: ") + index_path};
        // The reference opens the index on every call (IndexReader::open is an mmap there).  Here opening is an upload, so the
        // shard is kept -- but only while the directory is what it was: the segment list with the sizes and mtimes of the files
        // is compared on every call, and at most eight indexes stay resident (least recently used goes first).
        IndexDir now = scan_index_dir(index_path);
        const std::string fp = index_fingerprint(now);
        auto it = s->opened.find(index_path);
        if (it != s->opened.end() && it->second.fingerprint != fp) {
            s->opened.erase(it);
            it = s->opened.end();
        }
        if (it == s->opened.end()) {
            if (s->opened.size() >= 8) {
                auto victim = s->opened.begin();
                for (auto j = s->opened.begin(); j != s->opened.end(); ++j)
                    if (j->second.last_use < victim->second.last_use) victim = j;
                s->opened.erase(victim);
            }
            KjarniSearcher::Opened op;
            op.dir = std::move(now);
            op.fingerprint = fp;
            if (op.dir.total_rows > 0) {
                const std::vector<int> devs = shim_devices();
                op.idx.reset(new ShardedIndex(index_path, devs.data(), static_cast<int>(devs.size())));
            }
            it = s->opened.emplace(index_path, std::move(op)).first;
        }
        KjarniSearcher::Opened& op = it->second;
        op.last_use = ++s->use_counter;
        const int model_dim = s->embedder->m->enc->info().hidden_size;
        if (op.dir.dimension != model_dim)
            throw Fail{KJARNI_ERROR_INVALID_CONFIG, "Dimension mismatch: index has " + std::to_string(op.dir.dimension) + ", model has " + std::to_string(model_dim)};
        const int mode = o.mode >= 0 ? (o.mode == 0 ? 0 : (o.mode == 1 ? 1 : 2)) : static_cast<int>(s->default_mode);
        const size_t top_k = o.top_k > 0 ? o.top_k : s->default_top_k;
        const bool use_reranker = (o.use_reranker >= 0 ? o.use_reranker != 0 : static_cast<bool>(s->reranker)) && s->reranker;
        const size_t fetch_k = use_reranker ? top_k * 5 : top_k;
        const bool has_filter = (o.source_pattern && uni::valid_utf8(o.source_pattern)) || (o.filter_key && o.filter_value);
        const size_t limit = has_filter ? fetch_k * 3 : fetch_k;  // search_*_filtered: limit * 3 candidates, then filter, then take(limit)

        // IndexReader::search_semantic on the GPU shard: exact cosine top-`n` in global ids, (score desc, id asc)
        auto semantic = [&](size_t n) {
            std::vector<std::pair<uint64_t, float>> res;
            if (!op.idx || n == 0) return res;
            if (n > 256) throw Fail{KJARNI_ERROR_INVALID_CONFIG, "the GPU scan returns at most 256 candidates per query; top_k x rerank / filter / hybrid factors ask for " + std::to_string(n)};
            std::vector<float> q;
            s->embedder->embed({std::string(query)}, q);
            std::vector<uint64_t> ids(n);
            std::vector<float> sc(n);
            int32_t cnt = 0;
            op.idx->search_host(q.data(), 1, static_cast<int>(n), KJC_SCAN_SEGMENT, ids.data(), sc.data(), &cnt);
            for (int i = 0; i < cnt; ++i) res.emplace_back(ids[i], sc[i]);
            return res;
        };
        std::vector<std::pair<uint64_t, float>> ranked;
        if (mode == KJARNI_SEARCH_MODE_KEYWORD) ranked = keyword_search(op.dir, op.bm25, query, limit);
        else if (mode == KJARNI_SEARCH_MODE_SEMANTIC) ranked = semantic(limit);
        else  // IndexReader::search_hybrid (index_reader.rs:248-289): 2 x limit of each, reciprocal-rank fusion (hybrid.rs:3-31)
            ranked = rrf_fuse(keyword_search(op.dir, op.bm25, query, limit * 2), semantic(limit * 2), limit);
        std::vector<SearchHit> hits;
        for (const auto& r : ranked) {
            SearchHit h{r.second, r.first, {}, {}};
            fetch_doc(op.dir, r.first, h.text, h.meta);
            if (has_filter && !filter_matches(o, h)) continue;
            hits.push_back(std::move(h));
            if (hits.size() >= fetch_k) break;
        }
        if (use_reranker && !hits.empty()) {
            std::vector<std::string> docs;
            for (const SearchHit& h : hits) docs.push_back(h.text);
            std::vector<float> rs;
            s->reranker->score(query, docs, rs);
            std::vector<size_t> order(hits.size());
            for (size_t i = 0; i < order.size(); ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return rs[a] > rs[b]; });
            std::vector<SearchHit> re;
            for (size_t i = 0; i < order.size() && i < top_k; ++i) {
                re.push_back(hits[order[i]]);
                re.back().score = rs[order[i]];
            }
            hits.swap(re);
        }
        if (o.threshold > 0.0f) {
            std::vector<SearchHit> kept;
            for (SearchHit& h : hits)
                if (h.score >= o.threshold) kept.push_back(std::move(h));
            hits.swap(kept);
        }
        if (hits.size() > top_k) hits.resize(top_k);
        fill_results(hits, out);
    });
}
KjarniErrorCode kjarni_searcher_search(KjarniSearcher* s, const char* index_path, const char* query, KjarniSearchResults* out) {
    const KjarniSearchOptions o = kjarni_search_options_default();
    return kjarni_searcher_search_with_options(s, index_path, query, &o, out);
}

size_t kjarni_searcher_model_name(const KjarniSearcher* s, char* buf, size_t buf_len) { return s ? copy_name(s->model_name, buf, buf_len) : 0; }
size_t kjarni_searcher_reranker_model(const KjarniSearcher* s, char* buf, size_t buf_len) {
    return (s && s->reranker) ? copy_name(s->reranker_name, buf, buf_len) : 0;
}
KjarniErrorCode kjarni_search_keywords(const char* index_path, const char* query, size_t top_k, KjarniSearchResults* out) {
    KJ_NULLCHECK(!index_path || !query || !out);
    *out = KjarniSearchResults{nullptr, 0};
    KJ_UTF8(index_path);
    KJ_UTF8(query);
    return guarded(KJARNI_ERROR_INFERENCE_FAILED, [&] {  // Searcher::search_keywords: IndexReader::open + search_keywords, host only
        const IndexDir d = scan_index_dir(index_path);
        std::vector<std::unique_ptr<Bm25Index>> cache;
        std::vector<SearchHit> hits;
        for (const auto& r : keyword_search(d, cache, query, top_k)) {
            SearchHit h{r.second, r.first, {}, {}};
            fetch_doc(d, r.first, h.text, h.meta);
            hits.push_back(std::move(h));
        }
        fill_results(hits, out);
    });
}

// ------------------------------------------------------------------ cancellation tokens (kjarni-ffi/src/callback.rs:46-92)
KjarniCancelToken* kjarni_cancel_token_new(void) { return new (std::nothrow) KjarniCancelToken; }
void kjarni_cancel_token_cancel(KjarniCancelToken* t) { if (t) t->flag.store(true); }
bool kjarni_cancel_token_is_cancelled(const KjarniCancelToken* t) { return t && t->flag.load(); }
void kjarni_cancel_token_reset(KjarniCancelToken* t) { if (t) t->flag.store(false); }
void kjarni_cancel_token_free(KjarniCancelToken* t) { delete t; }

// ------------------------------------------------------------------ Indexer (kjarni-ffi/src/indexer.rs, kjarni/src/indexer/model.rs)
void kjarni_index_info_free(KjarniIndexInfo info) {
    free(info.path);
    free(info.embedding_model);
}
KjarniIndexerConfig kjarni_indexer_config_default(void) {
    return KjarniIndexerConfig{KJARNI_DEVICE_CPU, nullptr, nullptr, 512, 50, 32, nullptr, nullptr, 1, 0, 10u * 1024 * 1024, 0};
}
KjarniErrorCode kjarni_indexer_new(const KjarniIndexerConfig* config, KjarniIndexer** out) {
    KJ_NULLCHECK(!out);
    *out = nullptr;
    const KjarniIndexerConfig dflt = kjarni_indexer_config_default();
    const KjarniIndexerConfig& c = config ? *config : dflt;
    KJ_UTF8(c.cache_dir);
    KJ_UTF8(c.model_name);
    KJ_UTF8(c.extensions);
    KJ_UTF8(c.exclude_patterns);
    return guarded(KJARNI_ERROR_LOAD_FAILED, [&] {
        device_check(c.device);
        if (c.chunk_size == 0) throw Fail{KJARNI_ERROR_INVALID_CONFIG, "chunk_size must be greater than 0"};               // SplitterConfig::validate
        if (c.chunk_overlap >= c.chunk_size) throw Fail{KJARNI_ERROR_INVALID_CONFIG, "chunk_overlap must be less than chunk_size"};
        std::unique_ptr<KjarniIndexer> ix(new KjarniIndexer);
        ix->model_name = (c.model_name && *c.model_name) ? c.model_name : "minilm-l6-v2";
        ix->embedder.reset(new KjarniEmbedder);
        ix->embedder->m.reset(new TextModel(resolve_model_dir(c.cache_dir, c.model_name, nullptr, "minilm-l6-v2"), ix->model_name));
        ix->embedder->normalize = true;
        ix->splitter.chunk_size = c.chunk_size;
        ix->splitter.chunk_overlap = c.chunk_overlap;
        ix->batch_size = std::max<size_t>(c.batch_size, 1);
        ix->loader.extensions = split_csv(c.extensions, true);
        ix->loader.exclude_patterns = split_csv(c.exclude_patterns, false);
        ix->loader.recursive = c.recursive != 0;
        ix->loader.include_hidden = c.include_hidden != 0;
        if (c.max_file_size > 0) ix->loader.max_file_size = c.max_file_size;
        ix->quiet = c.quiet != 0;
        *out = ix.release();
    });
}
void kjarni_indexer_free(KjarniIndexer* ix) { delete ix; }
size_t kjarni_indexer_model_name(const KjarniIndexer* ix, char* buf, size_t buf_len) { return ix ? copy_name(ix->model_name, buf, buf_len) : 0; }
size_t kjarni_indexer_dimension(const KjarniIndexer* ix) { return ix ? static_cast<size_t>(ix->embedder->m->enc->info().hidden_size) : 0; }
size_t kjarni_indexer_chunk_size(const KjarniIndexer* ix) { return ix ? ix->splitter.chunk_size : 0; }

}  // extern "C"

namespace {
// Indexer::create_internal / add_internal (kjarni/src/indexer/model.rs:169-300,466-570): discover -> load + split -> embed -> write.
// Chunks are embedded in GPU-sized groups (at least the caller's batch_size, up to 1024 texts: rows are independent, so the
// grouping is invisible in the result) and every embedding row goes from the encoder's pinned output buffer straight into the
// segment's vectors.bin.
size_t run_indexer(KjarniIndexer* ix, IndexWriter& writer, const std::vector<std::string>& inputs, KjarniProgressCallbackFn cb, void* user,
                   const KjarniCancelToken* cancel, KjarniIndexStats* stats) {
    auto report = [&](KjarniProgressStage st, size_t cur, size_t total, const char* msg) {
        if (cb) cb(KjarniProgress{st, cur, total, msg}, user);
    };
    auto check_cancel = [&] {
        if (cancel && cancel->flag.load()) throw Fail{KJARNI_ERROR_CANCELLED, "Operation cancelled"};
    };
    report(KJARNI_PROGRESS_SCANNING, 0, 0, "Discovering files...");
    std::vector<std::string> files;
    try {
        files = collect_files(inputs, ix->loader);
    } catch (const Error& e) {
        throw Fail{KJARNI_ERROR_MODEL_NOT_FOUND, e.what()};
    }
    check_cancel();
    const size_t group = std::max<size_t>(ix->batch_size, 1024);
    const int dim = writer.dimension();
    std::vector<std::string> texts;
    std::vector<std::vector<std::pair<std::string, std::string>>> metas;
    size_t total_docs = 0, total_chunks = 0, processed = 0, skipped = 0;
    auto flush = [&] {
        size_t done = 0;
        while (done < texts.size()) {
            check_cancel();
            report(KJARNI_PROGRESS_EMBEDDING, total_docs, 0, nullptr);
            SegmentWriter& seg = writer.current();
            const size_t n = std::min(texts.size() - done, seg.room());
            std::vector<std::string> part(texts.begin() + done, texts.begin() + done + n);
            std::vector<size_t> row_id(n);
            for (size_t i = 0; i < n; ++i) row_id[i] = seg.add_text(part[i], metas[done + i]);
            const TextModel::TextSink sink = [&](const float* rows, size_t cnt, const size_t* index) {
                size_t i = 0;
                while (i < cnt) {  // runs of consecutive documents go out as one write
                    size_t j = i + 1;
                    while (j < cnt && index[j] == index[j - 1] + 1) ++j;
                    seg.write_rows(row_id[index[i]], rows + i * dim, j - i);
                    i = j;
                }
            };
            {
                std::lock_guard<std::mutex> lock(ix->embedder->mu);
                const KjcForwardOptions o{KJC_OUT_POOLED, KJC_POOL_MEAN, 1, KJC_MASK_AUTO};
                std::vector<float> unused;
                ix->embedder->m->run(part, {}, o, false, unused, static_cast<size_t>(dim), &sink);
            }
            writer.note_added(n);
            total_docs += n;
            done += n;
        }
        texts.clear();
        metas.clear();
    };
    for (size_t fi = 0; fi < files.size(); ++fi) {
        check_cancel();
        report(KJARNI_PROGRESS_LOADING, fi, files.size(), files[fi].c_str());
        std::string content;
        bool ok = true;
        try {
            content = read_text_file(files[fi], KJC_LOAD_FAILED);
            ok = uni::valid_utf8(content.c_str()) && content.find('\0') == std::string::npos;  // fs::read_to_string fails on invalid UTF-8
        } catch (const Error&) {
            ok = false;
        }
        if (!ok) {
            ++skipped;
            if (!ix->quiet) fprintf(stderr, "Warning: Failed to load %s\n", files[fi].c_str());
            continue;
        }
        const std::vector<std::string> chunks = ix->splitter.split(content);  // DocumentLoader::load_file, loader.rs:71-93
        total_chunks += chunks.size();
        ++processed;
        for (size_t ci = 0; ci < chunks.size(); ++ci) {
            texts.push_back(chunks[ci]);
            metas.push_back({{"source", files[fi]}, {"chunk_index", std::to_string(ci)}, {"total_chunks", std::to_string(chunks.size())}});
            if (texts.size() >= group) flush();
        }
    }
    if (!texts.empty()) flush();
    report(KJARNI_PROGRESS_COMMITTING, total_docs, total_docs, "Finalizing index...");
    writer.commit();
    if (stats) {
        stats->documents_indexed = total_docs;
        stats->chunks_created = total_chunks;
        stats->dimension = static_cast<size_t>(dim);
        stats->files_processed = processed;
        stats->files_skipped = skipped;
    }
    return total_docs;
}

bool parse_inputs(const char* const* inputs, size_t n, std::vector<std::string>& out, KjarniErrorCode& err) {
    for (size_t i = 0; i < n; ++i) {
        if (!inputs[i]) { err = KJARNI_ERROR_NULL_POINTER; return false; }
        if (!uni::valid_utf8(inputs[i])) { err = KJARNI_ERROR_INVALID_UTF8; return false; }
        out.emplace_back(inputs[i]);
    }
    return true;
}
bool path_exists(const std::string& p) {
    struct stat sb;
    return stat(p.c_str(), &sb) == 0;
}
}  // namespace

extern "C" {

KjarniErrorCode kjarni_indexer_create_with_callback(KjarniIndexer* ix, const char* index_path, const char* const* inputs, size_t num_inputs, int32_t force,
                                                    KjarniProgressCallbackFn cb, void* user_data, const KjarniCancelToken* cancel, KjarniIndexStats* out) {
    KJ_NULLCHECK(!ix || !index_path || !inputs || !out);
    *out = KjarniIndexStats{};
    KJ_UTF8(index_path);
    std::vector<std::string> in;
    KjarniErrorCode perr = KJARNI_OK;
    if (!parse_inputs(inputs, num_inputs, in, perr)) return perr;
    return guarded(KJARNI_ERROR_INFERENCE_FAILED, [&] {
        std::lock_guard<std::mutex> lock(ix->mu);
        const auto t0 = std::chrono::steady_clock::now();
        if (in.empty()) throw Fail{KJARNI_ERROR_INVALID_CONFIG, "No input paths specified"};
        if (path_exists(index_path)) {
            if (!force) throw Fail{KJARNI_ERROR_INVALID_CONFIG, std::string("Index already exists at ") + index_path + ". Use force=true to overwrite."};
            remove_tree(index_path);
        }
        IndexWriter writer(index_path, true, ix->embedder->m->enc->info().hidden_size, ix->max_docs_per_segment, ix->model_name);
        run_indexer(ix, writer, in, cb, user_data, cancel, out);
        out->size_bytes = tree_size(index_path);
        out->elapsed_ms = static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count());
    });
}
KjarniErrorCode kjarni_indexer_create(KjarniIndexer* ix, const char* index_path, const char* const* inputs, size_t num_inputs, int32_t force,
                                      KjarniIndexStats* out) {
    return kjarni_indexer_create_with_callback(ix, index_path, inputs, num_inputs, force, nullptr, nullptr, nullptr, out);
}
KjarniErrorCode kjarni_indexer_add_with_callback(KjarniIndexer* ix, const char* index_path, const char* const* inputs, size_t num_inputs,
                                                 KjarniProgressCallbackFn cb, void* user_data, const KjarniCancelToken* cancel, size_t* documents_added) {
    KJ_NULLCHECK(!ix || !index_path || !inputs || !documents_added);
    *documents_added = 0;
    if (num_inputs == 0) return KJARNI_OK;
    KJ_UTF8(index_path);
    std::vector<std::string> in;
    KjarniErrorCode perr = KJARNI_OK;
    if (!parse_inputs(inputs, num_inputs, in, perr)) return perr;
    return guarded(KJARNI_ERROR_INFERENCE_FAILED, [&] {
        std::lock_guard<std::mutex> lock(ix->mu);
        if (!path_exists(index_path)) throw Fail{KJARNI_ERROR_MODEL_NOT_FOUND, std::string("Index not found at ") + index_path};
        IndexWriter writer(index_path, false, 0, 0, "");
        const int model_dim = ix->embedder->m->enc->info().hidden_size;
        if (writer.dimension() != model_dim)
            throw Fail{KJARNI_ERROR_INVALID_CONFIG, "Dimension mismatch: index has " + std::to_string(writer.dimension()) + ", model produces " + std::to_string(model_dim)};
        *documents_added = run_indexer(ix, writer, in, cb, user_data, cancel, nullptr);
    });
}
KjarniErrorCode kjarni_indexer_add(KjarniIndexer* ix, const char* index_path, const char* const* inputs, size_t num_inputs, size_t* documents_added) {
    return kjarni_indexer_add_with_callback(ix, index_path, inputs, num_inputs, nullptr, nullptr, nullptr, documents_added);
}
KjarniErrorCode kjarni_index_info(const char* index_path, KjarniIndexInfo* out) {
    KJ_NULLCHECK(!index_path || !out);
    *out = KjarniIndexInfo{};
    KJ_UTF8(index_path);
    return guarded(KJARNI_ERROR_INFERENCE_FAILED, [&] {  // Indexer::info, kjarni/src/indexer/model.rs:95-124
        if (!path_exists(index_path)) throw Fail{KJARNI_ERROR_MODEL_NOT_FOUND, std::string("Index not found at ") + index_path};
        const IndexDir d = scan_index_dir(index_path);
        std::string model;
        {
            const std::string txt = read_text_file(std::string(index_path) + "/config.json", KJC_MODEL_NOT_FOUND);
            const Json cfg = JsonParser(txt.data(), txt.size()).parse();
            model = cfg.string("embedding_model", "");
        }
        out->path = dup_cstr(index_path);
        out->document_count = static_cast<size_t>(d.total_rows);
        out->segment_count = d.segments.size();
        out->dimension = static_cast<size_t>(d.dimension);
        out->size_bytes = tree_size(index_path);
        out->embedding_model = model.empty() ? nullptr : dup_cstr(model);
    });
}
KjarniErrorCode kjarni_index_delete(const char* index_path) {
    KJ_NULLCHECK(!index_path);
    KJ_UTF8(index_path);
    return guarded(KJARNI_ERROR_INFERENCE_FAILED, [&] {
        if (!path_exists(index_path)) throw Fail{KJARNI_ERROR_MODEL_NOT_FOUND, std::string("Index not found at ") + index_path};
        remove_tree(index_path);
    });
}

// ------------------------------------------------------------------ Chat: out of scope (LLM decode), load-compatibility stubs
KjarniChatConfig kjarni_chat_config_default(void) { return KjarniChatConfig{KJARNI_DEVICE_CPU, nullptr, nullptr, nullptr, nullptr, 0, 0}; }
KjarniGenerationConfig kjarni_generation_config_default(void) { return KjarniGenerationConfig{-1.0f, -1, -1.0f, -1.0f, -1.0f, -1, -1}; }
static KjarniErrorCode chat_unsupported() {
    set_last_error("chat / text generation is not served by the CUDA encoder backend (autoregressive decode is outside its scope)");
    return KJARNI_ERROR_INVALID_CONFIG;
}
KjarniErrorCode kjarni_chat_new(const KjarniChatConfig*, KjarniChat** out) {
    KJ_NULLCHECK(!out);
    *out = nullptr;
    return chat_unsupported();
}
void kjarni_chat_free(KjarniChat*) {}
KjarniErrorCode kjarni_chat_send(KjarniChat*, const char*, const KjarniGenerationConfig*, char** out) {
    if (out) *out = nullptr;
    return chat_unsupported();
}
KjarniErrorCode kjarni_chat_stream(KjarniChat*, const char*, const KjarniGenerationConfig*, KjarniStreamCallbackFn, void*, const KjarniCancelToken*) {
    return chat_unsupported();
}
KjarniErrorCode kjarni_chat_send_with_history(KjarniChat*, const int32_t*, const char* const*, size_t, const char*, const KjarniGenerationConfig*, char** out) {
    if (out) *out = nullptr;
    return chat_unsupported();
}
KjarniErrorCode kjarni_chat_conversation_new(KjarniChat*, KjarniChatConversation** out) {
    if (out) *out = nullptr;
    return chat_unsupported();
}
void kjarni_chat_conversation_free(KjarniChatConversation*) {}
KjarniErrorCode kjarni_chat_conversation_send(KjarniChatConversation*, const char*, const KjarniGenerationConfig*, char** out) {
    if (out) *out = nullptr;
    return chat_unsupported();
}
KjarniErrorCode kjarni_chat_conversation_stream(KjarniChatConversation*, const char*, const KjarniGenerationConfig*, KjarniStreamCallbackFn, void*,
                                                const KjarniCancelToken*) {
    return chat_unsupported();
}
size_t kjarni_chat_conversation_len(const KjarniChatConversation*) { return 0; }
void kjarni_chat_conversation_clear(KjarniChatConversation*, int32_t) {}
size_t kjarni_chat_model_name(const KjarniChat*, char*, size_t) { return 0; }
size_t kjarni_chat_context_size(const KjarniChat*) { return 0; }

// ------------------------------------------------------------------ tokenizer (inner ABI, include/kjarni_cuda.h)
struct KjcTokenizer {
    Tokenizer impl;
    KjcTokenizer(const char* p, int max_len) : impl(p, max_len) {}
};
int kjc_tokenizer_create(const char* tokenizer_json_path, int max_length, KjcTokenizer** out) {
    if (!tokenizer_json_path || !out) { set_last_error("null pointer argument"); return KJC_NULL_POINTER; }
    *out = nullptr;
    return guarded(KJARNI_ERROR_LOAD_FAILED, [&] { *out = new KjcTokenizer(tokenizer_json_path, max_length); });
}
void kjc_tokenizer_destroy(KjcTokenizer* t) { delete t; }
int kjc_tokenizer_token_to_id(const KjcTokenizer* t, const char* token, uint32_t* out_id) {
    if (!t || !token || !out_id) return 0;
    return t->impl.token_to_id(token, *out_id) ? 1 : 0;
}
int kjc_tokenizer_encode_batch(const KjcTokenizer* t, const char* const* texts, const char* const* pairs, int n, int add_special_tokens,
                               uint32_t* ids, float* mask, uint32_t* type_ids, int cap_seq, int* out_seq_len) {
    if (!t || !texts || !out_seq_len || n < 0) { set_last_error("null pointer argument"); return KJC_NULL_POINTER; }
    std::vector<std::string> a, b;
    for (int i = 0; i < n; ++i) {
        if (!texts[i] || (pairs && !pairs[i])) { set_last_error("null pointer argument"); return KJC_NULL_POINTER; }
        if (!uni::valid_utf8(texts[i]) || (pairs && !uni::valid_utf8(pairs[i]))) { set_last_error("invalid UTF-8"); return KJC_INVALID_UTF8; }
        a.emplace_back(texts[i]);
        if (pairs) b.emplace_back(pairs[i]);
    }
    return guarded(KJARNI_ERROR_INFERENCE_FAILED, [&] {
        std::vector<uint32_t> vi, vt;
        std::vector<float> vm;
        int S = 0;
        t->impl.encode_batch(a, b, add_special_tokens != 0, vi, vm, vt, S);
        *out_seq_len = S;
        if (!ids) return;  // query call: sequence length only
        if (S > cap_seq) throw Error(KJC_INVALID_CONFIG, "kjc_tokenizer_encode_batch: buffers hold " + std::to_string(cap_seq) + " tokens per row, need " + std::to_string(S));
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < cap_seq; ++k) {
                const bool in = k < S;
                ids[static_cast<size_t>(i) * cap_seq + k] = in ? vi[static_cast<size_t>(i) * S + k] : 0u;
                if (mask) mask[static_cast<size_t>(i) * cap_seq + k] = in ? vm[static_cast<size_t>(i) * S + k] : 0.f;
                if (type_ids) type_ids[static_cast<size_t>(i) * cap_seq + k] = in ? vt[static_cast<size_t>(i) * S + k] : 0u;
            }
    }, true);
}

}  // extern "C"
