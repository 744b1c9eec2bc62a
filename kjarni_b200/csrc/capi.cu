// extern "C" surface of libkjarni_cuda.so (include/kjarni_cuda.h): argument checks,
// exception -> status-code translation and the thread-local error message
// (same contract as kjarni-ffi/src/error.rs:7-101 in the reference).
#include <cmath>

#include "multi.hpp"
#include "../../include/kjarni_cuda_debug.h"

namespace kj {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
}  // namespace kj

struct KjcEncoder {
    kj::EncoderGroup grp;  // one replica (kjc_encoder_create) or one per device (kjc_encoder_create_multi)
    kj::Encoder& impl;     // replica 0: info, labels, device-pointer entry, profiling, debug hooks
    KjcEncoder(const char* d, const int* devs, int n) : grp(d, devs, n), impl(grp.replica(0)) {}
};
struct KjcShardedIndex {
    kj::ShardedIndex impl;
    KjcShardedIndex(int dim, uint64_t cap, const int* devs, int n) : impl(dim, cap, devs, n) {}
    KjcShardedIndex(const char* root, const int* devs, int n) : impl(std::string(root), devs, n) {}
};
struct KjcIndex {
    std::unique_ptr<kj::Index> own;
    kj::Index& impl;
    KjcIndex(int dim, uint64_t cap, uint64_t base, int dev) : own(new kj::Index(dim, cap, base, dev)), impl(*own) {}
    explicit KjcIndex(kj::Index* p) : own(p), impl(*own) {}
};

template <typename F>
static int guarded(F&& f) {
    try {
        f();
        return KJC_OK;
    } catch (const kj::Error& e) {
        kj::set_last_error(e.what());
        return e.status;
    } catch (const std::bad_alloc&) {
        kj::set_last_error("out of host memory");
        return KJC_INFERENCE_FAILED;
    } catch (const std::exception& e) {
        kj::set_last_error(e.what());
        return KJC_UNKNOWN;
    } catch (...) {
        kj::set_last_error("unknown error");
        return KJC_UNKNOWN;
    }
}
#define KJC_REQUIRE(ptr)                                             \
    do {                                                             \
        if ((ptr) == nullptr) {                                      \
            kj::set_last_error("null pointer argument: " #ptr);      \
            return KJC_NULL_POINTER;                                 \
        }                                                            \
    } while (0)

extern "C" {

const char* kjc_last_error_message(void) { return kj::g_last_error.empty() ? nullptr : kj::g_last_error.c_str(); }
void kjc_clear_error(void) { kj::g_last_error.clear(); }
const char* kjc_error_name(int s) {
    switch (s) {
        case KJC_OK: return "Ok";
        case KJC_NULL_POINTER: return "NullPointer";
        case KJC_INVALID_UTF8: return "InvalidUtf8";
        case KJC_MODEL_NOT_FOUND: return "ModelNotFound";
        case KJC_LOAD_FAILED: return "LoadFailed";
        case KJC_INFERENCE_FAILED: return "InferenceFailed";
        case KJC_GPU_UNAVAILABLE: return "GpuUnavailable";
        case KJC_INVALID_CONFIG: return "InvalidConfig";
        case KJC_CANCELLED: return "Cancelled";
        case KJC_TIMEOUT: return "Timeout";
        case KJC_STREAM_ENDED: return "StreamEnded";
        default: return "Unknown";
    }
}
const char* kjc_version(void) { return "kjarni-b200 0.1.0 (sm_100a)"; }
int kjc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return -1; }
    return n;
}

// ------------------------------------------------------------------ encoder
int kjc_encoder_create(const char* model_dir, int device, KjcEncoder** out) {
    KJC_REQUIRE(out);
    *out = nullptr;
    KJC_REQUIRE(model_dir);
    return guarded([&] { *out = new KjcEncoder(model_dir, &device, 1); });
}
int kjc_encoder_create_multi(const char* model_dir, const int* device_ids, int n_devices, KjcEncoder** out) {
    KJC_REQUIRE(out);
    *out = nullptr;
    KJC_REQUIRE(model_dir);
    KJC_REQUIRE(device_ids);
    return guarded([&] { *out = new KjcEncoder(model_dir, device_ids, n_devices); });
}
int kjc_encoder_device_count(const KjcEncoder* enc) { return enc ? enc->grp.size() : 0; }
void kjc_encoder_destroy(KjcEncoder* enc) { delete enc; }
int kjc_encoder_info(const KjcEncoder* enc, KjcEncoderInfo* out) {
    KJC_REQUIRE(enc);
    KJC_REQUIRE(out);
    *out = enc->impl.info();
    return KJC_OK;
}
const char* kjc_encoder_label(const KjcEncoder* enc, int i) {
    if (!enc || i < 0 || i >= static_cast<int>(enc->impl.labels().size())) return nullptr;
    return enc->impl.labels()[i].c_str();
}
static KjcForwardOptions default_opts() { return KjcForwardOptions{KJC_OUT_POOLED, KJC_POOL_MEAN, 1, KJC_MASK_AUTO}; }
int kjc_encoder_forward(KjcEncoder* enc, const uint32_t* ids, const float* mask, const uint32_t* type_ids, int batch, int seq_len,
                        const KjcForwardOptions* opts, float* out) {
    KJC_REQUIRE(enc);
    KJC_REQUIRE(ids);
    KJC_REQUIRE(out);
    const KjcForwardOptions o = opts ? *opts : default_opts();
    return guarded([&] { enc->grp.forward_host(ids, mask, type_ids, batch, seq_len, o, out); });
}
int kjc_encoder_forward_device_async(KjcEncoder* enc, const uint32_t* d_ids, const float* d_mask, const uint32_t* d_type_ids, int batch,
                                     int seq_len, const KjcForwardOptions* opts, float* d_out, void* stream) {
    KJC_REQUIRE(enc);
    KJC_REQUIRE(d_ids);
    KJC_REQUIRE(d_out);
    const KjcForwardOptions o = opts ? *opts : default_opts();
    return guarded([&] { enc->impl.forward_device(d_ids, d_mask, d_type_ids, batch, seq_len, o, d_out, static_cast<cudaStream_t>(stream)); });
}
int kjc_encoder_chained(const KjcEncoder* enc) { return enc && enc->impl.chained() ? 1 : 0; }
int kjc_encoder_set_fp32_residual(KjcEncoder* enc, int mode) {
    KJC_REQUIRE(enc);
    for (int p = 0; p < enc->grp.size(); ++p) enc->grp.replica(p).set_fp32_residual(mode);
    return KJC_OK;
}
int kjc_encoder_micro_batch(const KjcEncoder* enc, int seq_len) { return enc ? enc->impl.micro_batch(seq_len) : 0; }
int64_t kjc_encoder_last_launch_count(const KjcEncoder* enc) { return enc ? enc->grp.last_launches() : 0; }

int kjc_encoder_set_profiling(KjcEncoder* enc, int on) {
    KJC_REQUIRE(enc);
    return guarded([&] { enc->impl.set_profiling(on != 0); });
}
int kjc_encoder_get_profile(KjcEncoder* enc, double* ms, int64_t* launches) {
    KJC_REQUIRE(enc);
    KJC_REQUIRE(ms);
    KJC_REQUIRE(launches);
    return guarded([&] { enc->impl.get_profile(ms, launches); });
}

void kjc_softmax_rows(float* x, int rows, int cols) {
    // softmax_inplace, KT/activations.rs:223-242: max-subtract, exp, divide only if the sum is > 0
    if (!x) return;
    for (int r = 0; r < rows; ++r) {
        float* p = x + static_cast<size_t>(r) * cols;
        float m = -INFINITY;
        for (int c = 0; c < cols; ++c) m = fmaxf(m, p[c]);
        float s = 0.f;
        for (int c = 0; c < cols; ++c) { p[c] = expf(p[c] - m); s += p[c]; }
        if (s > 0.f) for (int c = 0; c < cols; ++c) p[c] /= s;
    }
}

void kjc_sigmoid_rows(float* x, int rows, int cols) {
    // multi-label scores: sigmoid(x) = 1 / (1 + exp(-x)) per logit (kjarni/src/classifier/model.rs:313-335,528-531)
    if (!x) return;
    const size_t n = static_cast<size_t>(rows) * cols;
    for (size_t i = 0; i < n; ++i) x[i] = 1.0f / (1.0f + expf(-x[i]));
}

// -------------------------------------------------------------------- index
int kjc_index_create(int dim, uint64_t capacity_rows, uint64_t id_base, int device, KjcIndex** out) {
    KJC_REQUIRE(out);
    *out = nullptr;
    return guarded([&] { *out = new KjcIndex(dim, capacity_rows, id_base, device); });
}
void kjc_index_destroy(KjcIndex* idx) { delete idx; }
uint64_t kjc_index_len(const KjcIndex* idx) { return idx ? idx->impl.len() : 0; }
int kjc_index_dim(const KjcIndex* idx) { return idx ? idx->impl.dim() : 0; }
int kjc_index_dir_info(const char* root, KjcIndexDirInfo* out) {
    KJC_REQUIRE(root);
    KJC_REQUIRE(out);
    memset(out, 0, sizeof *out);
    return guarded([&] {
        const kj::IndexDir d = kj::scan_index_dir(root);
        out->dimension = d.dimension;
        out->n_segments = static_cast<int32_t>(d.segments.size());
        out->n_skipped = d.skipped;
        out->total_rows = d.total_rows;
        out->max_docs_per_segment = d.max_docs_per_segment;
    });
}
int kjc_index_dir_segment_lens(const char* root, uint64_t* out_lens, int cap) {
    if (!root || (cap > 0 && !out_lens)) {
        kj::set_last_error("null pointer argument");
        return -1;
    }
    int n = -1;
    guarded([&] {
        const kj::IndexDir d = kj::scan_index_dir(root);
        for (size_t i = 0; i < d.segments.size() && static_cast<int>(i) < cap; ++i) out_lens[i] = d.segments[i].doc_count;
        n = static_cast<int>(d.segments.size());
    });
    return n;
}
int kjc_index_part_range(uint64_t total_rows, int part, int parts, uint64_t* lo, uint64_t* hi) {
    KJC_REQUIRE(lo);
    KJC_REQUIRE(hi);
    return guarded([&] { kj::index_part_range(total_rows, part, parts, lo, hi); });
}
int kjc_index_open_dir(const char* root, int device, int part, int parts, KjcIndex** out) {
    KJC_REQUIRE(root);
    KJC_REQUIRE(out);
    *out = nullptr;
    return guarded([&] {
        std::unique_ptr<kj::Index> p(kj::open_index_dir(root, device, part, parts));
        *out = new KjcIndex(p.get());
        p.release();
    });
}
uint64_t kjc_index_id_base(const KjcIndex* idx) { return idx ? idx->impl.id_base() : 0; }
int kjc_index_add_rows(KjcIndex* idx, const float* rows, uint64_t n) {
    KJC_REQUIRE(idx);
    if (n == 0) return KJC_OK;
    KJC_REQUIRE(rows);
    return guarded([&] { idx->impl.add_rows_host(rows, n); });
}
int kjc_index_load_vectors_bin(KjcIndex* idx, const char* path) {
    KJC_REQUIRE(idx);
    KJC_REQUIRE(path);
    return guarded([&] { idx->impl.load_vectors_bin(path); });
}
int kjc_index_append_synthetic(KjcIndex* idx, uint32_t seed, uint64_t row0, uint64_t n) {
    KJC_REQUIRE(idx);
    return guarded([&] { idx->impl.append_synthetic(seed, row0, n); });
}
int kjc_index_get_rows(const KjcIndex* idx, uint64_t row, uint64_t n, float* out) {
    KJC_REQUIRE(idx);
    KJC_REQUIRE(out);
    return guarded([&] { idx->impl.get_rows(row, n, out); });
}
int kjc_index_search(KjcIndex* idx, const float* queries, int nq, int k, int mode, uint64_t* out_ids, float* out_scores,
                     int32_t* out_counts) {
    KJC_REQUIRE(idx);
    KJC_REQUIRE(queries);
    KJC_REQUIRE(out_ids);
    KJC_REQUIRE(out_scores);
    return guarded([&] { idx->impl.search_host(queries, nq, k, mode, out_ids, out_scores, out_counts); });
}
int kjc_index_search_device_async(KjcIndex* idx, const float* d_queries, int nq, int k, int mode, uint64_t* d_out_ids, float* d_out_scores,
                                  int32_t* d_out_counts, void* stream) {
    KJC_REQUIRE(idx);
    KJC_REQUIRE(d_queries);
    KJC_REQUIRE(d_out_ids);
    KJC_REQUIRE(d_out_scores);
    return guarded([&] { idx->impl.search_device(d_queries, nq, k, mode, d_out_ids, d_out_scores, d_out_counts, static_cast<cudaStream_t>(stream)); });
}
int kjc_index_search_device(KjcIndex* idx, const float* d_queries, int nq, int k, int mode, uint64_t* d_out_ids, float* d_out_scores,
                            int32_t* d_out_counts, void* stream) {
    KJC_REQUIRE(idx);
    KJC_REQUIRE(d_queries);
    KJC_REQUIRE(d_out_ids);
    KJC_REQUIRE(d_out_scores);
    return guarded([&] {
        idx->impl.search_device(d_queries, nq, k, mode, d_out_ids, d_out_scores, d_out_counts, static_cast<cudaStream_t>(stream), /*may_sync=*/true);
    });
}
int kjc_topk_merge_device_async(int device, const uint64_t* d_cand_ids, const float* d_cand_scores, int n_lists, int nq, int k,
                                uint64_t* d_out_ids, float* d_out_scores, int32_t* d_out_counts, void* stream) {
    KJC_REQUIRE(d_cand_ids);
    KJC_REQUIRE(d_cand_scores);
    KJC_REQUIRE(d_out_ids);
    KJC_REQUIRE(d_out_scores);
    return guarded([&] {
        if (n_lists < 1 || nq < 1 || k < 1) throw kj::Error(KJC_INVALID_CONFIG, "n_lists, nq and k must be positive");
        KJ_CUDA(cudaSetDevice(device));
        kj::merge_lists_u64(d_cand_ids, d_cand_scores, n_lists, nq, k, d_out_ids, d_out_scores, d_out_counts, static_cast<cudaStream_t>(stream));
    });
}
// ------------------------------------------------------------ sharded index
int kjc_sharded_index_create(int dim, uint64_t capacity_rows, const int* device_ids, int n_devices, KjcShardedIndex** out) {
    KJC_REQUIRE(out);
    *out = nullptr;
    KJC_REQUIRE(device_ids);
    return guarded([&] { *out = new KjcShardedIndex(dim, capacity_rows, device_ids, n_devices); });
}
int kjc_sharded_index_open_dir(const char* root, const int* device_ids, int n_devices, KjcShardedIndex** out) {
    KJC_REQUIRE(out);
    *out = nullptr;
    KJC_REQUIRE(root);
    KJC_REQUIRE(device_ids);
    return guarded([&] { *out = new KjcShardedIndex(root, device_ids, n_devices); });
}
void kjc_sharded_index_destroy(KjcShardedIndex* idx) { delete idx; }
uint64_t kjc_sharded_index_len(const KjcShardedIndex* idx) { return idx ? idx->impl.len() : 0; }
int kjc_sharded_index_dim(const KjcShardedIndex* idx) { return idx ? idx->impl.dim() : 0; }
int kjc_sharded_index_shards(const KjcShardedIndex* idx) { return idx ? idx->impl.n_shards() : 0; }
uint64_t kjc_sharded_index_shard_len(const KjcShardedIndex* idx, int shard) {
    if (!idx || shard < 0 || shard >= idx->impl.n_shards()) return 0;
    return const_cast<KjcShardedIndex*>(idx)->impl.shard(shard).len();
}
int kjc_sharded_index_add_rows(KjcShardedIndex* idx, const float* rows, uint64_t n) {
    KJC_REQUIRE(idx);
    if (n == 0) return KJC_OK;
    KJC_REQUIRE(rows);
    return guarded([&] { idx->impl.add_rows_host(rows, n); });
}
int kjc_sharded_index_append_synthetic(KjcShardedIndex* idx, uint32_t seed, uint64_t n) {
    KJC_REQUIRE(idx);
    return guarded([&] { idx->impl.append_synthetic(seed, n); });
}
int kjc_sharded_index_search(KjcShardedIndex* idx, const float* queries, int nq, int k, int mode, uint64_t* out_ids, float* out_scores,
                             int32_t* out_counts) {
    KJC_REQUIRE(idx);
    KJC_REQUIRE(queries);
    KJC_REQUIRE(out_ids);
    KJC_REQUIRE(out_scores);
    return guarded([&] { idx->impl.search_host(queries, nq, k, mode, out_ids, out_scores, out_counts); });
}
int64_t kjc_sharded_index_last_launch_count(const KjcShardedIndex* idx) {
    if (!idx) return 0;
    int64_t s = 0;
    KjcShardedIndex* m = const_cast<KjcShardedIndex*>(idx);
    for (int p = 0; p < m->impl.n_shards(); ++p) s += m->impl.shard(p).last_launches();
    return s + (m->impl.n_shards() > 1 ? 1 : 0);
}
size_t kjc_packed_record_bytes(int nq, int k) { return kj::packed_record_bytes(nq, k); }
int kjc_topk_merge_packed_device_async(int device, const void* d_packed, int n_lists, int nq, int k, uint64_t* d_out_ids, float* d_out_scores,
                                       int32_t* d_out_counts, void* stream) {
    KJC_REQUIRE(d_packed);
    KJC_REQUIRE(d_out_ids);
    KJC_REQUIRE(d_out_scores);
    return guarded([&] {
        if (n_lists < 1 || nq < 1 || k < 1) throw kj::Error(KJC_INVALID_CONFIG, "n_lists, nq and k must be positive");
        KJ_CUDA(cudaSetDevice(device));
        // record of one list: nq*k u64 ids, then nq*k f32 scores, padded to a multiple of 16 bytes; strides in elements of each array
        const size_t rec = kj::packed_record_bytes(nq, k);
        const uint8_t* base = static_cast<const uint8_t*>(d_packed);
        kj::merge_lists_u64(reinterpret_cast<const uint64_t*>(base), reinterpret_cast<const float*>(base + static_cast<size_t>(nq) * k * 8), n_lists, nq, k,
                            d_out_ids, d_out_scores, d_out_counts, static_cast<cudaStream_t>(stream), rec / 8, rec / 4);
    });
}
int64_t kjc_index_last_launch_count(const KjcIndex* idx) { return idx ? idx->impl.last_launches() : 0; }
int64_t kjc_index_unverified_count(KjcIndex* idx) {
    if (!idx) return 0;
    int64_t v = -1;
    guarded([&] { v = idx->impl.unverified_count(); });
    return v;
}

float kjc_cosine_similarity(const float* a, const float* b, size_t len) {
    // kjarni_cosine_similarity (KF/src/lib.rs:177-188) -> VectorStore::cosine_similarity (KS/vector.rs:131-148)
    if (!a || !b || len == 0) return 0.0f;
    float dot = 0.f, na = 0.f, nb = 0.f;
    for (size_t i = 0; i < len; ++i) { dot += a[i] * b[i]; na += a[i] * a[i]; nb += b[i] * b[i]; }
    const float den = fmaxf(sqrtf(na) * sqrtf(nb), 1e-9f);
    return dot / den;
}

// -------------------------------------------------------------- debug hooks
// Single-kernel entry points used by tests/ to check each kernel in isolation (host buffers).
int kjc_dbg_gemm(const uint16_t* a_bf16, const uint16_t* w_bf16, const float* bias, const float* residual, int M, int N, int K, int epi,
                 int act, int block_n, void* out) {
    KJC_REQUIRE(a_bf16);
    KJC_REQUIRE(w_bf16);
    KJC_REQUIRE(out);
    return guarded([&] { kj::dbg_gemm(a_bf16, w_bf16, bias, residual, M, N, K, epi, act, block_n, out); });
}

int kjc_dbg_gemm_ln_gemm(const uint16_t* a_bf16, const uint16_t* w1_bf16, const float* bias1, const float* gamma, const float* beta, float eps,
                         const uint16_t* res_bf16, int M, int K1, const uint16_t* w2_bf16, const float* bias2, int N2, int epi2, int act,
                         uint16_t* out_x_bf16, uint16_t* out2_bf16, int iters, float* out_us) {
    return guarded([&] {
        kj::dbg_gemm_ln_gemm(a_bf16, w1_bf16, bias1, gamma, beta, eps, res_bf16, M, K1, w2_bf16, bias2, N2, epi2, act, out_x_bf16, out2_bf16, iters, out_us);
    });
}
int kjc_dbg_gemm_ln(const uint16_t* a_bf16, const uint16_t* w_bf16, const float* bias, const float* gamma, const float* beta, float eps,
                    const uint16_t* res_bf16, int M, int K, uint16_t* out_bf16, int iters, float* out_us) {
    KJC_REQUIRE(a_bf16); KJC_REQUIRE(w_bf16); KJC_REQUIRE(bias); KJC_REQUIRE(gamma); KJC_REQUIRE(beta); KJC_REQUIRE(res_bf16); KJC_REQUIRE(out_bf16);
    return guarded([&] { kj::dbg_gemm_ln(a_bf16, w_bf16, bias, gamma, beta, eps, res_bf16, M, 384, K, out_bf16, iters, out_us); });
}
int kjc_dbg_gemm_ln_h(const uint16_t* a_bf16, const uint16_t* w_bf16, const float* bias, const float* gamma, const float* beta, float eps,
                      const uint16_t* res_bf16, int M, int H, int K, uint16_t* out_bf16, int iters, float* out_us) {
    KJC_REQUIRE(a_bf16); KJC_REQUIRE(w_bf16); KJC_REQUIRE(bias); KJC_REQUIRE(gamma); KJC_REQUIRE(beta); KJC_REQUIRE(res_bf16); KJC_REQUIRE(out_bf16);
    return guarded([&] { kj::dbg_gemm_ln(a_bf16, w_bf16, bias, gamma, beta, eps, res_bf16, M, H, K, out_bf16, iters, out_us); });
}

int kjc_dbg_experimental_kernels(void) { return kj::experimental_kernels_built() ? 1 : 0; }
int kjc_dbg_ffn_ln(const uint16_t* x_bf16, const uint16_t* w1_bf16, const float* b1, const uint16_t* w2_bf16, const float* b2, const float* gamma,
                   const float* beta, float eps, int M, int I, int act, uint16_t* out_bf16, int iters, float* out_us) {
    KJC_REQUIRE(x_bf16); KJC_REQUIRE(w1_bf16); KJC_REQUIRE(b1); KJC_REQUIRE(w2_bf16); KJC_REQUIRE(b2); KJC_REQUIRE(gamma); KJC_REQUIRE(beta);
    KJC_REQUIRE(out_bf16);
    return guarded([&] { kj::dbg_ffn_ln(x_bf16, w1_bf16, b1, w2_bf16, b2, gamma, beta, eps, M, I, act, out_bf16, iters, out_us); });
}

int kjc_dbg_gemm_time(int M, int N, int K, int epi, int act, int block_n, int flags, int iters, float* out_us) {
    KJC_REQUIRE(out_us);
    return guarded([&] { *out_us = kj::dbg_gemm_time(M, N, K, epi, act, block_n, flags, iters); });
}

int kjc_dbg_attention(const uint16_t* qkv_bf16, const float* mask, int B, int S, int H, int heads, int nan_if_all_masked, uint16_t* ctx_bf16) {
    KJC_REQUIRE(qkv_bf16);
    KJC_REQUIRE(ctx_bf16);
    return guarded([&] { kj::dbg_attention(qkv_bf16, mask, B, S, H, heads, nan_if_all_masked, ctx_bf16); });
}

int kjc_dbg_index_set_filter(KjcIndex* idx, float eps, int min_queries) {
    KJC_REQUIRE(idx);
    return guarded([&] { idx->impl.set_filter(eps, min_queries); });
}

int kjc_dbg_encoder_head(KjcEncoder* enc, const float* hidden, int batch, int seq_len, float* logits) {
    KJC_REQUIRE(enc);
    KJC_REQUIRE(hidden);
    KJC_REQUIRE(logits);
    return guarded([&] { enc->impl.head_only_host(hidden, batch, seq_len, logits); });
}

}  // extern "C"
