// Index shard handle behind KjcIndex (see include/kjarni_cuda.h and index.cu).
#pragma once
#include <mutex>
#include <string>
#include <vector>

#include "encoder.hpp"

namespace kj {

// list l of the candidates starts at d_ids + l * stride / d_scores + l * stride (stride in elements; 0 = nq * k, densely packed)
void merge_lists_u64(const uint64_t* d_ids, const float* d_scores, int n_lists, int nq, int k, uint64_t* d_out_ids, float* d_out_scores,
                     int32_t* d_out_counts, cudaStream_t st, size_t ids_stride = 0, size_t scores_stride = 0);

// Bytes of one shard's packed candidate record: [nq,k] u64 ids | [nq,k] f32 scores, rounded up to 16 bytes.
inline size_t packed_record_bytes(int nq, int k) { return (static_cast<size_t>(nq) * k * 12 + 15) & ~static_cast<size_t>(15); }

class Index {
  public:
    Index(int dim, uint64_t capacity_rows, uint64_t id_base, int device);
    ~Index();
    Index(const Index&) = delete;
    Index& operator=(const Index&) = delete;

    uint64_t len() const { return len_; }
    int dim() const { return dim_; }
    uint64_t id_base() const { return id_base_; }
    int device() const { return device_; }
    int64_t last_launches() const { return launches_; }
    cudaStream_t stream() const { return stream_; }

    void add_rows_host(const float* rows, uint64_t n);
    void load_vectors_bin(const std::string& path);
    void load_vectors_bin_range(const std::string& path, uint64_t row0, uint64_t n_rows);
    void append_synthetic(uint32_t seed, uint64_t row0, uint64_t n);
    void get_rows(uint64_t row, uint64_t n, float* out) const;
    void search_host(const float* q, int nq, int k, int mode, uint64_t* ids, float* scores, int32_t* counts);
    // may_sync: the caller allows a stream synchronise (host-buffer API) so that queries the tensor-core filter could not
    // prove exact are re-run on the exact scan; without it they are counted in unverified_count().
    void search_device(const float* d_q, int nq, int k, int mode, uint64_t* d_ids, float* d_scores, int32_t* d_counts, cudaStream_t st,
                       bool may_sync = false);
    int64_t unverified_count();
    void set_filter(float eps, int min_queries);  // test hook: proof margin and the batch size from which the GEMM filter is used

  private:
    void compute_norms(uint64_t row0, uint64_t n, cudaStream_t st);
    void search_exact(const float* d_q, int nq, int k, int mode, uint64_t* d_ids, float* d_scores, int32_t* d_counts, cudaStream_t st);
    void search_gemm(const float* d_q, int nq, int k, int mode, uint64_t* d_ids, float* d_scores, int32_t* d_counts, cudaStream_t st,
                     bool may_sync);
    int dim_;
    uint64_t cap_, id_base_, len_ = 0;
    int device_, num_sms_ = 0;
    mutable std::recursive_mutex mu_;  // one lock per handle call, held from staging-in to the final synchronise (kjarni_cuda.h: calls on one handle are serialised)
    cudaStream_t stream_ = nullptr;
    float *rows_ = nullptr, *norms_ = nullptr;
    float *d_q_ = nullptr, *d_qn_ = nullptr, *d_cand_s_ = nullptr, *d_out_s_ = nullptr;
    uint32_t* d_cand_i_ = nullptr;
    uint64_t* d_out_i_ = nullptr;
    int32_t* d_out_c_ = nullptr;
    float* h_stage_ = nullptr;
    size_t q_cap_ = 0, qn_cap_ = 0, cand_cap_ = 0, out_cap_ = 0, outc_cap_ = 0;
    int64_t launches_ = 0;
    // tensor-core filter path (scan_gemm.cuh): bf16 shadow of the rows, 1/|r|, staging for query tiles and candidates
    bool gemm_ok_ = false;
    bool scan_t8_ = false;  // exact scan: lane-per-row kernel (dim % 32 == 0); otherwise the warp-per-row kernels
    CUtensorMap t_rows32_;  // fp32 rows as 32 x 32-float boxes, 128-byte swizzle
    float filter_eps_ = 0.0045f;
    int filter_min_q_ = 1;  // the filter pass reads half the bytes of the exact scan, so it wins from a single query on
    __nv_bfloat16 *rows16_ = nullptr, *d_q16_ = nullptr;
    float *d_gc_s_ = nullptr, *d_am_s_ = nullptr, *d_fix_q_ = nullptr, *d_fix_s_ = nullptr;
    uint32_t* d_gc_i_ = nullptr;
    uint64_t *d_am_i_ = nullptr, *d_fix_i_ = nullptr;
    int32_t *d_flags_ = nullptr, *d_nflag_ = nullptr, *d_fix_c_ = nullptr;
    float* d_seed_ = nullptr;
    int32_t* d_esc_map_ = nullptr;  // escalation of unproven queries: their indices, 128-candidate lists
    float* d_esc_s_ = nullptr;
    uint64_t* d_esc_i_ = nullptr;
    size_t esc_cap_ = 0;
    bool escalate_ = true;  // KJC_SCAN_NO_ESCALATE: unproven queries go straight to the exact scan (test / measurement hook)
    size_t q16_cap_ = 0, gc_cap_ = 0, am_cap_ = 0, flags_cap_ = 0, seed_cap_ = 0;
    CUtensorMap t_rows16_, t_rows16_half_;
    std::vector<int32_t> h_flags_;
};

}  // namespace kj

namespace kj {

// On-disk index directory as IndexWriter leaves it (kjarni-rag/src/index_writer.rs:128-170, config.rs:5-27):
//   <root>/config.json                         IndexConfig {dimension, max_docs_per_segment, ...}
//   <root>/segments/seg_%06d/segment.json      SegmentMeta {id, doc_count, dimension, created_at, total_bytes}
//   <root>/segments/seg_%06d/vectors.bin       raw LE f32 [doc_count, dimension]
//   (+ docs.bin, docs.idx, metadata.jsonl, bm25.bin: text / keyword side, stays with the Rust host)
// Segments are taken in file-name order and a segment the reference's Segment::open would reject (missing segment.json,
// vectors.bin, docs.idx or bm25.bin) is skipped exactly as IndexReader::open does (index_reader.rs:161-204), so the
// global ids (sum of preceding segment lengths + local id, index_reader.rs:313-319) agree with the host's.
struct IndexDirSegment {
    std::string dir;       // absolute path of the segment directory
    uint64_t doc_count;    // SegmentMeta::doc_count
    uint64_t global_base;  // global id of local row 0
};
struct IndexDir {
    int dimension = 0;
    uint64_t max_docs_per_segment = 0;
    uint64_t total_rows = 0;
    int skipped = 0;  // segment directories IndexReader::open would have skipped
    std::vector<IndexDirSegment> segments;
};
IndexDir scan_index_dir(const std::string& root);
// Rows [lo, hi) of global-id space of part `part` of `parts` (contiguous, sizes differ by at most one row).
void index_part_range(uint64_t total, int part, int parts, uint64_t* lo, uint64_t* hi);
// New shard on `device` holding this part's rows of the on-disk index; id_base = lo.
Index* open_index_dir(const std::string& root, int device, int part, int parts);

}  // namespace kj
