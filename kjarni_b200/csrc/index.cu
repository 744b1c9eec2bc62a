// Index shard handle behind KjcIndex: fp32 rows resident in HBM, row norms cached at
// append time, brute-force cosine top-k search and the candidate merge.
// Host-side mirror of Segment (kjarni-rag/src/segment.rs:195-337), VectorStore
// (kjarni-search/src/vector.rs:5-165) and the per-segment -> merge shape of
// IndexReader::search_semantic (kjarni-rag/src/index_reader.rs:207-228,313-319).
#include <dirent.h>

#include <algorithm>
#include <mutex>

#include "index.hpp"
#include "scan.cuh"
#include "scan_gemm.cuh"

namespace kj {

Index::Index(int dim, uint64_t capacity, uint64_t id_base, int device) : dim_(dim), cap_(capacity), id_base_(id_base), device_(device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) throw Error(KJC_GPU_UNAVAILABLE, "no CUDA device available");
    if (device < 0 || device >= ndev) throw Error(KJC_GPU_UNAVAILABLE, "device index out of range");
    if (dim <= 0 || dim % 4 != 0 || dim > 1024) throw Error(KJC_INVALID_CONFIG, "index dimension must be a multiple of 4 and <= 1024");
    if (capacity == 0 || capacity > 0xFFFFFFF0ull) throw Error(KJC_INVALID_CONFIG, "shard capacity must be in [1, 2^32-16] rows");
    cudaDeviceProp prop;
    KJ_CUDA(cudaGetDeviceProperties(&prop, device));
    num_sms_ = prop.multiProcessorCount;
    KJ_CUDA(cudaSetDevice(device));
    KJ_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    KJ_CUDA(cudaMalloc(&rows_, capacity * dim * sizeof(float)));
    KJ_CUDA(cudaMalloc(&norms_, capacity * sizeof(float) + 64));  // slack: tail bulk copies round up to 16 B
    // tensor-core filter path: bf16 shadow + 1/|r| (dims the resident 128-query tile supports)
    gemm_ok_ = dim % kSgBK == 0 && dim <= kSgMaxDStream && !getenv("KJC_SCAN_NO_GEMM");
    if (const char* e = getenv("KJC_SCAN_EPS")) filter_eps_ = static_cast<float>(atof(e));
    if (const char* e = getenv("KJC_SCAN_GEMM_MIN_Q")) filter_min_q_ = std::max(1, atoi(e));
    // exact scan: lane-per-row kernel on TMA-swizzled 32 x 32-float boxes when the dimension allows it (scan.cuh, scan_t8_kernel)
    escalate_ = getenv("KJC_SCAN_NO_ESCALATE") == nullptr;
    scan_t8_ = dim % 32 == 0 && getenv("KJC_SCAN_NO_T8") == nullptr;
    if (scan_t8_) t_rows32_ = make_tmap_2d(rows_, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, capacity, dim, kT8Rows, 32, 128);  // 64 rows x 32 floats
    if (gemm_ok_) {
        const uint64_t cap16 = std::max<uint64_t>(capacity, kSgRows);  // at least one TMA box of rows
        KJ_CUDA(cudaMalloc(&rows16_, cap16 * dim * sizeof(__nv_bfloat16)));
        KJ_CUDA(cudaMemsetAsync(rows16_, 0, cap16 * dim * sizeof(__nv_bfloat16), stream_));
        KJ_CUDA(cudaMalloc(&d_nflag_, 2 * sizeof(int32_t)));  // [0] flagged in the current search, [1] unverified (async) total
        KJ_CUDA(cudaMemsetAsync(d_nflag_, 0, 2 * sizeof(int32_t), stream_));
        KJ_CUDA(cudaMalloc(&d_fix_q_, 8 * dim * sizeof(float)));
        KJ_CUDA(cudaMalloc(&d_fix_s_, 8 * 256 * sizeof(float)));
        KJ_CUDA(cudaMalloc(&d_fix_i_, 8 * 256 * sizeof(uint64_t)));
        KJ_CUDA(cudaMalloc(&d_fix_c_, 8 * sizeof(int32_t)));
        KJ_CUDA(cudaStreamSynchronize(stream_));
        t_rows16_ = make_tmap_2d(rows16_, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, cap16, dim, kSgRows, kSgBK, 128);
        t_rows16_half_ = make_tmap_2d(rows16_, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, cap16, dim, kSgRows / 2, kSgBK, 128);  // CTA-pair filter pass: half a row tile per CTA
    }
}

Index::~Index() {
    cudaSetDevice(device_);
    for (void* p : {(void*)rows_, (void*)norms_, (void*)d_q_, (void*)d_qn_, (void*)d_cand_s_, (void*)d_cand_i_, (void*)d_out_s_,
                    (void*)d_out_i_, (void*)d_out_c_, (void*)rows16_, (void*)d_q16_, (void*)d_gc_s_, (void*)d_gc_i_,
                    (void*)d_am_s_, (void*)d_am_i_, (void*)d_fix_q_, (void*)d_fix_s_, (void*)d_fix_i_, (void*)d_fix_c_, (void*)d_flags_,
                    (void*)d_nflag_, (void*)d_seed_, (void*)d_esc_map_, (void*)d_esc_s_, (void*)d_esc_i_})
        if (p) cudaFree(p);
    if (h_stage_) cudaFreeHost(h_stage_);
    if (stream_) cudaStreamDestroy(stream_);
}

// |row| + bf16 shadow: NCH = float4 chunks per lane (dim <= 384: 3, <= 768: 6, <= 1024: 8)
static void launch_row_prep(const float* rows, float* norms, __nv_bfloat16* rows16, size_t n, int D, int normalise, cudaStream_t st) {
    const unsigned grid = static_cast<unsigned>((n + 7) / 8);
    if (D <= 384) row_prep_kernel<3><<<grid, 256, 0, st>>>(rows, norms, rows16, n, D, normalise);
    else if (D <= 768) row_prep_kernel<6><<<grid, 256, 0, st>>>(rows, norms, rows16, n, D, normalise);
    else row_prep_kernel<8><<<grid, 256, 0, st>>>(rows, norms, rows16, n, D, normalise);
}

void Index::compute_norms(uint64_t row0, uint64_t n, cudaStream_t st) {
    if (n == 0) return;
    if (gemm_ok_)
        launch_row_prep(rows_ + row0 * dim_, norms_ + row0, rows16_ + row0 * dim_, n, dim_, 1, st);
    else
        row_norm_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, st>>>(rows_ + row0 * dim_, norms_ + row0, n, dim_);
    KJ_CUDA(cudaGetLastError());
}

void Index::add_rows_host(const float* rows, uint64_t n) {
    std::lock_guard<std::recursive_mutex> lock(mu_);
    if (len_ + n > cap_) throw Error(KJC_INVALID_CONFIG, "index shard capacity exceeded");
    KJ_CUDA(cudaSetDevice(device_));
    // chunked pinned staging (true async DMA, bounded host memory)
    const size_t chunk_rows = std::max<size_t>(1, (32u << 20) / (dim_ * sizeof(float)));
    if (!h_stage_) KJ_CUDA(cudaMallocHost(&h_stage_, chunk_rows * dim_ * sizeof(float)));
    uint64_t done = 0;
    while (done < n) {
        const uint64_t m = std::min<uint64_t>(chunk_rows, n - done);
        memcpy(h_stage_, rows + done * dim_, m * dim_ * sizeof(float));
        KJ_CUDA(cudaMemcpyAsync(rows_ + (len_ + done) * dim_, h_stage_, m * dim_ * sizeof(float), cudaMemcpyHostToDevice, stream_));
        KJ_CUDA(cudaStreamSynchronize(stream_));
        done += m;
    }
    compute_norms(len_, n, stream_);
    KJ_CUDA(cudaStreamSynchronize(stream_));
    len_ += n;
}

void Index::load_vectors_bin(const std::string& path) { load_vectors_bin_range(path, 0, UINT64_MAX); }

void Index::load_vectors_bin_range(const std::string& path, uint64_t row0, uint64_t n_rows) {
    // vectors.bin = raw little-endian f32 [doc_count x dim] (kjarni-rag/src/segment.rs:87-123,211-262);
    // n_rows == UINT64_MAX: every row from row0 to the end of the file
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) throw Error(KJC_MODEL_NOT_FOUND, "cannot open " + path);
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); throw Error(KJC_LOAD_FAILED, "cannot stat " + path); }
    const size_t bytes = static_cast<size_t>(sb.st_size);
    const size_t row_bytes = dim_ * sizeof(float);
    if (bytes % row_bytes != 0) { close(fd); throw Error(KJC_LOAD_FAILED, "vectors.bin size is not a multiple of 4*dim: " + path); }
    const uint64_t file_rows = bytes / row_bytes;
    if (n_rows == UINT64_MAX) n_rows = file_rows > row0 ? file_rows - row0 : 0;
    if (row0 + n_rows > file_rows) { close(fd); throw Error(KJC_LOAD_FAILED, "vectors.bin holds fewer rows than segment.json declares: " + path); }
    if (n_rows == 0) { close(fd); return; }
    void* map = mmap(nullptr, bytes, PROT_READ, MAP_PRIVATE, fd, 0);
    if (map == MAP_FAILED) { close(fd); throw Error(KJC_LOAD_FAILED, "mmap failed for " + path); }
    try {
        add_rows_host(static_cast<const float*>(map) + row0 * dim_, n_rows);
    } catch (...) {
        munmap(map, bytes);
        close(fd);
        throw;
    }
    munmap(map, bytes);
    close(fd);
}

void Index::append_synthetic(uint32_t seed, uint64_t row0, uint64_t n) {
    std::lock_guard<std::recursive_mutex> lock(mu_);
    if (len_ + n > cap_) throw Error(KJC_INVALID_CONFIG, "index shard capacity exceeded");
    KJ_CUDA(cudaSetDevice(device_));
    const size_t total = n * dim_;
    const uint64_t per = 1ull << 30;  // elements per launch (grid-size limit)
    for (uint64_t off_rows = 0; off_rows < n;) {
        const uint64_t rows_now = std::min<uint64_t>(n - off_rows, per / dim_);
        const size_t elems = rows_now * dim_;
        synth_rows_kernel<<<static_cast<unsigned>((elems + 255) / 256), 256, 0, stream_>>>(rows_ + (len_ + off_rows) * dim_, seed,
                                                                                         row0 + off_rows, rows_now, dim_);
        KJ_CUDA(cudaGetLastError());
        off_rows += rows_now;
    }
    (void)total;
    compute_norms(len_, n, stream_);
    KJ_CUDA(cudaStreamSynchronize(stream_));
    len_ += n;
}

void Index::get_rows(uint64_t row, uint64_t n, float* out) const {
    std::lock_guard<std::recursive_mutex> lock(mu_);
    if (row + n > len_) throw Error(KJC_INVALID_CONFIG, "Document ID out of range");
    KJ_CUDA(cudaSetDevice(device_));
    KJ_CUDA(cudaMemcpy(out, rows_ + row * dim_, n * dim_ * sizeof(float), cudaMemcpyDeviceToHost));
}

template <int QT, int NCH, int CW = kScanConsumerWarps, int RU = 1>
static void launch_scan_inst(ScanParams p, int grid, cudaStream_t st) {
    static int configured[64] = {0};
    // rows per stage / stage count: as much as fits in ~200 KB next to the top-k lists
    p.rows_per_stage = p.D <= 512 ? 32 : 16;
    const size_t budget = 200 * 1024 - scan_list_bytes(QT, p.k, CW) - 512;
    p.nstages = static_cast<int>(std::min<size_t>(4, budget / scan_stage_bytes(p.D, p.rows_per_stage)));
    if (p.nstages < 2) throw Error(KJC_INVALID_CONFIG, "index dimension too large for the scan pipeline");
    const size_t smem = scan_smem_bytes(p.D, p.rows_per_stage, p.nstages, QT, p.k, CW);
    auto kern = scan_topk_kernel<QT, NCH, CW, RU>;
    ensure_smem_attr(kern, static_cast<int>(smem), configured);
    kern<<<grid, (CW + 1) * 32, smem, st>>>(p);
    KJ_CUDA(cudaGetLastError());
}
// Lane-per-row exact scan (scan_t8_kernel, dim % 32 == 0): <= 8 queries per pass; writes kT8Teams lists per CTA and query
constexpr int kT8Teams = 2;
static void launch_scan_t8(const CUtensorMap& t_rows, ScanT8Params p, int grid, cudaStream_t st) {
    static int configured[64] = {0};
    // floats of a row per stage: the largest multiple of 32 that divides D and is <= 192 (64 rows x 192 floats = 48 KB stages at most)
    int ds = 32;
    for (int c = 32; c <= 192 && c <= p.D; c += 32)
        if (p.D % c == 0) ds = c;
    p.ds = ds;
    const size_t fixed = scan_t8_smem_bytes(p.D, ds, 0, p.k, kT8Teams);
    p.nstages = static_cast<int>(std::min<size_t>(4, (226 * 1024 - fixed) / (static_cast<size_t>(ds) * kT8Rows * 4)));
    if (p.nstages < 2) throw Error(KJC_INVALID_CONFIG, "index dimension too large for the scan pipeline");
    const size_t smem = scan_t8_smem_bytes(p.D, ds, p.nstages, p.k, kT8Teams);
    ensure_smem_attr(scan_t8_kernel<kT8Teams>, static_cast<int>(smem), configured);
    scan_t8_kernel<kT8Teams><<<grid, t8_threads<kT8Teams>(), smem, st>>>(t_rows, p);
    KJ_CUDA(cudaGetLastError());
}

template <int QT>
static void launch_scan_qt(const ScanParams& p, int grid, cudaStream_t st) {
    const int nch = (p.D + 127) / 128;
    switch (nch) {
        case 1: launch_scan_inst<QT, 1>(p, grid, st); break;
        case 2: launch_scan_inst<QT, 2>(p, grid, st); break;
        case 3: launch_scan_inst<QT, 3>(p, grid, st); break;
        default:
            if constexpr (QT <= 2) {
                if (nch == 4) launch_scan_inst<QT, 4>(p, grid, st);
                else if (nch <= 6) launch_scan_inst<QT, 6>(p, grid, st);
                else launch_scan_inst<QT, 8>(p, grid, st);
            } else {
                throw Error(KJC_INVALID_CONFIG, "internal: 4 queries per warp need dim <= 384");
            }
            break;
    }
}

static void launch_topk_merge(const MergeParams& m, cudaStream_t st) {
    topk_merge_kernel<<<m.Q, 256, static_cast<size_t>(m.L) * sizeof(int), st>>>(m);
    KJ_CUDA(cudaGetLastError());
}

void Index::search_device(const float* d_q, int nq, int k, int mode, uint64_t* d_ids, float* d_scores, int32_t* d_counts, cudaStream_t st,
                          bool may_sync) {
    if (nq < 1) throw Error(KJC_INVALID_CONFIG, "nq must be >= 1");
    if (k < 1 || k > 256) throw Error(KJC_INVALID_CONFIG, "k must be in [1, 256]");
    if (mode != SCAN_SEGMENT && mode != SCAN_VECTORSTORE) throw Error(KJC_INVALID_CONFIG, "unknown scan mode");
    std::lock_guard<std::recursive_mutex> lock(mu_);
    KJ_CUDA(cudaSetDevice(device_));
    if (!st) st = stream_;
    launches_ = 0;
    // many queries: tensor-core filter + exact rescoring; few queries: the exact HBM-bound scan
    // the filter keeps C >= 2k approximate candidates per query: C = 32 for k <= 16, 128 for k <= 64 (a reranking Searcher fetches
    // top_k * 5 = 50); beyond that the exact scan
    if (gemm_ok_ && nq >= filter_min_q_ && 2 * k <= kSgCWide && len_ > 0) search_gemm(d_q, nq, k, mode, d_ids, d_scores, d_counts, st, may_sync);
    else search_exact(d_q, nq, k, mode, d_ids, d_scores, d_counts, st);
}

int64_t Index::unverified_count() {
    std::lock_guard<std::recursive_mutex> lock(mu_);
    if (!d_nflag_) return 0;
    KJ_CUDA(cudaSetDevice(device_));
    int32_t v[2] = {0, 0};
    KJ_CUDA(cudaDeviceSynchronize());
    KJ_CUDA(cudaMemcpy(v, d_nflag_, sizeof(v), cudaMemcpyDeviceToHost));
    return v[1];
}

void Index::set_filter(float eps, int min_queries) {
    std::lock_guard<std::recursive_mutex> lock(mu_);
    filter_eps_ = eps;
    filter_min_q_ = std::max(1, min_queries);
}

static __global__ void add_counter_kernel(const int32_t* src, int32_t* dst) { atomicAdd(dst, *src); }

// Tensor-core path (scan_gemm.cuh): query prep -> seed pass -> filter pass -> candidate select -> exact rescoring +
// proof check -> (sync callers) exact re-run of the queries that could not be proven.
void Index::search_gemm(const float* d_q_all, int nq_all, int k, int mode, uint64_t* d_ids_all, float* d_scores_all, int32_t* d_counts_all,
                        cudaStream_t st, bool may_sync) {
    static int configured[64] = {0}, configured_pair[64] = {0};
    // filter pass as CTA pairs (scan_gemm_kernel<true>): two query tiles against the same row tiles, half a row tile per CTA
    static const bool env_pair = getenv("KJC_SG_PAIR") ? atoi(getenv("KJC_SG_PAIR")) != 0 : true;
    static const int env_R = getenv("KJC_SG_R") ? std::max(1, atoi(getenv("KJC_SG_R"))) : 3;
    static const int env_dbg = getenv("KJC_SG_DBG") ? atoi(getenv("KJC_SG_DBG")) : 0;
    static const bool env_no_seed = getenv("KJC_SG_NO_SEED") != nullptr;
    ensure_smem_attr(scan_gemm_kernel<false>, kSgSmemBytes, configured);
    ensure_smem_attr(scan_gemm_kernel<true>, kSgSmemBytes, configured_pair);
    const uint32_t n_tiles = static_cast<uint32_t>((len_ + kSgRows - 1) / kSgRows);
    const int grid = static_cast<int>(std::min<uint32_t>(num_sms_, n_tiles));
    const int qb_max = std::min(nq_all, 4096);  // queries per pass over the shard (bounds the candidate buffers)
    const int C = 2 * k <= kSgC ? kSgC : kSgCWide;  // approximate candidates kept per query
    {
        const size_t q16 = static_cast<size_t>(std::max(qb_max, kSgQ)) * dim_;
        if (q16 > q16_cap_) {
            if (d_q16_) cudaFree(d_q16_);
            KJ_CUDA(cudaMalloc(&d_q16_, q16 * sizeof(__nv_bfloat16)));
            q16_cap_ = q16;
        }
        if (static_cast<size_t>(qb_max) > qn_cap_) {
            if (d_qn_) cudaFree(d_qn_);
            KJ_CUDA(cudaMalloc(&d_qn_, static_cast<size_t>(std::max(qb_max, 8)) * 4));
            qn_cap_ = std::max(qb_max, 8);
        }
        const size_t gc = static_cast<size_t>(qb_max) * kSgCap;
        if (gc > gc_cap_) {
            if (d_gc_s_) cudaFree(d_gc_s_);
            if (d_gc_i_) cudaFree(d_gc_i_);
            KJ_CUDA(cudaMalloc(&d_gc_s_, gc * 4));
            KJ_CUDA(cudaMalloc(&d_gc_i_, gc * 4));
            gc_cap_ = gc;
        }
        const size_t am = static_cast<size_t>(qb_max) * C;
        if (am > am_cap_) {
            if (d_am_s_) cudaFree(d_am_s_);
            if (d_am_i_) cudaFree(d_am_i_);
            KJ_CUDA(cudaMalloc(&d_am_s_, am * 4));
            KJ_CUDA(cudaMalloc(&d_am_i_, am * 8));
            am_cap_ = am;
        }
        if (static_cast<size_t>(qb_max) > flags_cap_) {  // flags | overflow | candidate counters
            if (d_flags_) cudaFree(d_flags_);
            KJ_CUDA(cudaMalloc(&d_flags_, static_cast<size_t>(qb_max) * 3 * 4));
            flags_cap_ = qb_max;
        }
        const size_t seed_elems = static_cast<size_t>(kSgSeedGroupsMax + 1) * qb_max;
        if (seed_elems > seed_cap_) {
            if (d_seed_) cudaFree(d_seed_);
            KJ_CUDA(cudaMalloc(&d_seed_, seed_elems * 4));
            seed_cap_ = seed_elems;
        }
    }
    for (int qoff = 0; qoff < nq_all; qoff += qb_max) {
        const int nq = std::min(qb_max, nq_all - qoff);
        const float* d_q = d_q_all + static_cast<size_t>(qoff) * dim_;
        uint64_t* d_ids = d_ids_all + static_cast<size_t>(qoff) * k;
        float* d_scores = d_scores_all + static_cast<size_t>(qoff) * k;
        int32_t* d_counts = d_counts_all ? d_counts_all + qoff : nullptr;
        int32_t* d_overflow = d_flags_ + flags_cap_;
        uint32_t* d_cnt = reinterpret_cast<uint32_t*>(d_flags_ + 2 * flags_cap_);
        float* d_seed = d_seed_;                                                 // [groups, nq]
        float* d_thr0 = d_seed_ + static_cast<size_t>(kSgSeedGroupsMax) * nq;    // [nq]

        // |q| and the bf16 copy of the queries (rows beyond nq of the last 128-query tile are zero-filled by TMA)
        launch_row_prep(d_q, d_qn_, d_q16_, static_cast<size_t>(nq), dim_, 0, st);
        KJ_CUDA(cudaGetLastError());
        if (nq < kSgQ)
            KJ_CUDA(cudaMemsetAsync(d_q16_ + static_cast<size_t>(nq) * dim_, 0, static_cast<size_t>(kSgQ - nq) * dim_ * 2, st));
        KJ_CUDA(cudaMemsetAsync(d_nflag_, 0, sizeof(int32_t), st));
        KJ_CUDA(cudaMemsetAsync(d_cnt, 0, static_cast<size_t>(nq) * 4, st));
        ++launches_;
        const CUtensorMap t_q = make_tmap_2d(d_q16_, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, std::max(nq, kSgQ), dim_, kSgQ, kSgBK, 128);
        ScanGemmParams sp;
        sp.thr0 = nullptr; sp.cand_scores = d_gc_s_; sp.cand_ids = d_gc_i_; sp.cand_cnt = d_cnt;
        sp.seed_max = nullptr; sp.n_rows = static_cast<uint32_t>(len_); sp.n_tiles = n_tiles; sp.n_tiles_total = n_tiles;
        sp.seed_chunks = 0; sp.D = dim_; sp.Q = nq; sp.R = env_R; sp.dbg = env_dbg; sp.stream_a = dim_ > kSgMaxD ? 1 : 0;
        // ---- seed pass
        int seed_groups = 0;
        if (len_ > static_cast<uint64_t>(kSgCap) && !env_no_seed) {
            ScanGemmParams ss = sp;
            ss.seed_max = d_seed;
            int sgrid;
            if (n_tiles < static_cast<uint32_t>(std::min(C, kSgSeedGroupsMax / 8))) {  // small shard: every tile, one group per 32-row chunk
                ss.seed_chunks = 1; ss.n_tiles = n_tiles; ss.R = 1; sgrid = static_cast<int>(n_tiles);
                seed_groups = static_cast<int>(n_tiles) * 8;
            } else {  // evenly spaced sample tiles, 8 per CTA, one group per CTA
                ss.R = 2; sgrid = std::min(grid, kSgSeedGroupsMax / 2);  // 4 L2-resident superblocks of 2 tiles per CTA
                ss.n_tiles = std::min<uint32_t>(n_tiles, static_cast<uint32_t>(sgrid) * 8);
                seed_groups = 2 * sgrid;  // per CTA and column half
            }
            scan_gemm_kernel<false><<<sgrid, kSgThreads, kSgSmemBytes, st>>>(t_q, t_rows16_, ss);
            KJ_CUDA(cudaGetLastError());
            ++launches_;
        }
        scan_seed_select_kernel<<<(nq + 7) / 8, 256, 0, st>>>(d_seed, seed_groups, nq, d_thr0, C);  // 0 groups: thr0 = -inf
        KJ_CUDA(cudaGetLastError());
        ++launches_;
        // ---- filter pass
        sp.thr0 = d_thr0;
        if (env_pair && !sp.stream_a && nq > kSgQ && num_sms_ >= 2) {
            // more than one query tile, resident query tiles: pairs of CTAs share every row tile (superblocks of half as many tiles,
            // twice as many per worker, so the L2-resident footprint per superblock is unchanged)
            ScanGemmParams pp = sp;
            pp.R = 2 * sp.R;
            const unsigned clusters = std::min<uint32_t>(static_cast<uint32_t>(num_sms_ / 2), n_tiles);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2 * clusters);
            cfg.blockDim = dim3(kSgThreads);
            cfg.dynamicSmemBytes = kSgSmemBytes;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            KJ_CUDA(cudaLaunchKernelEx(&cfg, scan_gemm_kernel<true>, t_q, t_rows16_half_, pp));
        } else {
            scan_gemm_kernel<false><<<grid, kSgThreads, kSgSmemBytes, st>>>(t_q, t_rows16_, sp);
        }
        KJ_CUDA(cudaGetLastError());
        ++launches_;
        CandSelectParams cs;
        cs.cand_scores = d_gc_s_; cs.cand_ids = d_gc_i_; cs.cand_cnt = d_cnt; cs.id_base = id_base_;
        cs.out_scores = d_am_s_; cs.out_ids = d_am_i_; cs.overflow = d_overflow; cs.C = C; cs.qmap = nullptr;
        scan_cand_select_kernel<<<nq, 256, 0, st>>>(cs);
        KJ_CUDA(cudaGetLastError());
        ++launches_;
        RescoreParams r;
        r.rows = rows_; r.norms = norms_; r.queries = d_q; r.qnorms = d_qn_; r.cand_ids = d_am_i_; r.cand_scores = d_am_s_; r.id_base = id_base_;
        r.out_ids = d_ids; r.out_scores = d_scores; r.out_counts = d_counts; r.overflow = d_overflow; r.thr0 = d_thr0;
        r.flags = d_flags_; r.n_flagged = d_nflag_;
        r.eps = filter_eps_; r.D = dim_; r.Q = nq; r.k = k; r.mode = mode; r.qmap = nullptr;
        if (C == kSgC) scan_rescore_kernel<1><<<(nq + 7) / 8, 256, 0, st>>>(r);
        else scan_rescore_kernel<kSgCWide / 32><<<(nq + 7) / 8, 256, 0, st>>>(r);
        KJ_CUDA(cudaGetLastError());
        ++launches_;
        if (!may_sync) {
            add_counter_kernel<<<1, 1, 0, st>>>(d_nflag_, d_nflag_ + 1);
            KJ_CUDA(cudaGetLastError());
            continue;
        }
        int32_t nflag = 0;
        KJ_CUDA(cudaMemcpyAsync(&nflag, d_nflag_, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        KJ_CUDA(cudaStreamSynchronize(st));
        if (nflag == 0) continue;
        h_flags_.resize(nq);
        KJ_CUDA(cudaMemcpyAsync(h_flags_.data(), d_flags_, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, st));
        KJ_CUDA(cudaStreamSynchronize(st));
        std::vector<int> todo;
        for (int i = 0; i < nq; ++i)
            if (h_flags_[i]) todo.push_back(i);
        // Escalation before the exact scan: the filter pass kept EVERY row above the seed bound (up to 2048 per query), so an unproven
        // query is first re-selected with 128 instead of 32 candidates from the same buffer and re-scored -- a wider list lowers the
        // bound on everything outside it (or, when it holds all passers, the bound becomes the seed bound itself).  Costs two small
        // kernels over the flagged queries only; an exact pass over the shard costs ~1.5 ms per 8 queries.
        if (escalate_ && C == kSgC && !todo.empty() && 2 * k <= kSgCWide) {
            const size_t ns = todo.size();
            if (ns > esc_cap_) {
                for (void* q : {(void*)d_esc_map_, (void*)d_esc_s_, (void*)d_esc_i_}) if (q) cudaFree(q);
                KJ_CUDA(cudaMalloc(&d_esc_map_, ns * 4));
                KJ_CUDA(cudaMalloc(&d_esc_s_, ns * kSgCWide * 4));
                KJ_CUDA(cudaMalloc(&d_esc_i_, ns * kSgCWide * 8));
                esc_cap_ = ns;
            }
            KJ_CUDA(cudaMemcpyAsync(d_esc_map_, todo.data(), ns * 4, cudaMemcpyHostToDevice, st));
            KJ_CUDA(cudaMemsetAsync(d_nflag_, 0, sizeof(int32_t), st));
            CandSelectParams c2 = cs;
            c2.out_scores = d_esc_s_; c2.out_ids = d_esc_i_; c2.C = kSgCWide; c2.qmap = d_esc_map_;
            scan_cand_select_kernel<<<static_cast<unsigned>(ns), 256, 0, st>>>(c2);
            KJ_CUDA(cudaGetLastError());
            RescoreParams r2 = r;
            r2.cand_ids = d_esc_i_; r2.cand_scores = d_esc_s_; r2.Q = static_cast<int>(ns); r2.qmap = d_esc_map_;
            scan_rescore_kernel<kSgCWide / 32><<<static_cast<unsigned>((ns + 7) / 8), 256, 0, st>>>(r2);
            KJ_CUDA(cudaGetLastError());
            launches_ += 2;
            KJ_CUDA(cudaMemcpyAsync(&nflag, d_nflag_, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            KJ_CUDA(cudaStreamSynchronize(st));
            if (nflag == 0) continue;
            KJ_CUDA(cudaMemcpyAsync(h_flags_.data(), d_flags_, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, st));
            KJ_CUDA(cudaStreamSynchronize(st));
            std::vector<int> still;
            for (int i : todo)
                if (h_flags_[i]) still.push_back(i);
            todo.swap(still);
        }
        for (size_t b = 0; b < todo.size(); b += 8) {
            const int nb = static_cast<int>(std::min<size_t>(8, todo.size() - b));
            for (int j = 0; j < nb; ++j)
                KJ_CUDA(cudaMemcpyAsync(d_fix_q_ + static_cast<size_t>(j) * dim_, d_q + static_cast<size_t>(todo[b + j]) * dim_, dim_ * 4,
                                        cudaMemcpyDeviceToDevice, st));
            search_exact(d_fix_q_, nb, k, mode, d_fix_i_, d_fix_s_, d_fix_c_, st);
            for (int j = 0; j < nb; ++j) {
                const size_t o = static_cast<size_t>(todo[b + j]) * k;
                KJ_CUDA(cudaMemcpyAsync(d_ids + o, d_fix_i_ + static_cast<size_t>(j) * k, static_cast<size_t>(k) * 8, cudaMemcpyDeviceToDevice, st));
                KJ_CUDA(cudaMemcpyAsync(d_scores + o, d_fix_s_ + static_cast<size_t>(j) * k, static_cast<size_t>(k) * 4, cudaMemcpyDeviceToDevice, st));
                if (d_counts) KJ_CUDA(cudaMemcpyAsync(d_counts + todo[b + j], d_fix_c_ + j, 4, cudaMemcpyDeviceToDevice, st));
            }
        }
    }
}

// Exact path.  Enqueue: query norms -> scan passes of <= 8 queries -> merge of the per-CTA lists.
void Index::search_exact(const float* d_q, int nq, int k, int mode, uint64_t* d_ids, float* d_scores, int32_t* d_counts, cudaStream_t st) {
    // queries per warp: bounded by the per-warp list memory (k) and the register file (dim)
    int qt_max = k > 64 ? 1 : (k > 32 ? 2 : 4);
    if (dim_ > 384) qt_max = std::min(qt_max, 2);  // 4 queries x more than 3 float4 chunks per lane would spill
    const int grid = std::max<int>(1, static_cast<int>(std::min<uint64_t>(num_sms_, (len_ + 31) / 32)));  // one CTA per SM
    const int lists_per_cta = scan_t8_ ? kT8Teams : 1;
    const size_t cand = static_cast<size_t>(grid) * lists_per_cta * nq * k;
    if (cand > cand_cap_) {
        if (d_cand_s_) cudaFree(d_cand_s_);
        if (d_cand_i_) cudaFree(d_cand_i_);
        KJ_CUDA(cudaMalloc(&d_cand_s_, cand * 4));
        KJ_CUDA(cudaMalloc(&d_cand_i_, cand * 4));
        cand_cap_ = cand;
    }
    if (static_cast<size_t>(nq) > qn_cap_) {
        if (d_qn_) cudaFree(d_qn_);
        KJ_CUDA(cudaMalloc(&d_qn_, static_cast<size_t>(nq) * 4));
        qn_cap_ = nq;
    }
    row_norm_kernel<<<(nq + 7) / 8, 256, 0, st>>>(d_q, d_qn_, nq, dim_);
    KJ_CUDA(cudaGetLastError());
    ++launches_;
    if (len_ > 0) {
        ScanParams p;
        p.rows = rows_; p.norms = norms_; p.queries = d_q; p.qnorms = d_qn_; p.out_scores = d_cand_s_; p.out_ids = d_cand_i_;
        p.n_rows = len_; p.D = dim_; p.Q = nq; p.k = k; p.mode = mode;
        for (int q0 = 0; q0 < nq;) {
            p.q0 = q0;
            const int rem = nq - q0;
            if (scan_t8_) {  // dim % 32 == 0: the lane-per-row kernel, 8 queries per pass
                ScanT8Params t;
                t.norms = norms_; t.queries = d_q; t.qnorms = d_qn_; t.out_scores = d_cand_s_; t.out_ids = d_cand_i_;
                t.n_rows = len_; t.D = dim_; t.Q = nq; t.k = k; t.q0 = q0; t.mode = mode;
                launch_scan_t8(t_rows32_, t, grid, st);
                q0 += 8;
                ++launches_;
                continue;
            }
            int qt = qt_max;
            while (qt > 1 && qt / 2 >= rem) qt /= 2;  // smallest per-warp tile that still covers the remainder in one group
            p.ngroups = rem > qt ? 2 : 1;
            if (qt == 4) launch_scan_qt<4>(p, grid, st);
            else if (qt == 2) launch_scan_qt<2>(p, grid, st);
            else launch_scan_qt<1>(p, grid, st);
            q0 += qt * p.ngroups;
            ++launches_;
        }
    }
    MergeParams m;
    m.in_scores = d_cand_s_; m.in_ids32 = d_cand_i_; m.in_ids64 = nullptr; m.id_base = id_base_; m.qnorms = d_qn_;
    m.out_scores = d_scores; m.out_ids = d_ids; m.out_counts = d_counts; m.L = len_ > 0 ? grid * lists_per_cta : 0; m.Q = nq; m.k = k; m.mode = mode;
    m.ids_stride = m.scores_stride = 0;
    launch_topk_merge(m, st);
    ++launches_;
}

void Index::search_host(const float* q, int nq, int k, int mode, uint64_t* ids, float* scores, int32_t* counts) {
    if (nq < 1) throw Error(KJC_INVALID_CONFIG, "nq must be >= 1");
    if (k < 1 || k > 256) throw Error(KJC_INVALID_CONFIG, "k must be in [1, 256]");
    // the staging buffers (d_q_, d_out_*) are shared by every caller of this handle: one lock from the H2D copy of the queries to
    // the synchronise after the D2H copies, so that two host threads cannot overwrite each other's queries or read each other's results
    std::lock_guard<std::recursive_mutex> lock(mu_);
    KJ_CUDA(cudaSetDevice(device_));
    const size_t qe = static_cast<size_t>(nq) * dim_, oe = static_cast<size_t>(nq) * k;
    {
        if (qe > q_cap_) {
            if (d_q_) cudaFree(d_q_);
            KJ_CUDA(cudaMalloc(&d_q_, qe * 4));
            q_cap_ = qe;
        }
        if (oe > out_cap_) {
            for (void* p : {(void*)d_out_s_, (void*)d_out_i_}) if (p) cudaFree(p);
            KJ_CUDA(cudaMalloc(&d_out_s_, oe * 4));
            KJ_CUDA(cudaMalloc(&d_out_i_, oe * 8));
            out_cap_ = oe;
        }
        if (static_cast<size_t>(nq) > outc_cap_) {
            if (d_out_c_) cudaFree(d_out_c_);
            KJ_CUDA(cudaMalloc(&d_out_c_, static_cast<size_t>(nq) * 4));
            outc_cap_ = nq;
        }
    }
    KJ_CUDA(cudaMemcpyAsync(d_q_, q, qe * 4, cudaMemcpyHostToDevice, stream_));
    search_device(d_q_, nq, k, mode, d_out_i_, d_out_s_, d_out_c_, stream_, /*may_sync=*/true);
    KJ_CUDA(cudaMemcpyAsync(ids, d_out_i_, oe * 8, cudaMemcpyDeviceToHost, stream_));
    KJ_CUDA(cudaMemcpyAsync(scores, d_out_s_, oe * 4, cudaMemcpyDeviceToHost, stream_));
    if (counts) KJ_CUDA(cudaMemcpyAsync(counts, d_out_c_, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, stream_));
    KJ_CUDA(cudaStreamSynchronize(stream_));
}

void merge_lists_u64(const uint64_t* d_ids, const float* d_scores, int n_lists, int nq, int k, uint64_t* d_out_ids, float* d_out_scores,
                     int32_t* d_out_counts, cudaStream_t st, size_t ids_stride, size_t scores_stride) {
    MergeParams m;
    m.ids_stride = ids_stride;
    m.scores_stride = scores_stride;
    m.in_scores = d_scores; m.in_ids32 = nullptr; m.in_ids64 = d_ids; m.id_base = 0; m.qnorms = nullptr;
    m.out_scores = d_out_scores; m.out_ids = d_out_ids; m.out_counts = d_out_counts; m.L = n_lists; m.Q = nq; m.k = k;
    m.mode = SCAN_VECTORSTORE;
    launch_topk_merge(m, st);
}

// ------------------------------------------------------------------ on-disk index directory (IndexReader::open)
static bool file_exists(const std::string& p) {
    struct stat sb;
    return stat(p.c_str(), &sb) == 0 && S_ISREG(sb.st_mode);
}

IndexDir scan_index_dir(const std::string& root) {
    IndexDir d;
    const std::string cfg_text = read_text_file(root + "/config.json", KJC_MODEL_NOT_FOUND);
    Json cfg;
    try {
        cfg = JsonParser(cfg_text.data(), cfg_text.size()).parse();
    } catch (const Error& e) {
        throw Error(KJC_LOAD_FAILED, root + "/config.json: " + e.what());
    }
    // IndexConfig: `dimension` and `max_docs_per_segment` are required serde fields (kjarni-rag/src/config.rs:5-14)
    if (!cfg.has("dimension") || !cfg.has("max_docs_per_segment"))
        throw Error(KJC_LOAD_FAILED, root + "/config.json: missing field `dimension` / `max_docs_per_segment`");
    d.dimension = static_cast<int>(cfg.number("dimension", 0));
    d.max_docs_per_segment = static_cast<uint64_t>(cfg.number("max_docs_per_segment", 0));
    if (d.dimension <= 0) throw Error(KJC_INVALID_CONFIG, root + "/config.json: dimension must be positive");
    const std::string seg_root = root + "/segments";
    DIR* dp = opendir(seg_root.c_str());
    if (!dp) return d;  // `if segments_dir.exists()`: an index without segments is empty, not an error
    std::vector<std::string> names;
    while (struct dirent* e = readdir(dp)) {
        const std::string name = e->d_name;
        if (name == "." || name == "..") continue;
        struct stat sb;
        if (stat((seg_root + "/" + name).c_str(), &sb) == 0 && S_ISDIR(sb.st_mode)) names.push_back(name);
    }
    closedir(dp);
    std::sort(names.begin(), names.end());  // entries.sort_by_key(file_name)
    for (const std::string& name : names) {
        const std::string sd = seg_root + "/" + name;
        // Segment::open (segment.rs:212-238) fails -- and IndexReader::open then skips the segment with a warning --
        // when any of these is missing or segment.json does not parse.
        if (!file_exists(sd + "/segment.json") || !file_exists(sd + "/vectors.bin") || !file_exists(sd + "/docs.idx") ||
            !file_exists(sd + "/bm25.bin")) {
            ++d.skipped;
            continue;
        }
        Json meta;
        try {
            const std::string mt = read_text_file(sd + "/segment.json", KJC_LOAD_FAILED);
            meta = JsonParser(mt.data(), mt.size()).parse();
        } catch (const Error&) {
            ++d.skipped;
            continue;
        }
        if (!meta.has("id") || !meta.has("doc_count") || !meta.has("dimension") || !meta.has("created_at") || !meta.has("total_bytes")) {
            ++d.skipped;
            continue;
        }
        if (static_cast<int>(meta.number("dimension", 0)) != d.dimension)
            throw Error(KJC_LOAD_FAILED, sd + ": segment dimension differs from the index dimension (the reference would score no row of "
                                              "it, segment.rs:308-310; a mixed-dimension index is not loadable as one GPU shard)");
        IndexDirSegment s;
        s.dir = sd;
        s.doc_count = static_cast<uint64_t>(meta.number("doc_count", 0));
        s.global_base = d.total_rows;
        d.total_rows += s.doc_count;
        d.segments.push_back(std::move(s));
    }
    return d;
}

void index_part_range(uint64_t total, int part, int parts, uint64_t* lo, uint64_t* hi) {
    if (parts < 1 || part < 0 || part >= parts) throw Error(KJC_INVALID_CONFIG, "part must be in [0, parts)");
    const uint64_t base = total / parts, rem = total % parts, p = static_cast<uint64_t>(part);
    *lo = p * base + std::min<uint64_t>(p, rem);
    *hi = *lo + base + (p < rem ? 1 : 0);
}

Index* open_index_dir(const std::string& root, int device, int part, int parts) {
    const IndexDir d = scan_index_dir(root);
    uint64_t lo, hi;
    index_part_range(d.total_rows, part, parts, &lo, &hi);
    std::unique_ptr<Index> idx(new Index(d.dimension, std::max<uint64_t>(hi - lo, 1), lo, device));
    for (const IndexDirSegment& s : d.segments) {
        const uint64_t a = std::max(lo, s.global_base), b = std::min(hi, s.global_base + s.doc_count);
        if (a >= b) continue;
        idx->load_vectors_bin_range(s.dir + "/vectors.bin", a - s.global_base, b - a);
    }
    return idx.release();
}

}  // namespace kj
