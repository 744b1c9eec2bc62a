// Index shard handle behind KjcIndex: fp32 rows resident in HBM, row norms cached at
// append time, brute-force cosine top-k search and the candidate merge.
// Host-side mirror of Segment (kjarni-rag/src/segment.rs:195-337), VectorStore
// (kjarni-search/src/vector.rs:5-165) and the per-segment -> merge shape of
// IndexReader::search_semantic (kjarni-rag/src/index_reader.rs:207-228,313-319).
#include <algorithm>
#include <mutex>

#include "index.hpp"
#include "scan.cuh"

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
}

Index::~Index() {
    cudaSetDevice(device_);
    for (void* p : {(void*)rows_, (void*)norms_, (void*)d_q_, (void*)d_qn_, (void*)d_cand_s_, (void*)d_cand_i_, (void*)d_out_s_,
                    (void*)d_out_i_, (void*)d_out_c_})
        if (p) cudaFree(p);
    if (h_stage_) cudaFreeHost(h_stage_);
    if (stream_) cudaStreamDestroy(stream_);
}

void Index::compute_norms(uint64_t row0, uint64_t n, cudaStream_t st) {
    if (n == 0) return;
    row_norm_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, st>>>(rows_ + row0 * dim_, norms_ + row0, n, dim_);
    KJ_CUDA(cudaGetLastError());
}

void Index::add_rows_host(const float* rows, uint64_t n) {
    std::lock_guard<std::mutex> lock(mu_);
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

void Index::load_vectors_bin(const std::string& path) {
    // vectors.bin = raw little-endian f32 [doc_count x dim] (kjarni-rag/src/segment.rs:87-123,211-262)
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) throw Error(KJC_MODEL_NOT_FOUND, "cannot open " + path);
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); throw Error(KJC_LOAD_FAILED, "cannot stat " + path); }
    const size_t bytes = static_cast<size_t>(sb.st_size);
    if (bytes % (dim_ * sizeof(float)) != 0) { close(fd); throw Error(KJC_LOAD_FAILED, "vectors.bin size is not a multiple of 4*dim: " + path); }
    const uint64_t n = bytes / (dim_ * sizeof(float));
    if (n == 0) { close(fd); return; }
    void* map = mmap(nullptr, bytes, PROT_READ, MAP_PRIVATE, fd, 0);
    if (map == MAP_FAILED) { close(fd); throw Error(KJC_LOAD_FAILED, "mmap failed for " + path); }
    try {
        add_rows_host(static_cast<const float*>(map), n);
    } catch (...) {
        munmap(map, bytes);
        close(fd);
        throw;
    }
    munmap(map, bytes);
    close(fd);
}

void Index::append_synthetic(uint32_t seed, uint64_t row0, uint64_t n) {
    std::lock_guard<std::mutex> lock(mu_);
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
    if (row + n > len_) throw Error(KJC_INVALID_CONFIG, "Document ID out of range");
    KJ_CUDA(cudaSetDevice(device_));
    KJ_CUDA(cudaMemcpy(out, rows_ + row * dim_, n * dim_ * sizeof(float), cudaMemcpyDeviceToHost));
}

template <int QT, int NCH>
static void launch_scan_inst(ScanParams p, int grid, cudaStream_t st) {
    static int configured[64] = {0};
    // rows per stage / stage count: as much as fits in ~200 KB next to the top-k lists
    p.rows_per_stage = p.D <= 512 ? 32 : 16;
    const size_t budget = 200 * 1024 - scan_list_bytes(QT, p.k) - 512;
    p.nstages = static_cast<int>(std::min<size_t>(4, budget / scan_stage_bytes(p.D, p.rows_per_stage)));
    if (p.nstages < 2) throw Error(KJC_INVALID_CONFIG, "index dimension too large for the scan pipeline");
    const size_t smem = scan_smem_bytes(p.D, p.rows_per_stage, p.nstages, QT, p.k);
    auto kern = scan_topk_kernel<QT, NCH>;
    ensure_smem_attr(kern, static_cast<int>(smem), configured);
    kern<<<grid, kScanCtaThreads, smem, st>>>(p);
    KJ_CUDA(cudaGetLastError());
}
template <int QT>
static void launch_scan_qt(const ScanParams& p, int grid, cudaStream_t st) {
    const int nch = (p.D + 127) / 128;
    switch (nch) {
        case 1: launch_scan_inst<QT, 1>(p, grid, st); break;
        case 2: launch_scan_inst<QT, 2>(p, grid, st); break;
        case 3: launch_scan_inst<QT, 3>(p, grid, st); break;
        case 4: launch_scan_inst<QT, 4>(p, grid, st); break;
        case 5: case 6: launch_scan_inst<QT, 6>(p, grid, st); break;
        default: launch_scan_inst<QT, 8>(p, grid, st); break;
    }
}

static void launch_topk_merge(const MergeParams& m, cudaStream_t st) {
    topk_merge_kernel<<<m.Q, 256, static_cast<size_t>(m.L) * sizeof(int), st>>>(m);
    KJ_CUDA(cudaGetLastError());
}

// Enqueue: query norms -> scan passes of <= 8 queries -> merge of the per-CTA lists.
void Index::search_device(const float* d_q, int nq, int k, int mode, uint64_t* d_ids, float* d_scores, int32_t* d_counts, cudaStream_t st) {
    if (nq < 1) throw Error(KJC_INVALID_CONFIG, "nq must be >= 1");
    if (k < 1 || k > 256) throw Error(KJC_INVALID_CONFIG, "k must be in [1, 256]");
    if (mode != SCAN_SEGMENT && mode != SCAN_VECTORSTORE) throw Error(KJC_INVALID_CONFIG, "unknown scan mode");
    std::lock_guard<std::mutex> lock(mu_);
    KJ_CUDA(cudaSetDevice(device_));
    if (!st) st = stream_;
    launches_ = 0;
    // queries per warp: bounded by the per-warp list memory (k) and the register file (dim)
    int qt_max = k > 64 ? 1 : (k > 32 ? 2 : 4);
    if (dim_ > 512) qt_max = std::min(qt_max, 2);
    const int grid = std::max<int>(1, static_cast<int>(std::min<uint64_t>(num_sms_, (len_ + 31) / 32)));  // one CTA per SM
    const size_t cand = static_cast<size_t>(grid) * nq * k;
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
    m.out_scores = d_scores; m.out_ids = d_ids; m.out_counts = d_counts; m.L = len_ > 0 ? grid : 0; m.Q = nq; m.k = k; m.mode = mode;
    launch_topk_merge(m, st);
    ++launches_;
}

void Index::search_host(const float* q, int nq, int k, int mode, uint64_t* ids, float* scores, int32_t* counts) {
    if (nq < 1) throw Error(KJC_INVALID_CONFIG, "nq must be >= 1");
    if (k < 1 || k > 256) throw Error(KJC_INVALID_CONFIG, "k must be in [1, 256]");
    KJ_CUDA(cudaSetDevice(device_));
    const size_t qe = static_cast<size_t>(nq) * dim_, oe = static_cast<size_t>(nq) * k;
    {
        std::lock_guard<std::mutex> lock(mu_);
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
    search_device(d_q_, nq, k, mode, d_out_i_, d_out_s_, d_out_c_, stream_);
    KJ_CUDA(cudaMemcpyAsync(ids, d_out_i_, oe * 8, cudaMemcpyDeviceToHost, stream_));
    KJ_CUDA(cudaMemcpyAsync(scores, d_out_s_, oe * 4, cudaMemcpyDeviceToHost, stream_));
    if (counts) KJ_CUDA(cudaMemcpyAsync(counts, d_out_c_, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, stream_));
    KJ_CUDA(cudaStreamSynchronize(stream_));
}

void merge_lists_u64(const uint64_t* d_ids, const float* d_scores, int n_lists, int nq, int k, uint64_t* d_out_ids, float* d_out_scores,
                     int32_t* d_out_counts, cudaStream_t st) {
    MergeParams m;
    m.in_scores = d_scores; m.in_ids32 = nullptr; m.in_ids64 = d_ids; m.id_base = 0; m.qnorms = nullptr;
    m.out_scores = d_out_scores; m.out_ids = d_out_ids; m.out_counts = d_out_counts; m.L = n_lists; m.Q = nq; m.k = k;
    m.mode = SCAN_VECTORSTORE;
    launch_topk_merge(m, st);
}

}  // namespace kj
