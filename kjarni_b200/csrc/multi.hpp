// Multi-GPU inside ONE process, behind the C ABI (include/kjarni_cuda.h: kjc_encoder_create_multi, kjc_sharded_index_*).
// The consumer of this library is Rust / C# / Go (kjarni-ffi/src/lib.rs:24-45): there is no torchrun there, so the two ways the
// hot path shards are also available without torch.distributed:
//   * EncoderGroup  -- one full weight replica per GPU, one persistent host thread + stream + pinned staging slice per GPU; a
//                      batch is split by contiguous rows, NO collective (sequences are independent: traits.rs:66-139);
//   * ShardedIndex  -- contiguous row shards (global id = shard base + local id, index_reader.rs:313-319); every GPU scans its
//                      shard for the same query batch (always exact: unproven queries are re-run on the exact scan), the per-shard
//                      [Q,k] candidates travel to GPU 0 as peer copies over NVLink (Q*k*12 B per shard) and one merge kernel does
//                      the concat -> sort -> truncate of IndexReader::search_semantic (index_reader.rs:207-228).
#pragma once
#include <condition_variable>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "index.hpp"

namespace kj {

// n persistent host threads; run(fn) executes fn(p) on thread p for every p and returns when all are done (first exception rethrown).
class DeviceWorkers {
  public:
    explicit DeviceWorkers(int n) : n_(n), pending_(n, false), errors_(n) {
        for (int p = 1; p < n; ++p) threads_.emplace_back([this, p] { loop(p); });  // slot 0 runs on the calling thread
    }
    ~DeviceWorkers() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (std::thread& t : threads_) t.join();
    }
    int size() const { return n_; }
    void run(const std::function<void(int)>& fn) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn;
            for (int p = 1; p < n_; ++p) pending_[p] = true;
            left_ = n_ - 1;
        }
        cv_.notify_all();
        std::exception_ptr first;
        try {
            fn(0);
        } catch (...) {
            first = std::current_exception();
        }
        {
            std::unique_lock<std::mutex> lk(mu_);
            done_.wait(lk, [this] { return left_ == 0; });
            fn_ = nullptr;
            for (int p = 1; p < n_ && !first; ++p)
                if (errors_[p]) first = errors_[p];
            for (auto& e : errors_) e = nullptr;
        }
        if (first) std::rethrow_exception(first);
    }

  private:
    void loop(int p) {
        for (;;) {
            const std::function<void(int)>* fn;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || pending_[p]; });
                if (stop_) return;
                pending_[p] = false;
                fn = fn_;
            }
            std::exception_ptr err;
            try {
                (*fn)(p);
            } catch (...) {
                err = std::current_exception();
            }
            {
                std::lock_guard<std::mutex> lk(mu_);
                errors_[p] = err;
                --left_;
            }
            done_.notify_one();
        }
    }
    int n_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::vector<char> pending_;
    std::vector<std::exception_ptr> errors_;
    const std::function<void(int)>* fn_ = nullptr;
    int left_ = 0;
    bool stop_ = false;
    std::vector<std::thread> threads_;
};

inline void split_rows(uint64_t total, int part, int parts, uint64_t* lo, uint64_t* hi) { index_part_range(total, part, parts, lo, hi); }

inline std::vector<int> checked_devices(const int* device_ids, int n) {
    if (!device_ids || n < 1) throw Error(KJC_INVALID_CONFIG, "device list must name at least one device");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) throw Error(KJC_GPU_UNAVAILABLE, "no CUDA device available");
    std::vector<int> d(device_ids, device_ids + n);
    // a device may be listed more than once (several replicas / shards on one GPU): no use in production, but it lets a one-GPU box
    // run the split, gather and merge logic
    for (int i = 0; i < n; ++i)
        if (d[i] < 0 || d[i] >= ndev) throw Error(KJC_GPU_UNAVAILABLE, "device index out of range");
    return d;
}

// One encoder replica per device; a host-buffer forward is split by contiguous rows over the replicas.
class EncoderGroup {
  public:
    EncoderGroup(const std::string& model_dir, const int* device_ids, int n) : devs_(checked_devices(device_ids, n)), workers_(n) {
        reps_.resize(n);
        workers_.run([&](int p) { reps_[p].reset(new Encoder(model_dir, devs_[p])); });  // replicas load in parallel
    }
    int size() const { return static_cast<int>(reps_.size()); }
    Encoder& replica(int p) { return *reps_[p]; }
    const Encoder& replica(int p) const { return *reps_[p]; }
    int64_t last_launches() const {
        int64_t s = 0;
        for (const auto& r : reps_) s += r->last_launches();
        return s;
    }
    // Same contract as Encoder::forward_host.  The padding-mask convention is resolved once for the WHOLE batch
    // (ComputeStrategy::select looks at batch x seq, strategy.rs:29-47), not per slice, so the split is invisible in the results.
    void forward_host(const uint32_t* ids, const float* mask, const uint32_t* types, int B, int S, const KjcForwardOptions& o, float* out,
                      const Encoder::RowSink* sink = nullptr) {
        const int n = size();
        if (n == 1 || B < 2) {
            reps_[0]->forward_host(ids, mask, types, B, S, o, out, sink);
            return;
        }
        std::lock_guard<std::mutex> lock(mu_);
        const KjcForwardOptions oc = reps_[0]->resolved_options(B, S, o);  // also validates
        const size_t row = reps_[0]->out_row_elems(oc, S);
        std::mutex sink_mu;
        workers_.run([&](int p) {
            uint64_t lo, hi;
            split_rows(static_cast<uint64_t>(B), p, n, &lo, &hi);
            if (hi == lo) return;
            const size_t t0 = static_cast<size_t>(lo) * S;
            const Encoder::RowSink wrapped = [&](const float* rows, size_t first, size_t cnt) {
                std::lock_guard<std::mutex> lk(sink_mu);  // sinks are written for one caller thread
                (*sink)(rows, static_cast<size_t>(lo) + first, cnt);
            };
            reps_[p]->forward_host(ids + t0, mask ? mask + t0 : nullptr, types ? types + t0 : nullptr, static_cast<int>(hi - lo), S, oc,
                                   out ? out + lo * row : nullptr, sink ? &wrapped : nullptr);
        });
    }

  private:
    std::vector<int> devs_;
    DeviceWorkers workers_;
    std::vector<std::unique_ptr<Encoder>> reps_;
    std::mutex mu_;
};

// Row-sharded index over several devices of one process.
class ShardedIndex {
  public:
    // Empty shards: shard p will hold global rows [lo_p, hi_p) of `capacity_rows` (contiguous, sizes differ by at most one row).
    ShardedIndex(int dim, uint64_t capacity_rows, const int* device_ids, int n) : dim_(dim), cap_(capacity_rows), devs_(checked_devices(device_ids, n)), workers_(n) {
        if (capacity_rows < 1) throw Error(KJC_INVALID_CONFIG, "capacity must be positive");
        shards_.resize(n);
        workers_.run([&](int p) {
            uint64_t lo, hi;
            split_rows(cap_, p, n, &lo, &hi);
            shards_[p].reset(new Index(dim, std::max<uint64_t>(hi - lo, 1), lo, devs_[p]));
        });
        init_buffers();
    }
    // IndexReader::open split by rows: shard p uploads part p of the on-disk index (same split as kjc_index_open_dir(root, dev, p, n)).
    ShardedIndex(const std::string& root, const int* device_ids, int n) : devs_(checked_devices(device_ids, n)), workers_(n) {
        const IndexDir d = scan_index_dir(root);
        dim_ = d.dimension;
        cap_ = std::max<uint64_t>(d.total_rows, 1);
        shards_.resize(n);
        workers_.run([&](int p) { shards_[p].reset(open_index_dir(root, devs_[p], p, n)); });
        len_ = d.total_rows;
        init_buffers();
    }
    ~ShardedIndex() {
        for (size_t p = 0; p < bufs_.size(); ++p) {
            cudaSetDevice(devs_[p]);
            Buf& b = bufs_[p];
            for (void* q : {(void*)b.d_q, (void*)b.d_rec}) if (q) cudaFree(q);
            if (b.done) cudaEventDestroy(b.done);
        }
        cudaSetDevice(devs_[0]);
        for (void* q : {(void*)g_rec_, (void*)o_ids_, (void*)o_sc_, (void*)o_cnt_}) if (q) cudaFree(q);
        if (h_out_) cudaFreeHost(h_out_);
        if (stream0_) cudaStreamDestroy(stream0_);
    }
    int n_shards() const { return static_cast<int>(shards_.size()); }
    int dim() const { return dim_; }
    uint64_t len() const { return len_; }
    uint64_t capacity() const { return cap_; }
    Index& shard(int p) { return *shards_[p]; }

    // Rows are appended in global-id order: row g goes to the shard whose range holds g.
    void add_rows_host(const float* rows, uint64_t n) {
        std::lock_guard<std::mutex> lock(mu_);
        if (len_ + n > cap_) throw Error(KJC_INVALID_CONFIG, "index capacity exceeded");
        const uint64_t g0 = len_, g1 = len_ + n;
        const int ns = n_shards();
        workers_.run([&](int p) {
            uint64_t lo, hi;
            split_rows(cap_, p, ns, &lo, &hi);
            const uint64_t a = std::max(lo, g0), b = std::min(hi, g1);
            if (a < b) shards_[p]->add_rows_host(rows + (a - g0) * dim_, b - a);
        });
        len_ = g1;
    }
    // The next `n` rows of the deterministic synthetic sequence (row g = synth row g of `seed`), generated on each shard's device.
    void append_synthetic(uint32_t seed, uint64_t n) {
        std::lock_guard<std::mutex> lock(mu_);
        if (len_ + n > cap_) throw Error(KJC_INVALID_CONFIG, "index capacity exceeded");
        const uint64_t g0 = len_, g1 = len_ + n;
        const int ns = n_shards();
        workers_.run([&](int p) {
            uint64_t lo, hi;
            split_rows(cap_, p, ns, &lo, &hi);
            const uint64_t a = std::max(lo, g0), b = std::min(hi, g1);
            if (a < b) shards_[p]->append_synthetic(seed, a, b - a);
        });
        len_ = g1;
    }

    // Host queries [nq, dim] -> global top-k (ids u64 [nq,k] with UINT64_MAX where empty, scores, optional counts); exact.
    void search_host(const float* q, int nq, int k, int mode, uint64_t* ids, float* scores, int32_t* counts) {
        if (nq < 1) throw Error(KJC_INVALID_CONFIG, "nq must be >= 1");
        if (k < 1 || k > 256) throw Error(KJC_INVALID_CONFIG, "k must be in [1, 256]");
        const int ns = n_shards();
        if (ns == 1) {
            shards_[0]->search_host(q, nq, k, mode, ids, scores, counts);
            return;
        }
        std::lock_guard<std::mutex> lock(mu_);
        const size_t qe = static_cast<size_t>(nq) * dim_, oe = static_cast<size_t>(nq) * k;
        const size_t rec = packed_record_bytes(nq, k);  // one shard's candidates: ids | scores, one peer copy
        KJ_CUDA(cudaSetDevice(devs_[0]));
        if (oe > out_cap_) {
            for (void* p : {(void*)g_rec_, (void*)o_ids_, (void*)o_sc_, (void*)o_cnt_}) if (p) cudaFree(p);
            if (h_out_) cudaFreeHost(h_out_);
            KJ_CUDA(cudaMalloc(&g_rec_, rec * ns));
            KJ_CUDA(cudaMalloc(&o_ids_, oe * 8));
            KJ_CUDA(cudaMalloc(&o_sc_, oe * 4));
            KJ_CUDA(cudaMalloc(&o_cnt_, (oe + 1) * 4));  // nq <= oe
            KJ_CUDA(cudaMallocHost(&h_out_, oe * 16 + 16));
            out_cap_ = oe;
        }
        workers_.run([&](int p) {
            KJ_CUDA(cudaSetDevice(devs_[p]));
            Buf& b = bufs_[p];
            if (qe > b.q_cap) {
                if (b.d_q) cudaFree(b.d_q);
                KJ_CUDA(cudaMalloc(&b.d_q, qe * 4));
                b.q_cap = qe;
            }
            if (rec > b.o_cap) {
                if (b.d_rec) cudaFree(b.d_rec);
                KJ_CUDA(cudaMalloc(&b.d_rec, rec));
                b.o_cap = rec;
            }
            uint64_t* d_ids = reinterpret_cast<uint64_t*>(b.d_rec);
            float* d_sc = reinterpret_cast<float*>(b.d_rec + oe * 8);
            cudaStream_t st = shards_[p]->stream();
            // straight from the caller's buffer, like Index::search_host: each device's thread stages its own copy concurrently
            KJ_CUDA(cudaMemcpyAsync(b.d_q, q, qe * 4, cudaMemcpyHostToDevice, st));
            // per-shard top-k, proven exact (queries the tensor-core filter cannot prove are re-run on the exact scan before this returns)
            shards_[p]->search_device(b.d_q, nq, k, mode, d_ids, d_sc, nullptr, st, /*may_sync=*/true);
            // candidate gather: this shard's packed record into slot p of GPU 0's buffer (ONE NVLink peer copy, 12 B per candidate)
            KJ_CUDA(cudaMemcpyPeerAsync(g_rec_ + p * rec, devs_[0], b.d_rec, devs_[p], rec, st));
            KJ_CUDA(cudaEventRecord(b.done, st));
        });
        KJ_CUDA(cudaSetDevice(devs_[0]));
        for (int p = 0; p < ns; ++p) KJ_CUDA(cudaStreamWaitEvent(stream0_, bufs_[p].done, 0));
        merge_lists_u64(reinterpret_cast<const uint64_t*>(g_rec_), reinterpret_cast<const float*>(g_rec_ + oe * 8), ns, nq, k, o_ids_, o_sc_, o_cnt_,
                        stream0_, rec / 8, rec / 4);
        uint8_t* h = static_cast<uint8_t*>(h_out_);
        KJ_CUDA(cudaMemcpyAsync(h, o_ids_, oe * 8, cudaMemcpyDeviceToHost, stream0_));
        KJ_CUDA(cudaMemcpyAsync(h + oe * 8, o_sc_, oe * 4, cudaMemcpyDeviceToHost, stream0_));
        KJ_CUDA(cudaMemcpyAsync(h + oe * 12, o_cnt_, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, stream0_));
        KJ_CUDA(cudaStreamSynchronize(stream0_));
        memcpy(ids, h, oe * 8);
        memcpy(scores, h + oe * 8, oe * 4);
        if (counts) memcpy(counts, h + oe * 12, static_cast<size_t>(nq) * 4);
    }

  private:
    struct Buf {
        float* d_q = nullptr;
        uint8_t* d_rec = nullptr;  // [nq,k] u64 ids | [nq,k] f32 scores
        size_t q_cap = 0, o_cap = 0;
        cudaEvent_t done = nullptr;
    };
    void init_buffers() {
        const int ns = n_shards();
        bufs_.resize(ns);
        for (int p = 0; p < ns; ++p) {
            KJ_CUDA(cudaSetDevice(devs_[p]));
            KJ_CUDA(cudaEventCreateWithFlags(&bufs_[p].done, cudaEventDisableTiming));
            for (int q = 0; q < ns; ++q) {  // direct NVLink peer copies where the topology allows; otherwise the copies stage through the host
                int can = 0;
                if (devs_[q] != devs_[p] && cudaDeviceCanAccessPeer(&can, devs_[p], devs_[q]) == cudaSuccess && can) {
                    const cudaError_t e = cudaDeviceEnablePeerAccess(devs_[q], 0);
                    if (e != cudaSuccess) cudaGetLastError();  // already enabled by another handle
                }
            }
        }
        KJ_CUDA(cudaSetDevice(devs_[0]));
        KJ_CUDA(cudaStreamCreateWithFlags(&stream0_, cudaStreamNonBlocking));
    }
    int dim_ = 0;
    uint64_t cap_ = 0, len_ = 0;
    std::vector<int> devs_;
    DeviceWorkers workers_;
    std::vector<std::unique_ptr<Index>> shards_;
    std::vector<Buf> bufs_;
    std::mutex mu_;
    cudaStream_t stream0_ = nullptr;
    uint8_t* g_rec_ = nullptr;  // n_shards packed records on device 0
    uint64_t* o_ids_ = nullptr;
    float* o_sc_ = nullptr;
    int32_t* o_cnt_ = nullptr;
    void* h_out_ = nullptr;
    size_t out_cap_ = 0;
};

}  // namespace kj
