// Brute-force cosine top-k scan over a row-major fp32 index shard resident in HBM.
// Replaces the per-row loops of Segment::search_vectors (reference:
// kjarni-rag/src/segment.rs:307-337,355-370) and VectorStore::search
// (kjarni-search/src/vector.rs:131-165); the per-shard top-k followed by a merge is
// the shape IndexReader::search_semantic already has (kjarni-rag/src/index_reader.rs:207-228).
//
// HBM-bound design: every index byte is read exactly once per pass of up to 8 queries;
// one warp owns one row at a time (fully coalesced 16-byte loads, several rows in
// flight per warp), queries live in registers, per-warp sorted top-k lists live in
// shared memory and are touched only when a score beats the current k-th best.
// Order everywhere is (score desc, id asc) = the reference's stable sorts.
#pragma once
#include "ptx.cuh"

namespace kj {

enum ScanMode : int {
    SCAN_SEGMENT = 0,      // segment.rs: empty result if |q| < 1e-9, score 0 for |r| < 1e-9, else dot/(|q||r|)
    SCAN_VECTORSTORE = 1,  // vector.rs: dot / max(|q||r|, 1e-9)
};

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr uint32_t kNoId32 = 0xFFFFFFFFu;
constexpr uint64_t kNoId64 = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// |row| for rows [0, n): one warp per row.
__global__ void __launch_bounds__(256) row_norm_kernel(const float* __restrict__ rows, float* __restrict__ norms, size_t n, int D) {
    const size_t row = static_cast<size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const int lane = threadIdx.x & 31;
    const float* r = rows + row * D;
    float s = 0.f;
    for (int c = lane * 4; c < D; c += 128) {
        const float4 v = ld_stream_f4(r + c);
        s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    s = warp_sum(s);
    if (lane == 0) norms[row] = sqrtf(s);
}

// Insert (s, id) into a descending list of capacity k kept in shared memory (warp-cooperative).
// Entries already present have smaller ids, so the new one goes after every entry with score >= s.
__device__ __forceinline__ void warp_topk_insert(float* sc, uint32_t* id, int k, int& n, float& thr, float s, uint32_t rid, int lane) {
    int cnt = 0;
    for (int j = lane; j < n; j += 32) cnt += (sc[j] >= s) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    const int pos = cnt;
    const int end = min(n, k - 1);  // entries [pos, end) move to [pos+1, end+1)
    for (int top = end - 1; top >= pos; top -= 32) {  // warp-uniform bounds; 32 entries per step, tail first
        const int j = top - lane;
        const bool act = j >= pos;
        float ts = 0.f;
        uint32_t ti = 0;
        if (act) { ts = sc[j]; ti = id[j]; }
        __syncwarp();
        if (act) { sc[j + 1] = ts; id[j + 1] = ti; }
        __syncwarp();
    }
    if (lane == 0) { sc[pos] = s; id[pos] = rid; }
    n = min(n + 1, k);
    __syncwarp();
    thr = (n == k) ? sc[k - 1] : -INFINITY;
}

struct ScanParams {
    const float* rows;      // [n_rows, D]
    const float* norms;     // [n_rows]
    const float* queries;   // [Q, D]
    const float* qnorms;    // [Q]
    float* out_scores;      // [gridDim.x, Q, k]  per-CTA sorted candidates
    uint32_t* out_ids;      // [gridDim.x, Q, k]  local row index, kNoId32 = empty
    size_t n_rows;
    int D, Q, k, q0, mode;  // q0: first query of this pass
};

// QT queries per pass, NCH = ceil(D/128) float4 chunks per lane, RU rows in flight per warp.
template <int QT, int NCH, int RU>
__global__ void __launch_bounds__(kScanThreads) scan_topk_kernel(ScanParams p) {
    extern __shared__ __align__(16) uint8_t smem_scan[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = p.k, D = p.D;
    // per-warp lists: [warp][QT][k]
    float* lsc = reinterpret_cast<float*>(smem_scan) + (static_cast<size_t>(warp) * QT) * k;
    uint32_t* lid = reinterpret_cast<uint32_t*>(smem_scan + static_cast<size_t>(kScanWarps) * QT * k * 4) + (static_cast<size_t>(warp) * QT) * k;
    const int nq = min(QT, p.Q - p.q0);

    float4 q[QT][NCH];
    float qn[QT];
#pragma unroll
    for (int t = 0; t < QT; ++t) {
        const int qi = p.q0 + min(t, nq - 1);
        qn[t] = p.qnorms[qi];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int col = (lane + 32 * c) * 4;
            q[t][c] = col < D ? *reinterpret_cast<const float4*>(p.queries + static_cast<size_t>(qi) * D + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    int cnt[QT];
    float thr[QT];
#pragma unroll
    for (int t = 0; t < QT; ++t) { cnt[t] = 0; thr[t] = -INFINITY; }

    const size_t gw = static_cast<size_t>(blockIdx.x) * kScanWarps + warp;
    const size_t nw = static_cast<size_t>(gridDim.x) * kScanWarps;
    for (size_t r0 = gw * RU; r0 < p.n_rows; r0 += nw * RU) {
        float4 v[RU][NCH];
        float rn[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const size_t r = min(r0 + u, p.n_rows - 1);
            const float* rp = p.rows + r * D;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int col = (lane + 32 * c) * 4;
                v[u][c] = col < D ? ld_stream_f4(rp + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            rn[u] = __ldg(p.norms + r);
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            if (r0 + u >= p.n_rows) break;
            float d[QT];
#pragma unroll
            for (int t = 0; t < QT; ++t) {
                float a = 0.f;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    a = fmaf(v[u][c].x, q[t][c].x, a);
                    a = fmaf(v[u][c].y, q[t][c].y, a);
                    a = fmaf(v[u][c].z, q[t][c].z, a);
                    a = fmaf(v[u][c].w, q[t][c].w, a);
                }
                d[t] = a;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int t = 0; t < QT; ++t) d[t] += __shfl_xor_sync(0xffffffffu, d[t], o);
            }
#pragma unroll
            for (int t = 0; t < QT; ++t) {
                if (t >= nq) break;
                float s;
                if (p.mode == SCAN_SEGMENT) s = rn[u] < 1e-9f ? 0.0f : d[t] / (qn[t] * rn[u]);
                else s = d[t] / fmaxf(qn[t] * rn[u], 1e-9f);
                if (s > thr[t]) warp_topk_insert(lsc + t * k, lid + t * k, k, cnt[t], thr[t], s, static_cast<uint32_t>(r0 + u), lane);
            }
        }
    }
    // publish list sizes, then merge the 8 per-warp lists of each query into the CTA's output list
    __shared__ int s_cnt[kScanWarps][QT];
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < QT; ++t) s_cnt[warp][t] = cnt[t];
    }
    __syncthreads();
    float* all_sc = reinterpret_cast<float*>(smem_scan);
    uint32_t* all_id = reinterpret_cast<uint32_t*>(smem_scan + static_cast<size_t>(kScanWarps) * QT * k * 4);
    for (int t = warp; t < nq; t += kScanWarps) {
        // lanes 0..7 each own the head of one warp's list; k rounds of an 8-way argmax
        int head = 0;
        const int mycnt = lane < kScanWarps ? s_cnt[lane][t] : 0;
        const float* msc = all_sc + (static_cast<size_t>(lane < kScanWarps ? lane : 0) * QT + t) * k;
        const uint32_t* mid = all_id + (static_cast<size_t>(lane < kScanWarps ? lane : 0) * QT + t) * k;
        float* os = p.out_scores + (static_cast<size_t>(blockIdx.x) * p.Q + p.q0 + t) * k;
        uint32_t* oi = p.out_ids + (static_cast<size_t>(blockIdx.x) * p.Q + p.q0 + t) * k;
        for (int j = 0; j < k; ++j) {
            float s = -INFINITY;
            uint32_t id = kNoId32;
            if (head < mycnt) { s = msc[head]; id = mid[head]; }
            float bs = s;
            uint32_t bi = id;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                const float os2 = __shfl_xor_sync(0xffffffffu, bs, o);
                const uint32_t oi2 = __shfl_xor_sync(0xffffffffu, bi, o);
                if (os2 > bs || (os2 == bs && oi2 < bi)) { bs = os2; bi = oi2; }
            }
            bs = __shfl_sync(0xffffffffu, bs, 0);
            bi = __shfl_sync(0xffffffffu, bi, 0);
            if (bi != kNoId32 && bi == id && s == bs) ++head;
            if (lane == 0) { os[j] = bi == kNoId32 ? -INFINITY : bs; oi[j] = bi; }
        }
    }
}

// Merge L sorted candidate lists per query into the final top-k, order (score desc, id asc).
// Used for (a) the per-CTA lists of one shard (ids = local u32 + id_base) and (b) the
// per-rank lists after the NVLink candidate gather (ids already global u64).
struct MergeParams {
    const float* in_scores;   // [L, Q, k]
    const uint32_t* in_ids32; // [L, Q, k] or nullptr
    const uint64_t* in_ids64; // [L, Q, k] or nullptr
    uint64_t id_base;         // added to 32-bit ids
    const float* qnorms;      // [Q] or nullptr; SCAN_SEGMENT: |q| < 1e-9 -> empty result
    float* out_scores;        // [Q, k]
    uint64_t* out_ids;        // [Q, k]  kNoId64 = empty
    int* out_counts;          // [Q] or nullptr
    int L, Q, k, mode;
};
__global__ void __launch_bounds__(256) topk_merge_kernel(MergeParams p) {
    const int qi = blockIdx.x;
    const int tid = threadIdx.x;
    extern __shared__ int heads[];  // [L]
    __shared__ float ws[8];
    __shared__ uint64_t wi[8];
    __shared__ int wl[8];
    __shared__ int win_list;
    for (int l = tid; l < p.L; l += 256) heads[l] = 0;
    __syncthreads();
    const bool empty_query = p.mode == SCAN_SEGMENT && p.qnorms != nullptr && p.qnorms[qi] < 1e-9f;
    int produced = 0;
    for (int j = 0; j < p.k; ++j) {
        float bs = -INFINITY;
        uint64_t bi = kNoId64;
        int bl = -1;
        if (!empty_query) {
            for (int l = tid; l < p.L; l += 256) {
                const int h = heads[l];
                if (h < p.k) {
                    const size_t off = (static_cast<size_t>(l) * p.Q + qi) * p.k + h;
                    uint64_t id;
                    if (p.in_ids64) id = p.in_ids64[off];
                    else { const uint32_t i32 = p.in_ids32[off]; id = i32 == kNoId32 ? kNoId64 : p.id_base + i32; }
                    if (id != kNoId64) {
                        const float s = p.in_scores[off];
                        if (bl < 0 || s > bs || (s == bs && id < bi)) { bs = s; bi = id; bl = l; }
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const uint64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
            if (ol >= 0 && (bl < 0 || os > bs || (os == bs && oi < bi))) { bs = os; bi = oi; bl = ol; }
        }
        if ((tid & 31) == 0) { ws[tid >> 5] = bs; wi[tid >> 5] = bi; wl[tid >> 5] = bl; }
        __syncthreads();
        if (tid == 0) {
            float fs = -INFINITY; uint64_t fi = kNoId64; int fl = -1;
            for (int w = 0; w < 8; ++w)
                if (wl[w] >= 0 && (fl < 0 || ws[w] > fs || (ws[w] == fs && wi[w] < fi))) { fs = ws[w]; fi = wi[w]; fl = wl[w]; }
            win_list = fl;
            p.out_scores[static_cast<size_t>(qi) * p.k + j] = fl >= 0 ? fs : -INFINITY;
            p.out_ids[static_cast<size_t>(qi) * p.k + j] = fl >= 0 ? fi : kNoId64;
            if (fl >= 0) heads[fl] += 1;
        }
        __syncthreads();
        if (win_list >= 0) ++produced;
    }
    if (tid == 0 && p.out_counts) p.out_counts[qi] = produced;
}

// Counter-based synthetic rows shared with the CPU oracle (oracle/kjarni_oracle.py:hash32/synth_rows):
// x[r,c] = (hash32(seed, r, c) >> 8) * 2^-24 - 0.5.  Bench/test input generator, not part of the search path.
__device__ __forceinline__ uint32_t hash32(uint32_t seed, uint64_t r, uint32_t c) {
    uint32_t x = static_cast<uint32_t>(r * 0x9E3779B1ull + static_cast<uint64_t>(c) * 0x85EBCA77ull + static_cast<uint64_t>(seed) * 0xC2B2AE3Dull);
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}
__global__ void synth_rows_kernel(float* __restrict__ out, uint32_t seed, uint64_t row0, size_t n, int D) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n * D) return;
    const uint64_t r = row0 + i / D;
    const uint32_t c = static_cast<uint32_t>(i % D);
    out[i] = static_cast<float>(hash32(seed, r, c) >> 8) * 5.9604644775390625e-08f - 0.5f;
}

}  // namespace kj
