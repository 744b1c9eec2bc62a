// Brute-force cosine top-k scan over a row-major fp32 index shard resident in HBM.
// Replaces the per-row loops of Segment::search_vectors (reference:
// kjarni-rag/src/segment.rs:307-337,355-370) and VectorStore::search
// (kjarni-search/src/vector.rs:131-165); the per-shard top-k followed by a merge is
// the shape IndexReader::search_semantic already has (kjarni-rag/src/index_reader.rs:207-228).
//
// HBM-bound design: every index byte is read exactly once per pass of up to 8 queries.
// Rows are streamed through a ring of shared-memory stages by the TMA engine (cp.async.bulk,
// ~190 KB in flight per SM, independent of register pressure); 8 consumer warps own one row
// at a time, queries live in registers, per-warp sorted top-k lists live in shared memory
// and are touched only when a score beats the current k-th best.
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
    const float* norms;     // [n_rows] (+16 B slack: the tail chunk's bulk copy is rounded up to 16 B)
    const float* queries;   // [Q, D]
    const float* qnorms;    // [Q]
    float* out_scores;      // [gridDim.x, Q, k]  per-CTA sorted candidates
    uint32_t* out_ids;      // [gridDim.x, Q, k]  local row index, kNoId32 = empty
    size_t n_rows;
    int D, Q, k, q0, mode;  // q0: first query of this pass
    int rows_per_stage, nstages;
    int ngroups;            // 1 or 2 warp groups; group g scores queries q0 + g*QT .. q0 + g*QT + QT - 1
};

constexpr int kScanConsumerWarps = 16;
constexpr int kScanCtaThreads = (kScanConsumerWarps + 1) * 32;  // + 1 producer warp

// 1-D bulk async copy global -> shared (TMA engine, no tensor map), completion counted in bytes on an mbarrier.
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :
                 : "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

inline size_t scan_stage_bytes(int D, int rows_per_stage) { return static_cast<size_t>(rows_per_stage) * D * 4 + ((rows_per_stage * 4 + 15) & ~15); }
inline size_t scan_list_bytes(int QT, int k, int consumer_warps = kScanConsumerWarps) { return static_cast<size_t>(consumer_warps) * QT * k * 8; }
inline size_t scan_smem_bytes(int D, int rows_per_stage, int nstages, int QT, int k, int consumer_warps = kScanConsumerWarps) {
    return 128 /*align*/ + nstages * scan_stage_bytes(D, rows_per_stage) + 2 * nstages * 8 + scan_list_bytes(QT, k, consumer_warps);
}

// QT queries per warp (1/2/4/8), NCH = ceil(D/128) float4 chunks per lane, CW consumer warps, RU rows per warp per turn.
// One CTA per SM: a producer thread streams chunks of `rows_per_stage` consecutive rows (+ their cached norms) through
// a ring of shared-memory stages with cp.async.bulk (~190 KB in flight per SM); 16 consumer warps in 1 or 2 groups take
// one row each per turn from the landed stage (with 2 groups, both read every row and score different queries).
// Dot products run on the packed FFMA2 pipe; the QT sums of a row are reduced across the warp with a transposed
// butterfly, so the warp ends up with one finished score per 32/QT lanes and does one divide per row.
//
// The 8-query pass at dim <= 384 runs as <8, NCH, 8, 2>: ONE group of 8 warps scores all 8 queries of a row (round 1 ran two groups
// of 4 queries, i.e. every row was fetched from shared memory and reduced twice: 232 issue slots per row, 57 % issue utilisation,
// 43 % of HBM in ncu), the queries take 96 registers per thread (hence 8 warps, not 16), and every warp works on two rows at a
// time so that the two dependent chains (FFMA2 -> butterfly -> divide) overlap.
template <int QT, int NCH, int CW = kScanConsumerWarps, int RU = 1>
__global__ void __launch_bounds__((CW + 1) * 32, 1) scan_topk_kernel(ScanParams p) {
    constexpr int LOGQ = QT == 8 ? 3 : (QT == 4 ? 2 : (QT == 2 ? 1 : 0));
    constexpr int GROUP_SHIFT = 5 - LOGQ;  // lanes per query group = 1 << GROUP_SHIFT
    extern __shared__ uint8_t smem_scan_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_scan_raw) + 127) & ~uintptr_t(127));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = p.k, D = p.D, rps = p.rows_per_stage, nst = p.nstages;
    const size_t row_bytes = static_cast<size_t>(D) * 4;
    const size_t stage_rows_bytes = static_cast<size_t>(rps) * row_bytes;
    const size_t stage_bytes = stage_rows_bytes + ((rps * 4 + 15) & ~15);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + nst * stage_bytes);
    uint64_t* empty_bar = full_bar + nst;
    uint8_t* lists = reinterpret_cast<uint8_t*>(empty_bar + nst);
    float* all_sc = reinterpret_cast<float*>(lists);
    uint32_t* all_id = reinterpret_cast<uint32_t*>(lists + static_cast<size_t>(CW) * QT * k * 4);
    __shared__ int s_cnt[CW][QT];

    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], CW);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const size_t n_chunks = (p.n_rows + rps - 1) / rps;
    const int wpg = CW / p.ngroups;  // warps per group
    const int nq_pass = min(QT * p.ngroups, p.Q - p.q0);

    if (warp == CW) {
        // ------------------------------------------------------------ producer
        if (lane == 0) {
            int it = 0;
            for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
                const int stage = it % nst;
                const uint32_t phase = (it / nst) & 1;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                const size_t row0 = c * rps;
                const uint32_t rows_in = static_cast<uint32_t>(min(static_cast<size_t>(rps), p.n_rows - row0));
                const uint32_t b_rows = rows_in * static_cast<uint32_t>(row_bytes);
                const uint32_t b_norm = (rows_in * 4 + 15) & ~15u;
                uint8_t* dst = smem + stage * stage_bytes;
                mbar_arrive_expect_tx(&full_bar[stage], b_rows + b_norm);
                bulk_load_1d(dst, p.rows + row0 * D, b_rows, &full_bar[stage]);
                bulk_load_1d(dst + stage_rows_bytes, p.norms + row0, b_norm, &full_bar[stage]);
            }
        }
    } else {
        // ----------------------------------------------------------- consumers
        const int grp = warp / wpg, wg = warp % wpg;
        const int gq0 = p.q0 + grp * QT;                       // first query of this warp's group
        const int nq = max(0, min(QT, p.Q - gq0));             // live queries of this group (0: idle group, only hand-shakes)
        float* lsc = all_sc + (static_cast<size_t>(warp) * QT) * k;
        uint32_t* lid = all_id + (static_cast<size_t>(warp) * QT) * k;
        uint64_t q[QT][NCH][2];  // packed pairs (x,y) and (z,w)
#pragma unroll
        for (int t = 0; t < QT; ++t) {
            const int qi = min(gq0 + min(t, max(nq - 1, 0)), p.Q - 1);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int col = (lane + 32 * c) * 4;
                const float4 f = col < D ? *reinterpret_cast<const float4*>(p.queries + static_cast<size_t>(qi) * D + col) : make_float4(0.f, 0.f, 0.f, 0.f);
                q[t][c][0] = f2_pack(f.x, f.y);
                q[t][c][1] = f2_pack(f.z, f.w);
            }
        }
        const int t_mine = lane >> GROUP_SHIFT;
        const bool owner = (lane & ((1 << GROUP_SHIFT) - 1)) == 0 && t_mine < nq;
        const float qn_mine = p.qnorms[min(gq0 + min(t_mine, max(nq - 1, 0)), p.Q - 1)];
        int cnt_mine = 0;
        float thr_mine = -INFINITY;

        int it = 0;
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
            const int stage = it % nst;
            const uint32_t phase = (it / nst) & 1;
            const size_t row0 = c * rps;
            const int rows_in = static_cast<int>(min(static_cast<size_t>(rps), p.n_rows - row0));
            mbar_wait(&full_bar[stage], phase);
            const uint8_t* sbase = smem + stage * stage_bytes;
            const float* snorm = reinterpret_cast<const float*>(sbase + stage_rows_bytes);
            if (nq > 0) {
                const uint32_t sbase_u32 = smem_u32(sbase);
                for (int r0 = wg; r0 < rows_in; r0 += RU * wpg) {
                    // RU rows of this warp's turn: r0, r0 + wpg, ...  (a row index >= rows_in reads stale shared memory and is discarded)
                    float sc_row[RU];
#pragma unroll
                    for (int u = 0; u < RU; ++u) {
                        const int r = r0 + u * wpg;
                        const uint32_t raddr = sbase_u32 + static_cast<uint32_t>(min(r, rps - 1)) * static_cast<uint32_t>(row_bytes);
                        ulonglong2 v[NCH];
#pragma unroll
                        for (int cc = 0; cc < NCH; ++cc) {
                            if ((lane + 32 * cc) * 4 < D) {
                                const uint4 t = ld_shared_v4(raddr + (lane + 32 * cc) * 16);
                                v[cc].x = (static_cast<unsigned long long>(t.y) << 32) | t.x;
                                v[cc].y = (static_cast<unsigned long long>(t.w) << 32) | t.z;
                            } else {
                                v[cc] = make_ulonglong2(0ull, 0ull);
                            }
                        }
                        const float rn = snorm[min(r, rps - 1)];
                        float acc[QT];
#pragma unroll
                        for (int t = 0; t < QT; ++t) {
                            uint64_t a2 = 0ull;  // (0.f, 0.f)
#pragma unroll
                            for (int cc = 0; cc < NCH; ++cc) {
                                a2 = f2_fma(v[cc].x, q[t][cc][0], a2);
                                a2 = f2_fma(v[cc].y, q[t][cc][1], a2);
                            }
                            float lo, hi;
                            f2_unpack(a2, lo, hi);
                            acc[t] = lo + hi;
                        }
#pragma unroll
                        for (int step = 0; step < 5; ++step) {
                            const int o = 16 >> step;
                            if (step < LOGQ) {
                                const bool upper = (lane & o) != 0;
                                const int half = QT >> (step + 1);
#pragma unroll
                                for (int i = 0; i < half; ++i) {
                                    const float send = upper ? acc[i] : acc[i + half];
                                    const float keep = upper ? acc[i + half] : acc[i];
                                    acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                                }
                            } else {
                                acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o);
                            }
                        }
                        float sres;
                        if (p.mode == SCAN_SEGMENT) sres = rn < 1e-9f ? 0.0f : acc[0] / (qn_mine * rn);
                        else sres = acc[0] / fmaxf(qn_mine * rn, 1e-9f);
                        sc_row[u] = sres;
                    }
#pragma unroll
                    for (int u = 0; u < RU; ++u) {  // ascending row order: equal scores keep ascending ids in the lists
                        const int r = r0 + u * wpg;
                        const float s = sc_row[u];
                        uint32_t need = __ballot_sync(0xffffffffu, owner && r < rows_in && s > thr_mine);
                        while (need) {
                            const int b = __ffs(need) - 1;
                            need &= need - 1;
                            const int t = b >> GROUP_SHIFT;
                            const float sb = __shfl_sync(0xffffffffu, s, b);
                            int n = __shfl_sync(0xffffffffu, cnt_mine, b);
                            float thr;
                            warp_topk_insert(lsc + t * k, lid + t * k, k, n, thr, sb, static_cast<uint32_t>(row0 + r), lane);
                            if (t_mine == t) { cnt_mine = n; thr_mine = thr; }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
        }
        if ((lane & ((1 << GROUP_SHIFT) - 1)) == 0 && t_mine < QT) s_cnt[warp][t_mine] = cnt_mine;
    }
    __syncthreads();
    // merge the per-warp lists of each query (the wpg warps of its group) into the CTA's output list
    for (int tq = warp; tq < nq_pass; tq += CW + 1) {
        const int grp = tq / QT, t = tq % QT;
        const int src_warp = grp * wpg + min(lane, wpg - 1);
        // lanes 0..wpg-1 each own the head of one warp's list; k rounds of a wpg-way argmax
        int head = 0;
        const int mycnt = lane < wpg ? s_cnt[src_warp][t] : 0;
        const float* msc = all_sc + (static_cast<size_t>(src_warp) * QT + t) * k;
        const uint32_t* mid = all_id + (static_cast<size_t>(src_warp) * QT + t) * k;
        float* os = p.out_scores + (static_cast<size_t>(blockIdx.x) * p.Q + p.q0 + tq) * k;
        uint32_t* oi = p.out_ids + (static_cast<size_t>(blockIdx.x) * p.Q + p.q0 + tq) * k;
        for (int j = 0; j < k; ++j) {
            float s = -INFINITY;
            uint32_t id = kNoId32;
            if (head < mycnt) { s = msc[head]; id = mid[head]; }
            float bs = s;
            uint32_t bi = id;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
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
    size_t ids_stride, scores_stride;  // elements between consecutive lists of in_ids / in_scores (0 = Q * k, densely packed); they
                                       // differ for the packed per-rank record (ids | scores) of the one-collective gather
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
                    const size_t dense = static_cast<size_t>(p.Q) * p.k, in_list = static_cast<size_t>(qi) * p.k + h;
                    const size_t off_i = static_cast<size_t>(l) * (p.ids_stride ? p.ids_stride : dense) + in_list;
                    const size_t off_s = static_cast<size_t>(l) * (p.scores_stride ? p.scores_stride : dense) + in_list;
                    uint64_t id;
                    if (p.in_ids64) id = p.in_ids64[off_i];
                    else { const uint32_t i32 = p.in_ids32[off_i]; id = i32 == kNoId32 ? kNoId64 : p.id_base + i32; }
                    if (id != kNoId64) {
                        const float s = p.in_scores[off_s];
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

namespace kj {

// ------------------------------------------------------------------------------------------------------------------------------
// Exact fp32 scan, round 2 (dim % 32 == 0): lane = row.
// The kernels above reduce every row's dot products across the 32 lanes of a warp with shuffles (9 per row for 8 queries) and are
// latency-bound on that butterfly (ncu: 41 % issue utilisation, 0.54 of the HBM copy rate).  Here a warp never reduces across lanes:
//   * rows arrive as TMA boxes of 64 rows x 32 floats with the 128-byte swizzle, so lane l reads the 16-byte chunk c of ITS rows
//     (l and l + 32) conflict-free (physical chunk c ^ (row & 7)) while the chunk index c is the same for all lanes;
//   * the <= 8 queries of the pass sit in shared memory and are read with warp-uniform (broadcast) 16-byte loads;
//   * the 8 consumer warps split the dimension: warp w owns chunk w of every box (chunk j of a row belongs to group j % 8) and keeps
//     8 packed accumulators (one per query) for its row; after the last box the 8 group partials of every (query, row) go through
//     shared memory and warp q sums them IN GROUP ORDER, divides by the norms and maintains the CTA's top-k list of query q.
// Fixed arithmetic order (scan_rescore_kernel reproduces it bit for bit): group partial = sequential FFMA2 over the group's chunks
// in ascending order, (x,y) then (z,w), lo + hi; score numerator = ((((p0 + p1) + p2) + ...) + p7).
// Issue slots per row: ~90 (was 139 + shuffle latency); shared-memory wavefronts per row: 36; HBM budget at 80 %: 83 clk per row.
struct ScanT8Params {
    const float* norms;     // [n_rows]
    const float* queries;   // [Q, D]
    const float* qnorms;    // [Q]
    float* out_scores;      // [gridDim.x, Q, k]  per-CTA sorted candidates
    uint32_t* out_ids;      // [gridDim.x, Q, k]  local row index, kNoId32 = empty
    size_t n_rows;
    int D, Q, k, q0, mode;
    int ds;                 // floats of a row per stage (multiple of 32 dividing D, <= 192): wide rows take D / ds stages per tile
    int nstages;
};
constexpr int kT8Warps = 8;         // warps of one team = chunk groups
constexpr int kT8Rows = 64;         // rows per tile: lane l owns rows l and l + 32 (the query loads are shared by both)
constexpr int kT8BoxBytes = kT8Rows * 128;
template <int TEAMS>
constexpr int t8_threads() { return (TEAMS * kT8Warps + 1) * 32; }

inline size_t scan_t8_smem_bytes(int D, int ds, int nstages, int k, int teams) {
    return 1024 /*align*/ + static_cast<size_t>(nstages) * ds * kT8Rows * 4 + 8u * D * 4 +
           static_cast<size_t>(teams) * (2u * kT8Warps * 8 * kT8Rows * 4 + 8u * k * 8) + 3u * 4 * 8 + 64;
}

__device__ __forceinline__ void tma_load_2d_plain(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :
                 : "r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// 16 bytes of shared memory as two packed fp32 pairs (no register shuffling between the load and the FFMA2s)
__device__ __forceinline__ void ld_shared_2x64(uint32_t addr, uint64_t& lo, uint64_t& hi) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(addr));
}

// TEAMS independent teams of 8 consumer warps share the stage ring: the CTA's i-th row tile goes to team i % TEAMS, which keeps its
// own partial buffers, barrier and top-k lists (the CTA writes TEAMS lists per query; the merge kernel takes them all).  Two teams
// double the warps that hide the shared-memory and barrier latencies of a tile.
template <int TEAMS>
__global__ void __launch_bounds__(t8_threads<TEAMS>(), 1) scan_t8_kernel(const __grid_constant__ CUtensorMap tmap_rows, ScanT8Params p) {
    extern __shared__ uint8_t smem_t8_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_t8_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = p.D, k = p.k, ds = p.ds, nst = p.nstages;
    const int boxes = ds / 32;                    // TMA boxes (64 rows x 128 B) per stage
    const int stage_bytes = boxes * kT8BoxBytes;  // 64 rows x ds floats
    const int n_parts = D / ds;
    uint8_t* stages = smem;
    float* s_q = reinterpret_cast<float*>(stages + static_cast<size_t>(nst) * stage_bytes);  // [8][D]
    float* s_part_all = s_q + 8 * D;                                                         // [TEAMS][2][8 groups][8 queries][64 rows]
    float* l_sc_all = s_part_all + TEAMS * 2 * kT8Warps * 8 * kT8Rows;                       // [TEAMS][8][k]
    uint32_t* l_id_all = reinterpret_cast<uint32_t*>(l_sc_all + TEAMS * 8 * k);              // [TEAMS][8][k]
    // full barriers are PER TEAM: a parity wait is only exact for an observer that sees every phase of the barrier in order, and a
    // team skips the other team's tiles (with one shared barrier per stage, team B's first wait -- parity 1 on a fresh barrier --
    // would pass before anything had landed).  The producer alone observes the empty barriers, in order, so one per stage suffices.
    uint64_t* full_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(l_id_all + TEAMS * 8 * k) + 7) & ~uintptr_t(7));  // [TEAMS][nst]
    uint64_t* empty_bar = full_bar + TEAMS * nst;                                                                                     // [nst]

    const int nq = min(8, p.Q - p.q0);  // live queries of this pass
    for (int i = threadIdx.x; i < 8 * D; i += t8_threads<TEAMS>()) {
        const int q = i / D, c = i - q * D;
        s_q[i] = p.queries[static_cast<size_t>(p.q0 + min(q, nq - 1)) * D + c];  // dead query slots repeat the last live one
    }
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_rows);
        for (int i = 0; i < nst; ++i) mbar_init(&empty_bar[i], kT8Warps);
        for (int i = 0; i < TEAMS * nst; ++i) mbar_init(&full_bar[i], 1);
        fence_mbar_init();
    }
    __syncthreads();

    const size_t n_tiles = (p.n_rows + kT8Rows - 1) / kT8Rows;
    if (warp == TEAMS * kT8Warps) {
        // ------------------------------------------------------------ producer: one stage = `boxes` TMA boxes of one row tile
        if (lane == 0) {
            int it = 0, tile_i = 0;
            for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_i) {
                uint64_t* fb = full_bar + (tile_i % TEAMS) * nst;  // the consuming team's barriers
                for (int part = 0; part < n_parts; ++part, ++it) {
                    const int stage = it % nst;
                    mbar_wait(&empty_bar[stage], ((it / nst) & 1) ^ 1);
                    mbar_arrive_expect_tx(&fb[stage], static_cast<uint32_t>(stage_bytes));
                    uint8_t* dst = stages + static_cast<size_t>(stage) * stage_bytes;
                    for (int b = 0; b < boxes; ++b)
                        tma_load_2d_plain(dst + b * kT8BoxBytes, &tmap_rows, &fb[stage], (part * boxes + b) * 32, static_cast<int32_t>(t * kT8Rows));
                }
            }
        }
    } else {
        // ----------------------------------------------------------- consumers: team = tile parity, warp = chunk group, lane = 2 rows
        const int team = warp / kT8Warps, w = warp % kT8Warps;
        const uint32_t q_u32 = smem_u32(s_q);
        const uint32_t my_chunk = (static_cast<uint32_t>(w) ^ static_cast<uint32_t>(lane & 7)) << 4;  // physical position of logical chunk `w`
        const float qn = p.qnorms[p.q0 + min(w, nq - 1)];  // reducer role: warp q of the team owns query q
        float* s_part = s_part_all + team * 2 * kT8Warps * 8 * kT8Rows;
        float* my_sc = l_sc_all + (team * 8 + w) * k;
        uint32_t* my_id = l_id_all + (team * 8 + w) * k;
        int cnt = 0;
        float thr = -INFINITY;
        uint64_t* fb = full_bar + team * nst;
        uint32_t uses = 0;       // bit s: parity of how often this team has consumed stage s = the phase parity of its barrier
        int ti = 0, tile_i = 0;  // ti: tiles of this team so far; tile_i: tiles of the CTA so far
        for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_i) {
            if (tile_i % TEAMS != team) continue;
            // the reducer needs these rows' norms at the very end of the tile: fetch them now, under the dot products
            const size_t row0 = t * kT8Rows + lane, row1 = row0 + 32;
            const bool valid0 = row0 < p.n_rows, valid1 = row1 < p.n_rows;
            const float rn0 = (w < nq && valid0) ? __ldg(p.norms + row0) : 1.0f;
            const float rn1 = (w < nq && valid1) ? __ldg(p.norms + row1) : 1.0f;
            uint64_t a0[8], a1[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) a0[q] = a1[q] = 0ull;
            for (int part = 0; part < n_parts; ++part) {
                const int it = tile_i * n_parts + part;
                const int stage = it % nst;
                mbar_wait(&fb[stage], (uses >> stage) & 1);
                uses ^= 1u << stage;
                const uint32_t rbase = smem_u32(stages + static_cast<size_t>(stage) * stage_bytes) + lane * 128 + my_chunk;
                const uint32_t qbase = q_u32 + static_cast<uint32_t>((part * ds + w * 4) * 4);
#pragma unroll 2
                for (int b = 0; b < boxes; ++b) {
                    uint64_t r0xy, r0zw, r1xy, r1zw;
                    ld_shared_2x64(rbase + b * kT8BoxBytes, r0xy, r0zw);
                    ld_shared_2x64(rbase + b * kT8BoxBytes + 32 * 128, r1xy, r1zw);  // row + 32: same swizzle phase (32 % 8 == 0)
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        uint64_t qxy, qzw;
                        ld_shared_2x64(qbase + static_cast<uint32_t>((q * D + b * 32) * 4), qxy, qzw);  // warp-uniform: broadcast
                        a0[q] = f2_fma(r0xy, qxy, a0[q]);
                        a0[q] = f2_fma(r0zw, qzw, a0[q]);
                        a1[q] = f2_fma(r1xy, qxy, a1[q]);
                        a1[q] = f2_fma(r1zw, qzw, a1[q]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
            }
            const uint32_t pp = smem_u32(s_part + ((ti & 1) * kT8Warps + w) * 8 * kT8Rows) + lane * 4;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float lo, hi;
                f2_unpack(a0[q], lo, hi);
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(pp + (q * kT8Rows) * 4), "f"(lo + hi) : "memory");
                f2_unpack(a1[q], lo, hi);
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(pp + (q * kT8Rows + 32) * 4), "f"(lo + hi) : "memory");
            }
            named_bar_sync(1 + team, kT8Warps * 32);
            if (w < nq) {
                const uint32_t pr = smem_u32(s_part + (ti & 1) * kT8Warps * 8 * kT8Rows + w * kT8Rows) + lane * 4;
                float acc0, acc1;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(acc0) : "r"(pr));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(acc1) : "r"(pr + 128));
#pragma unroll
                for (int g = 1; g < kT8Warps; ++g) {
                    float x0, x1;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x0) : "r"(pr + g * 8 * kT8Rows * 4));
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x1) : "r"(pr + g * 8 * kT8Rows * 4 + 128));
                    acc0 += x0;
                    acc1 += x1;
                }
                float s0, s1;
                if (p.mode == SCAN_SEGMENT) {
                    s0 = rn0 < 1e-9f ? 0.0f : acc0 / (qn * rn0);
                    s1 = rn1 < 1e-9f ? 0.0f : acc1 / (qn * rn1);
                } else {
                    s0 = acc0 / fmaxf(qn * rn0, 1e-9f);
                    s1 = acc1 / fmaxf(qn * rn1, 1e-9f);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {  // rows t*64 + [0, 32), then + [32, 64): ascending ids, equal scores keep ascending ids
                    const float s = h ? s1 : s0;
                    uint32_t need = __ballot_sync(0xffffffffu, (h ? valid1 : valid0) && s > thr);
                    while (need) {
                        const int b = __ffs(need) - 1;
                        need &= need - 1;
                        const float sb = __shfl_sync(0xffffffffu, s, b);
                        if (sb > thr) warp_topk_insert(my_sc, my_id, k, cnt, thr, sb, static_cast<uint32_t>(t * kT8Rows + h * 32 + b), lane);
                    }
                }
            }
            ++ti;
        }
        // the team's sorted list of query `w`: list index = blockIdx.x * TEAMS + team
        if (w < nq) {
            const size_t list = static_cast<size_t>(blockIdx.x) * TEAMS + team;
            float* os = p.out_scores + (list * p.Q + p.q0 + w) * k;
            uint32_t* oi = p.out_ids + (list * p.Q + p.q0 + w) * k;
            __syncwarp();
            for (int j = lane; j < k; j += 32) {
                os[j] = j < cnt ? my_sc[j] : -INFINITY;
                oi[j] = j < cnt ? my_id[j] : kNoId32;
            }
        }
    }
}

}  // namespace kj
