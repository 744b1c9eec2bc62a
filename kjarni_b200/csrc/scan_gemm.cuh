// Query-by-index similarity GEMM with a fused per-query candidate filter (tcgen05 / TMEM / TMA), the
// many-query path of the cosine top-k scan.  Replaces, for query batches, the per-row loop of
// Segment::search_vectors (reference: kjarni-rag/src/segment.rs:307-337,355-370) / VectorStore::search
// (kjarni-search/src/vector.rs:131-165) run once per query by IndexReader::search_semantic
// (kjarni-rag/src/index_reader.rs:207-228).
//
// Exactness by construction (top-k ids must be bit-exact on the fp32 index):
//   1. FILTER (this kernel)  scores S~[q,r] = <bf16(q), bf16(r)> / |r| on the tensor cores against a bf16
//      shadow of the index (built once at append time); every CTA keeps, per query, the C = 32 best
//      approximate scores of the rows it streamed.
//   2. merge of the per-CTA lists to the 32 best approximate candidates per query (topk_merge_kernel),
//   3. RESCORE (scan_rescore_kernel): exact fp32 cosine of those 32 rows with the same arithmetic as the
//      exact scan kernel (scan.cuh), final order (score desc, id asc), and a proof check:
//      every row outside the candidate list has approximate cosine <= m32 (the smallest kept one), hence
//      exact cosine <= m32 + eps with eps >= 2^-8 (bf16 rounding of both operands, Cauchy-Schwarz);
//      if exact_kth > m32 + eps the exact top-k is proven.  Otherwise the query is flagged and re-run on
//      the exact scan kernel.
//
// Kernel shape: one CTA per SM, 256 threads.  The 128-query tile (bf16, K-major, 128 B swizzle) is loaded once
// and stays resident in shared memory (96 KB at D = 384); index rows stream through a 3-stage TMA ring in
// 256-row x 64-column boxes; one elected thread issues tcgen05.mma 128 x 256 x 16 into one of two TMEM
// accumulators (2 x 256 columns) so the filter epilogue of tile i overlaps the MMAs of tile i+1.
// Epilogue: 4 warps, thread = query (TMEM lane), tcgen05.ld 32 columns at a time, scale by 1/|r| from smem,
// compare the chunk maximum against the thread's threshold; only then touch the candidate list.
#pragma once
#include <cuda.h>

#include "ptx.cuh"
#include "scan.cuh"

namespace kj {

constexpr int kSgThreads = 256;
constexpr int kSgQ = 128;     // queries per launch tile (TMEM lanes)
constexpr int kSgRows = 256;  // index rows per MMA tile (N)
constexpr int kSgBK = 64;     // bf16 per k-block = one 128-byte swizzle atom
constexpr int kSgStages = 3;
constexpr int kSgC = 32;      // approximate candidates kept per query per CTA
constexpr int kSgMaxD = 384;
constexpr int kSgABlockBytes = kSgQ * kSgBK * 2;                 // 16 KB per k-block of the resident query tile
constexpr int kSgABytes = kSgABlockBytes * (kSgMaxD / kSgBK);    // 96 KB
constexpr int kSgBBytes = kSgRows * kSgBK * 2;                   // 32 KB per stage
constexpr int kSgListBytes = kSgQ * kSgC * 8;                    // 32 KB
constexpr int kSgNormBytes = 2 * kSgRows * 4;                    // inverse row norms of the two tiles in flight
constexpr int kSgSmemBytes = kSgABytes + kSgStages * kSgBBytes + kSgListBytes + kSgNormBytes + 256;  // 231,680 B

struct ScanGemmParams {
    const float* inv_norms;  // [n_rows (+16 slack)]  1/|r|, 0 where |r| < 1e-9
    float* out_scores;       // [gridDim.x, Q, C]  per-CTA candidates sorted (approx score desc, id asc); scores are <q,r>/|r|
    uint32_t* out_ids;       // [gridDim.x, Q, C]  local row index, kNoId32 = empty
    uint32_t n_rows;
    int D, Q, q0;            // q0: first query of this launch's 128-query tile
    const float* thr0;       // [Q] or nullptr: per-query lower bound on the 32nd best approximate score (seed pass), exclusive
    float* seed_max;         // non-null = SEED MODE: [gridDim.x, Q] best approximate score of the CTA's rows; no lists are kept
    int dbg;                 // microbenchmark switches (env KJC_SG_DBG): 1 = no epilogue work, 2 = no MMA issue, 4 = no row loads
};

__global__ void __launch_bounds__(kSgThreads, 1)
scan_gemm_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_rows, ScanGemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_sg[];
    if (smem_u32(smem_sg) & 1023) __trap();
    uint8_t* smem_a = smem_sg;
    uint8_t* smem_b = smem_a + kSgABytes;
    float* l_sc = reinterpret_cast<float*>(smem_b + kSgStages * kSgBBytes);  // [C][128]
    uint32_t* l_id = reinterpret_cast<uint32_t*>(l_sc + kSgQ * kSgC);       // [C][128]
    float* s_inv = reinterpret_cast<float*>(l_id + kSgQ * kSgC);            // [2][256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_inv + 2 * kSgRows);
    uint64_t* full_bar = bars;                       // [stages]
    uint64_t* empty_bar = bars + kSgStages;          // [stages]
    uint64_t* a_bar = bars + 2 * kSgStages;          // [1]
    uint64_t* tmem_full = a_bar + 1;                 // [2]
    uint64_t* tmem_empty = tmem_full + 2;            // [2]
    uint64_t* norm_full = tmem_empty + 2;            // [2]
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(norm_full + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int k_blocks = p.D / kSgBK;
    const uint32_t n_tiles = (p.n_rows + kSgRows - 1) / kSgRows;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_rows);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kSgStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(a_bar, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
            mbar_init(&norm_full[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(a_bar, k_blocks * kSgABlockBytes);
            for (int kb = 0; kb < k_blocks; ++kb)
                tma_load_2d(smem_a + kb * kSgABlockBytes, &tmap_q, a_bar, kb * kSgBK, p.q0, kEvictLast);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                // the 1/|r| buffer of this accumulator slot is free once the epilogue of tile it-2 has released the slot
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                const uint32_t row0 = tile * kSgRows;
                const uint32_t rows_in = min(static_cast<uint32_t>(kSgRows), p.n_rows - row0);
                const uint32_t nb = (rows_in * 4 + 15) & ~15u;
                mbar_arrive_expect_tx(&norm_full[acc], nb);
                bulk_load_1d(s_inv + acc * kSgRows, p.inv_norms + row0, nb, &norm_full[acc]);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (p.dbg & 4) {
                        mbar_arrive(&full_bar[stage]);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], kSgBBytes);
                        tma_load_2d(smem_b + stage * kSgBBytes, &tmap_rows, &full_bar[stage], kb * kSgBK, static_cast<int32_t>(row0), kEvictFirst);
                    }
                    if (++stage == kSgStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(1 /*bf16*/, kSgQ, kSgRows);
            mbar_wait(a_bar, 0);
            tc_fence_after();
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * kSgRows;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + kb * kSgABlockBytes));
                    const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + stage * kSgBBytes));
                    if (!(p.dbg & 2)) {
#pragma unroll
                        for (int k = 0; k < kSgBK / 16; ++k) umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kb == k_blocks - 1) umma_commit(&tmem_full[acc]);
                    if (++stage == kSgStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------- filter epilogue: thread = query
        const int quad = warp & 3;
        const int tq = quad * 32 + lane;  // query inside the tile = TMEM lane
        float* my_sc = l_sc + tq;         // entry j at my_sc[j * 128]
        uint32_t* my_id = l_id + tq;
        const int qi = p.q0 + tq;
        const bool seed_mode = p.seed_max != nullptr;
        // candidates must beat thr: the seed bound until the list is full, then the smallest kept score; padded lanes never insert
        float thr = qi < p.Q ? (p.thr0 != nullptr ? p.thr0[qi] : -INFINITY) : INFINITY;
        float best = -INFINITY;           // seed mode: running maximum
        int cnt = 0, minpos = 0;
        int it = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            const uint32_t row0 = tile * kSgRows;
            const int rows_in = static_cast<int>(min(static_cast<uint32_t>(kSgRows), p.n_rows - row0));
            mbar_wait(&norm_full[acc], ph);
            mbar_wait(&tmem_full[acc], ph);
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kSgRows;
            const float* inv = s_inv + acc * kSgRows;
#pragma unroll 1
            for (int c = 0; c < kSgRows / 32; ++c) {
                if (c * 32 >= rows_in || (p.dbg & 1)) break;  // warp-uniform
                uint32_t v[32];
                tmem_ld_32x32(taddr0 + c * 32, v);
                float w[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 f = *reinterpret_cast<const float4*>(inv + c * 32 + 4 * j);  // broadcast read
                    w[4 * j] = f.x; w[4 * j + 1] = f.y; w[4 * j + 2] = f.z; w[4 * j + 3] = f.w;
                }
                tmem_ld_wait();
                float s[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) s[j] = __uint_as_float(v[j]) * w[j];
                if (c * 32 + 32 > rows_in) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c * 32 + j >= rows_in) s[j] = -INFINITY;
                }
                float m = s[0];
#pragma unroll
                for (int j = 1; j < 32; ++j) m = fmaxf(m, s[j]);
                best = fmaxf(best, m);
                if (!seed_mode && m > thr) {
                    const uint32_t rid0 = row0 + c * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (s[j] > thr) {
                            int pos = cnt;
                            if (cnt == kSgC) pos = minpos; else ++cnt;
                            my_sc[pos * kSgQ] = s[j];
                            my_id[pos * kSgQ] = rid0 + j;
                            if (cnt == kSgC) {  // list full: new threshold = smallest kept score
                                float mn = INFINITY;
                                int mp = 0;
#pragma unroll 1
                                for (int e = 0; e < kSgC; ++e) {
                                    const float x = my_sc[e * kSgQ];
                                    if (x < mn) { mn = x; mp = e; }
                                }
                                thr = mn;
                                minpos = mp;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        // sort my list (approx score desc, id asc) and write it out
        for (int i = 1; i < cnt; ++i) {
            const float si = my_sc[i * kSgQ];
            const uint32_t ii = my_id[i * kSgQ];
            int j = i - 1;
            while (j >= 0) {
                const float sj = my_sc[j * kSgQ];
                const uint32_t ij = my_id[j * kSgQ];
                if (sj > si || (sj == si && ij < ii)) break;
                my_sc[(j + 1) * kSgQ] = sj;
                my_id[(j + 1) * kSgQ] = ij;
                --j;
            }
            my_sc[(j + 1) * kSgQ] = si;
            my_id[(j + 1) * kSgQ] = ii;
        }
        if (seed_mode) {
            if (qi < p.Q) p.seed_max[static_cast<size_t>(blockIdx.x) * p.Q + qi] = best;
        } else if (qi < p.Q) {
            float* os = p.out_scores + (static_cast<size_t>(blockIdx.x) * p.Q + qi) * kSgC;
            uint32_t* oi = p.out_ids + (static_cast<size_t>(blockIdx.x) * p.Q + qi) * kSgC;
            for (int j = 0; j < kSgC; ++j) {
                os[j] = j < cnt ? my_sc[j * kSgQ] : -INFINITY;
                oi[j] = j < cnt ? my_id[j * kSgQ] : kNoId32;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// |row|, 1/|row| and the bf16 shadow of rows [0, n): one warp per row (append-time preparation, also used for queries).
__global__ void __launch_bounds__(256)
row_prep_kernel(const float* __restrict__ rows, float* __restrict__ norms, float* __restrict__ inv_norms, __nv_bfloat16* __restrict__ rows16,
                size_t n, int D) {
    const size_t row = static_cast<size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const int lane = threadIdx.x & 31;
    const float* r = rows + row * D;
    float s = 0.f;
    for (int c = lane * 4; c < D; c += 128) {
        const float4 v = ld_stream_f4(r + c);
        s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);  // same order as row_norm_kernel: identical cached norms
        if (rows16) {
            uint2 o;
            o.x = pack_bf16(v.x, v.y);
            o.y = pack_bf16(v.z, v.w);
            *reinterpret_cast<uint2*>(rows16 + row * D + c) = o;
        }
    }
    s = warp_sum(s);
    if (lane == 0) {
        const float nm = sqrtf(s);
        if (norms) norms[row] = nm;
        if (inv_norms) inv_norms[row] = nm < 1e-9f ? 0.0f : 1.0f / nm;
    }
}

// Seed bound for the filter: the 32nd largest of the per-CTA maxima of a sample of the shard (one 256-row tile per CTA).
// The maxima belong to distinct rows, so at least 32 rows of the shard score >= that value: it is a valid lower bound on
// the 32nd best approximate score, and lets every CTA skip list maintenance for all but ~0.1 % of its rows.
// thr0[q] is exclusive (rows must score > thr0), hence the small margin below the selected value.
__global__ void __launch_bounds__(256) scan_seed_select_kernel(const float* __restrict__ seed_max, int L, int Q, float* __restrict__ thr0) {
    const int qi = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (qi >= Q) return;
    const int lane = threadIdx.x & 31;
    if (L < kSgC) {
        if (lane == 0) thr0[qi] = -INFINITY;
        return;
    }
    constexpr int kPer = 8;  // up to 256 CTAs
    float v[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) v[i] = lane + 32 * i < L ? seed_max[static_cast<size_t>(lane + 32 * i) * Q + qi] : -INFINITY;
    float sel = -INFINITY;
    for (int r = 0; r < kSgC; ++r) {
        float m = v[0];
        int mi = 0;
#pragma unroll
        for (int i = 1; i < kPer; ++i)
            if (v[i] > m) { m = v[i]; mi = i; }
        float wm = m;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
        const uint32_t who = __ballot_sync(0xffffffffu, m == wm);
        if (lane == __ffs(who) - 1) {
#pragma unroll
            for (int i = 0; i < kPer; ++i)
                if (i == mi) v[i] = -INFINITY;
        }
        sel = wm;
    }
    if (lane == 0) thr0[qi] = sel == -INFINITY ? -INFINITY : sel - fabsf(sel) * 1e-5f - 1e-12f;
}

struct RescoreParams {
    const float* rows;         // [n_rows, D] fp32 index
    const float* norms;        // [n_rows]
    const float* queries;      // [Q, D] fp32
    const float* qnorms;       // [Q]
    const uint64_t* cand_ids;  // [Q, C] merged approximate candidates (global ids; kNoId64 = empty), sorted by approx score desc
    const float* cand_scores;  // [Q, C] approximate <q,r>/|r|
    uint64_t id_base;
    uint64_t* out_ids;         // [Q, k]
    float* out_scores;         // [Q, k]
    int32_t* out_counts;       // [Q] or nullptr
    int32_t* flags;            // [Q] 1 = not proven exact, needs the exact scan
    int32_t* n_flagged;        // running count of flagged queries
    float eps;                 // bound on |approx cosine - exact cosine|
    int D, Q, k, mode;
};

// One warp per query: exact fp32 cosine of the 32 candidates (arithmetic identical to scan_topk_kernel: per-lane packed
// FMA over float4 chunks, lo + hi, xor-butterfly 16..1, one divide), final order (score desc, id asc), proof check.
__global__ void __launch_bounds__(256) scan_rescore_kernel(RescoreParams p) {
    const int qi = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (qi >= p.Q) return;
    const int lane = threadIdx.x & 31;
    const float qn = p.qnorms[qi];
    const uint64_t my_gid = p.cand_ids[static_cast<size_t>(qi) * kSgC + lane];
    const float my_approx = p.cand_scores[static_cast<size_t>(qi) * kSgC + lane];
    const int ncand = __popc(__ballot_sync(0xffffffffu, my_gid != kNoId64));
    float my_s = -INFINITY;
    const float* q = p.queries + static_cast<size_t>(qi) * p.D;
    for (int c = 0; c < ncand; ++c) {
        const uint64_t gid = __shfl_sync(0xffffffffu, my_gid, c);  // candidates are packed at the front (sorted lists)
        const size_t r = static_cast<size_t>(gid - p.id_base);
        const float* rp = p.rows + r * p.D;
        uint64_t a2 = 0ull;
        for (int col = lane * 4; col < p.D; col += 128) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(rp + col);
            const float4 f = *reinterpret_cast<const float4*>(q + col);
            a2 = f2_fma(v.x, f2_pack(f.x, f.y), a2);
            a2 = f2_fma(v.y, f2_pack(f.z, f.w), a2);
        }
        float lo, hi;
        f2_unpack(a2, lo, hi);
        float acc = lo + hi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const float rn = p.norms[r];
        float s;
        if (p.mode == SCAN_SEGMENT) s = rn < 1e-9f ? 0.0f : acc / (qn * rn);
        else s = acc / fmaxf(qn * rn, 1e-9f);
        if (lane == c) my_s = s;
    }
    // rank among the candidates by (score desc, id asc)
    int rank = 0;
#pragma unroll 1
    for (int j = 0; j < kSgC; ++j) {
        const float sj = __shfl_sync(0xffffffffu, my_s, j);
        const uint64_t ij = __shfl_sync(0xffffffffu, my_gid, j);
        if (j < ncand && (sj > my_s || (sj == my_s && ij < my_gid))) ++rank;
    }
    const bool valid = lane < ncand;
    const bool empty_query = p.mode == SCAN_SEGMENT && qn < 1e-9f;
    const int nres = empty_query ? 0 : min(p.k, ncand);
    uint64_t* oi = p.out_ids + static_cast<size_t>(qi) * p.k;
    float* os = p.out_scores + static_cast<size_t>(qi) * p.k;
    if (valid && rank < nres) { oi[rank] = my_gid; os[rank] = my_s; }
    for (int j = nres + lane; j < p.k; j += 32) { oi[j] = kNoId64; os[j] = -INFINITY; }
    if (lane == 0 && p.out_counts) p.out_counts[qi] = nres;
    // proof: rows outside the list have approximate cosine <= m32 (the smallest kept one) => exact <= m32 + eps
    const uint32_t kth_mask = __ballot_sync(0xffffffffu, valid && rank == p.k - 1);
    const int kth_lane = kth_mask ? __ffs(kth_mask) - 1 : 0;
    const float kth = __shfl_sync(0xffffffffu, my_s, kth_lane);
    const float m32 = __shfl_sync(0xffffffffu, my_approx, kSgC - 1);  // lists are sorted descending: last = smallest
    if (lane == 0 && !empty_query) {
        bool proven = ncand < kSgC;
        if (!proven && kth_mask != 0 && qn >= 1e-9f) proven = kth > m32 / qn + p.eps;
        p.flags[qi] = proven ? 0 : 1;
        if (!proven) atomicAdd(p.n_flagged, 1);
    } else if (lane == 0) {
        p.flags[qi] = 0;
    }
}

}  // namespace kj
