// Query-by-index similarity GEMM with a fused per-query candidate filter (tcgen05 / TMEM / TMA), the
// many-query path of the cosine top-k scan.  Replaces, for query batches, the per-row loop of
// Segment::search_vectors (reference: kjarni-rag/src/segment.rs:307-337,355-370) / VectorStore::search
// (kjarni-search/src/vector.rs:131-165) run once per query by IndexReader::search_semantic
// (kjarni-rag/src/index_reader.rs:207-228).
//
// Exactness by construction (top-k ids must be bit-exact on the fp32 index):
//   0. SEED    the same kernel in seed mode scores a sample of the shard (evenly spaced 256-row tiles) and records
//              group maxima; the 32nd largest maximum is a lower bound thr0[q] on the 32nd best approximate score
//              of the whole shard (the maxima belong to 32 distinct rows).
//   1. FILTER  scores S~[q,r] = <bf16(q), bf16(r/|r|)> on the tensor cores against a normalised bf16 shadow of the
//              index (built once at append time); rows with S~ > thr0[q] (about 0.01 % of them) are appended to the
//              query's candidate buffer in global memory.
//   2. SELECT  scan_cand_select_kernel keeps the 32 best approximate candidates per query.
//   3. RESCORE scan_rescore_kernel: exact fp32 cosine of those 32 rows with the same arithmetic as the exact scan
//              kernel (scan.cuh), final order (score desc, id asc), and a proof check: every row outside the list
//              has approximate cosine <= m32 (the smallest kept one), hence exact cosine <= m32 + eps with
//              eps >= 2^-8 (bf16 rounding of both operands, Cauchy-Schwarz); if exact_kth > m32 + eps the exact
//              top-k is proven.  Otherwise (or if the candidate buffer overflowed) the query is flagged and re-run
//              on the exact scan kernel.
//
// Kernel shape: one CTA per SM, 256 threads, persistent over (row superblock, query tile, row tile):
//   * the shard is cut into superblocks of gridDim.x * R row tiles (R * 196 KB per CTA, ~60-90 MB in total) that stay
//     L2-resident while ALL query tiles are scored against them, so HBM is read once per search, not once per
//     128 queries;
//   * the current 128-query tile (bf16, K-major, 128 B swizzle, 6 k-blocks of 16 KB) lives in shared memory and is
//     replaced k-block by k-block as soon as the last MMA that reads a k-block has retired (warp 3), so switching
//     query tiles costs no pipeline drain;
//   * row tiles stream through a 4-stage TMA ring of 256-row x 64-column boxes (warp 0);
//   * one elected thread (warp 1) issues tcgen05.mma 128 x 256 x 16 into one of two TMEM accumulators (2 x 256
//     columns), so the filter epilogue of tile i overlaps the MMAs of tile i+1;
//   * epilogue (warps 4-11, two per TMEM lane quadrant, 128 columns each): thread = query (TMEM lane), tcgen05.ld
//     64 columns at a time, tree maximum against the thread's threshold; passers are appended with one atomicAdd each.
#pragma once
#include <cuda.h>

#include "ptx.cuh"
#include "scan.cuh"

namespace kj {

constexpr int kSgThreads = 384;
constexpr int kSgEpiWarps = 8;
constexpr int kSgQ = 128;     // queries per tile (TMEM lanes)
constexpr int kSgRows = 256;  // index rows per MMA tile (N)
constexpr int kSgBK = 64;     // bf16 per k-block = one 128-byte swizzle atom
constexpr int kSgStages = 4;
constexpr int kSgC = 32;      // approximate candidates kept per query for the exact rescoring when k <= 16 ...
constexpr int kSgCWide = 128; // ... and when 16 < k <= 64 (a reranking Searcher fetches top_k * 5 = 50, kjarni/src/searcher/model.rs:117-121)
constexpr int kSgCap = 2048;  // candidate buffer entries per query
constexpr int kSgMaxD = 384;                                     // widest index whose 128-query tile stays resident in shared memory
constexpr int kSgMaxDStream = 1024;                              // wider indexes (768-dim embedders) stream the query tile k-block by k-block
constexpr int kSgMaxKB = kSgMaxD / kSgBK;                        // 6
constexpr int kSgABlockBytes = kSgQ * kSgBK * 2;                 // 16 KB per k-block of the resident query tile
constexpr int kSgABytes = kSgABlockBytes * kSgMaxKB;             // 96 KB
constexpr int kSgBBytes = kSgRows * kSgBK * 2;                   // 32 KB per stage
constexpr int kSgPairStages = 8;                                 // kPair: the same 128 KB as a ring of eight half tiles
constexpr int kSgSmemBytes = kSgABytes + kSgStages * kSgBBytes + 512;  // 229,888 B
constexpr int kSgSeedGroupsMax = 320;                            // seed maxima per query (select kernel: 10 per lane)

struct ScanGemmParams {
    const float* thr0;       // [Q] or nullptr: exclusive lower bound a row's approximate score must beat
    float* cand_scores;      // [Q, kSgCap]  approximate <q, r/|r|> of the passers (unordered)
    uint32_t* cand_ids;      // [Q, kSgCap]  local row index
    uint32_t* cand_cnt;      // [Q]          passers seen (may exceed kSgCap: overflow)
    float* seed_max;         // non-null = SEED MODE: [groups, Q] group maxima, no candidates are written
    uint32_t n_rows;
    uint32_t n_tiles;        // tiles this launch visits (seed mode: sample size); tile i covers shard tile tile_of(i)
    uint32_t n_tiles_total;  // ceil(n_rows / 256)
    int seed_chunks;         // seed mode: 1 = one group per 32-row chunk (group = tile*8 + chunk), 0 = one group per CTA
    int D, Q, R;             // R: row tiles per CTA per superblock
    int dbg;                 // microbenchmark switches (env KJC_SG_DBG): 1 = no epilogue work, 2 = no MMA issue, 4 = no row loads
    int stream_a;            // D > 384: the query tile does not fit beside the row ring, so every stage carries the query k-block
                             // (16 KB) next to the row k-block (32 KB); the query tile is re-read from L2 once per row tile
};
constexpr int kSgStreamStageBytes = kSgABlockBytes + kSgBBytes;  // 48 KB
static_assert(kSgStages * kSgStreamStageBytes + 512 <= kSgSmemBytes, "streaming stages fit in the same allocation");

__device__ __forceinline__ uint32_t sg_tile_of(const ScanGemmParams& p, uint32_t i) {
    return p.n_tiles == p.n_tiles_total ? i : static_cast<uint32_t>(static_cast<uint64_t>(i) * p.n_tiles_total / p.n_tiles);
}

// kPair: two CTAs of a cluster work on the SAME row tiles for two DIFFERENT query tiles (tcgen05.mma.cta_group::2, M = 256: 128 queries
// per CTA).  Each CTA loads only half of every row tile (128 of its 256 rows) and the tensor core reads both halves.  With the query
// tile resident, a 128 x 256 x 64 k-block needs 32 KB of rows per 512 clk of MMAs: 64 B/clk per SM, 18 TB/s for 148 SMs streaming out
// of the L2-resident superblock against the ~10 TB/s L2 delivers -- the one-CTA filter pass measures 0.65 of the tensor roof, and the
// chained encoder kernels had the same bound.  Half the row bytes per SM, and the same 128 KB as a ring of eight
// half tiles instead of four whole ones.  Queries stay owned by one CTA (its TMEM lanes): the filter epilogue is unchanged, it only
// releases accumulators on the leader's barrier (relaxed cluster arrive).  `tmap_rows` has 128-row boxes in this mode; query tile of
// iteration j = 2j + rank (a tile beyond the batch is all padding); filter pass with a resident query tile only (dim <= 384).
template <bool kPair>
__global__ void __launch_bounds__(kSgThreads, 1)
scan_gemm_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_rows, ScanGemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_sg[];
    if (smem_u32(smem_sg) & 1023) __trap();
    uint8_t* smem_a = smem_sg;
    uint8_t* smem_b = smem_a + kSgABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + kSgStages * kSgBBytes);
    uint64_t* full_bar = bars;                       // [stages]
    constexpr int kNS = kPair ? kSgPairStages : kSgStages;  // row stages (kPair: half tiles, resident query tile only)
    constexpr int kBStride = kPair ? kSgBBytes / 2 : kSgBBytes;
    uint64_t* empty_bar = full_bar + kNS;            // [stages]
    uint64_t* a_full = empty_bar + kNS;              // [6]
    uint64_t* a_empty = a_full + kSgMaxKB;           // [6]
    uint64_t* tmem_full = a_empty + kSgMaxKB;        // [2]
    uint64_t* tmem_empty = tmem_full + 2;            // [2]
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int k_blocks = p.D / kSgBK;
    const int n_qt_all = (p.Q + kSgQ - 1) / kSgQ;
    const bool stream_a = p.stream_a != 0;
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const uint32_t widx = kPair ? cluster_id_x() : blockIdx.x;          // worker (CTA or CTA pair) index / count: row tiles are dealt to workers
    const uint32_t nworkers = kPair ? cluster_nctaid_x() : gridDim.x;
    const int n_qt = kPair ? (n_qt_all + 1) / 2 : n_qt_all;              // query-tile iterations; this CTA's tile of iteration j:
    auto my_qt = [&](int j) { return kPair ? 2 * j + static_cast<int>(rank) : j; };
    constexpr uint32_t kBLoadBytes = kPair ? kSgBBytes / 2 : kSgBBytes;  // row bytes this CTA loads per stage
    constexpr int kRowsLoaded = kPair ? kSgRows / 2 : kSgRows;
    // resident mode: [query tile 96 KB][4 row stages of 32 KB]; streaming mode: 4 stages of [query k-block 16 KB | row k-block 32 KB]
    auto stage_a = [&](int stage) -> uint8_t* { return smem_sg + stage * kSgStreamStageBytes; };
    auto stage_b = [&](int stage) -> uint8_t* { return stream_a ? smem_sg + stage * kSgStreamStageBytes + kSgABlockBytes : smem_b + stage * kBStride; };
    const uint32_t tiles_per_sb = nworkers * static_cast<uint32_t>(p.R);
    const uint32_t n_sb = (p.n_tiles + tiles_per_sb - 1) / tiles_per_sb;
    // tiles of this worker in superblock sb: i = sb*tiles_per_sb + r*nworkers + widx, r < R, while i < n_tiles

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_rows);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kNS; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < kSgMaxKB; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], kPair ? 2 * kSgEpiWarps : kSgEpiWarps);  // one arrive per epilogue warp (kPair: of both CTAs, on the leader's)
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        if constexpr (kPair) tmem_alloc_2sm<512>(tmem_base_smem);
        else tmem_alloc<512>(tmem_base_smem);
    }
    tc_fence_before();
    if constexpr (kPair) cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / TMA completion
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer: row tiles + their 1/|r|
        if (lane == 0) {
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (uint32_t sb = 0; sb < n_sb; ++sb) {
                for (int qt = 0; qt < n_qt; ++qt) {
                    for (int r = 0; r < p.R; ++r) {
                        const uint32_t i = sb * tiles_per_sb + r * nworkers + widx;
                        if (i >= p.n_tiles) break;
                        const uint32_t row0 = sg_tile_of(p, i) * kSgRows + rank * kRowsLoaded;  // kPair: this CTA's half of the row tile
                        for (int kb = 0; kb < k_blocks; ++kb) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            if (p.dbg & 4) {
                                if (leader) mbar_arrive(&full_bar[stage]);
                            } else if constexpr (kPair) {
                                const uint32_t lbar = mapa_shared(smem_u32(&full_bar[stage]), 0);  // all bytes are counted on the leader's barrier
                                if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (kBLoadBytes + (stream_a ? kSgABlockBytes : 0)));
                                if (stream_a) tma_load_2d_2sm(stage_a(stage), &tmap_q, lbar, kb * kSgBK, my_qt(qt) * kSgQ, kEvictLast);
                                tma_load_2d_2sm(stage_b(stage), &tmap_rows, lbar, kb * kSgBK, static_cast<int32_t>(row0), kEvictNormal);
                            } else {
                                mbar_arrive_expect_tx(&full_bar[stage], stream_a ? kSgStreamStageBytes : kSgBBytes);
                                if (stream_a) tma_load_2d(stage_a(stage), &tmap_q, &full_bar[stage], kb * kSgBK, qt * kSgQ, kEvictLast);
                                tma_load_2d(stage_b(stage), &tmap_rows, &full_bar[stage], kb * kSgBK, static_cast<int32_t>(row0), kEvictNormal);
                            }
                            if (++stage == kNS) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                        ++it;
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------ TMA producer: query tiles, k-block by k-block (resident mode only)
        if (lane == 0 && !stream_a) {
            uint32_t ai = 0;
            for (uint32_t sb = 0; sb < n_sb; ++sb) {
                if (sb * tiles_per_sb + widx >= p.n_tiles) break;  // no tile of this worker in the last superblock
                for (int qt = 0; qt < n_qt; ++qt, ++ai) {
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        mbar_wait(&a_empty[kb], (ai & 1) ^ 1);
                        if constexpr (kPair) {
                            if (leader) mbar_arrive_expect_tx(&a_full[kb], 2 * kSgABlockBytes);
                            tma_load_2d_2sm(smem_a + kb * kSgABlockBytes, &tmap_q, mapa_shared(smem_u32(&a_full[kb]), 0), kb * kSgBK, my_qt(qt) * kSgQ,
                                            kEvictLast);
                        } else {
                            mbar_arrive_expect_tx(&a_full[kb], kSgABlockBytes);
                            tma_load_2d(smem_a + kb * kSgABlockBytes, &tmap_q, &a_full[kb], kb * kSgBK, qt * kSgQ, kEvictLast);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer
        if ((KJ_MMA_UNIFORM != 0 || lane == 0) && leader) {  // kPair: the leader issues for both CTAs
            constexpr uint32_t idesc = umma_idesc(1 /*bf16*/, kPair ? 2 * kSgQ : kSgQ, kSgRows);
            auto commit = [&](uint64_t* bar) {
                if constexpr (kPair) umma_commit_2sm(bar, 3);  // the barrier at this offset in both CTAs
                else umma_commit(bar);
            };
            int stage = 0, it = 0;
            uint32_t phase = 0, ai = 0;
            for (uint32_t sb = 0; sb < n_sb; ++sb) {
                if (sb * tiles_per_sb + widx >= p.n_tiles) break;
                for (int qt = 0; qt < n_qt; ++qt, ++ai) {
                    for (int r = 0; r < p.R; ++r) {
                        const uint32_t i = sb * tiles_per_sb + r * nworkers + widx;
                        if (i >= p.n_tiles) break;
                        const bool last_r = (r + 1 == p.R) || (i + nworkers >= p.n_tiles);
                        const int acc = it & 1;
                        if constexpr (kPair) mbar_wait_cluster(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);  // the peer's arrives are remote
                        else mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                        tc_fence_after();
                        const uint32_t tmem_d = tmem_base + acc * kSgRows;
                        for (int kb = 0; kb < k_blocks; ++kb) {
                            if (r == 0 && !stream_a) mbar_wait(&a_full[kb], ai & 1);
                            mbar_wait(&full_bar[stage], phase);
                            tc_fence_after();
                            const uint64_t da = umma_desc_k_sw128(smem_u32(stream_a ? stage_a(stage) : smem_a + kb * kSgABlockBytes));
                            const uint64_t db = umma_desc_k_sw128(smem_u32(stage_b(stage)));
                            if (mma_issuer_lane()) {
                                if (!(p.dbg & 2)) {
#pragma unroll
                                    for (int k = 0; k < kSgBK / 16; ++k) {
                                        if constexpr (kPair) umma_f16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                                        else umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                                    }
                                }
                                commit(&empty_bar[stage]);
                                if (last_r && !stream_a) commit(&a_empty[kb]);  // this k-block of the query tile may be replaced
                                if (kb == k_blocks - 1) commit(&tmem_full[acc]);
                            }
                            mma_issuer_sync();
                            if (++stage == kNS) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                        ++it;
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------- filter epilogue: thread = query
        const int quad = warp & 3;
        const int half = (warp - 4) >> 2;  // column half of the 256-row tile handled by this warp
        const int tq = quad * 32 + lane;   // query inside the tile = TMEM lane
        const bool seed_mode = p.seed_max != nullptr;
        constexpr int kHalfCols = kSgRows / 2;  // 128
        int it = 0;
        const uint32_t leader_empty[2] = {kPair ? mapa_shared(smem_u32(&tmem_empty[0]), 0) : 0u, kPair ? mapa_shared(smem_u32(&tmem_empty[1]), 0) : 0u};
        for (uint32_t sb = 0; sb < n_sb; ++sb) {
            if (sb * tiles_per_sb + widx >= p.n_tiles) break;
            for (int qt = 0; qt < n_qt; ++qt) {
                const int qi = my_qt(qt) * kSgQ + tq;
                const bool live = qi < p.Q;
                // rows must beat thr; padded lanes never pass
                const float thr = live ? (p.thr0 != nullptr ? p.thr0[qi] : -INFINITY) : INFINITY;
                float best = -INFINITY;  // seed mode, one group per CTA: running maximum over this CTA's sample tiles
                for (int r = 0; r < p.R; ++r) {
                    const uint32_t i = sb * tiles_per_sb + r * nworkers + widx;
                    if (i >= p.n_tiles) break;
                    const int acc = it & 1;
                    const uint32_t ph = (it >> 1) & 1;
                    const uint32_t row0 = sg_tile_of(p, i) * kSgRows;
                    const int rows_in = static_cast<int>(min(static_cast<uint32_t>(kSgRows), p.n_rows - row0));
                    mbar_wait(&tmem_full[acc], ph);
                    tc_fence_after();
                    const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kSgRows + half * kHalfCols;
#pragma unroll 1
                    for (int c = 0; c < kHalfCols / 64; ++c) {
                        const int col0 = half * kHalfCols + c * 64;  // first tile column of this 64-column chunk
                        float m0 = -INFINITY, m1 = -INFINITY;       // maxima of its two 32-row groups
                        if (col0 < rows_in && !(p.dbg & 1)) {       // warp-uniform
                            uint32_t v0[32], v1[32];
                            tmem_ld_32x32(taddr0 + c * 64, v0);
                            tmem_ld_32x32(taddr0 + c * 64 + 32, v1);
                            tmem_ld_wait();
                            float s[64];
#pragma unroll
                            for (int j = 0; j < 32; ++j) { s[j] = __uint_as_float(v0[j]); s[32 + j] = __uint_as_float(v1[j]); }
                            if (col0 + 64 > rows_in) {
#pragma unroll
                                for (int j = 0; j < 64; ++j)
                                    if (col0 + j >= rows_in) s[j] = -INFINITY;
                            }
                            float t[32];
#pragma unroll
                            for (int j = 0; j < 16; ++j) { t[j] = fmaxf(s[2 * j], s[2 * j + 1]); t[16 + j] = fmaxf(s[32 + 2 * j], s[33 + 2 * j]); }
#pragma unroll
                            for (int w = 8; w > 0; w >>= 1) {
#pragma unroll
                                for (int j = 0; j < w; ++j) { t[j] = fmaxf(t[j], t[j + w]); t[16 + j] = fmaxf(t[16 + j], t[16 + j + w]); }
                            }
                            m0 = t[0];
                            m1 = t[16];
                            if (!seed_mode && fmaxf(m0, m1) > thr) {
                                const uint32_t rid0 = row0 + col0;
#pragma unroll
                                for (int j = 0; j < 64; ++j) {
                                    if (s[j] > thr) {
                                        const uint32_t pos = atomicAdd(p.cand_cnt + qi, 1u);
                                        if (pos < static_cast<uint32_t>(kSgCap)) {
                                            p.cand_scores[static_cast<size_t>(qi) * kSgCap + pos] = s[j];
                                            p.cand_ids[static_cast<size_t>(qi) * kSgCap + pos] = rid0 + j;
                                        }
                                    }
                                }
                            }
                        }
                        if (seed_mode) {
                            if (p.seed_chunks) {
                                if (live) {
                                    p.seed_max[static_cast<size_t>(i * 8 + (col0 >> 5)) * p.Q + qi] = m0;
                                    p.seed_max[static_cast<size_t>(i * 8 + (col0 >> 5) + 1) * p.Q + qi] = m1;
                                }
                            } else {
                                best = fmaxf(best, fmaxf(m0, m1));
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (kPair) mbar_arrive_cluster_relaxed(leader_empty[acc]);  // "finished reading": no release needed (ptx.cuh)
                        else mbar_arrive(&tmem_empty[acc]);
                    }
                    ++it;
                }
                // one group per worker and column half (2 * nworkers groups)
                if (seed_mode && !p.seed_chunks && live) {
                    float* g = p.seed_max + static_cast<size_t>(widx * 2 + half) * p.Q + qi;  // only this thread touches it
                    *g = sb == 0 ? best : fmaxf(*g, best);
                }
            }
        }
    }

    tc_fence_before();
    if constexpr (kPair) cluster_sync_all();  // peer shared memory / barriers stay valid until both CTAs are done
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (kPair) tmem_dealloc_2sm<512>(tmem_base);
        else tmem_dealloc<512>(tmem_base);
    }
}

// The 32 best approximate candidates of every query, sorted (score desc, id asc): one CTA per query, 32 rounds of a
// block-wide arg-max over the query's candidate buffer staged in shared memory.
struct CandSelectParams {
    const float* cand_scores;  // [Q, kSgCap]
    const uint32_t* cand_ids;  // [Q, kSgCap]
    const uint32_t* cand_cnt;  // [Q]
    uint64_t id_base;
    float* out_scores;         // [Q, kSgC]  -inf = empty
    uint64_t* out_ids;         // [Q, kSgC]  kNoId64 = empty
    int32_t* overflow;         // [Q] 1 = the buffer overflowed (candidates were dropped)
    int C;                     // candidates kept per query: kSgC or kSgCWide
    const int32_t* qmap;       // nullptr, or [gridDim.x] query indices: block b selects for query qmap[b] and writes list b (escalation)
};
__global__ void __launch_bounds__(256) scan_cand_select_kernel(CandSelectParams p) {
    __shared__ float sc[kSgCap];
    __shared__ uint32_t id[kSgCap];
    __shared__ float ws[8];
    __shared__ uint32_t wi[8];
    __shared__ int wp[8];
    const int qi = p.qmap ? p.qmap[blockIdx.x] : static_cast<int>(blockIdx.x), tid = threadIdx.x;
    const int slot = blockIdx.x;  // output list
    const uint32_t cnt = p.cand_cnt[qi];
    const int n = static_cast<int>(min(cnt, static_cast<uint32_t>(kSgCap)));
    for (int j = tid; j < n; j += 256) {
        sc[j] = p.cand_scores[static_cast<size_t>(qi) * kSgCap + j];
        id[j] = p.cand_ids[static_cast<size_t>(qi) * kSgCap + j];
    }
    if (tid == 0) p.overflow[qi] = cnt > static_cast<uint32_t>(kSgCap) ? 1 : 0;
    __syncthreads();
    for (int round = 0; round < p.C; ++round) {
        float bs = -INFINITY;
        uint32_t bi = kNoId32;
        int bp = -1;
        for (int j = tid; j < n; j += 256) {
            const float s = sc[j];
            const uint32_t ii = id[j];
            if (ii != kNoId32 && (bp < 0 || s > bs || (s == bs && ii < bi))) { bs = s; bi = ii; bp = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int op = __shfl_xor_sync(0xffffffffu, bp, o);
            if (op >= 0 && (bp < 0 || os > bs || (os == bs && oi < bi))) { bs = os; bi = oi; bp = op; }
        }
        if ((tid & 31) == 0) { ws[tid >> 5] = bs; wi[tid >> 5] = bi; wp[tid >> 5] = bp; }
        __syncthreads();
        if (tid == 0) {
            float fs = -INFINITY;
            uint32_t fi = kNoId32;
            int fp = -1;
            for (int w = 0; w < 8; ++w)
                if (wp[w] >= 0 && (fp < 0 || ws[w] > fs || (ws[w] == fs && wi[w] < fi))) { fs = ws[w]; fi = wi[w]; fp = wp[w]; }
            p.out_scores[static_cast<size_t>(slot) * p.C + round] = fp >= 0 ? fs : -INFINITY;
            p.out_ids[static_cast<size_t>(slot) * p.C + round] = fp >= 0 ? p.id_base + fi : kNoId64;
            if (fp >= 0) id[fp] = kNoId32;  // taken
        }
        __syncthreads();
    }
}

// |row| and the bf16 shadow of rows [0, n): one warp per row (append-time preparation, also used for queries).
// normalise = 1: the shadow holds bf16(r / |r|) (zero rows stay zero), so the filter GEMM yields cosine * |q| directly.
template <int NCH>
__global__ void __launch_bounds__(256)
row_prep_kernel(const float* __restrict__ rows, float* __restrict__ norms, __nv_bfloat16* __restrict__ rows16, size_t n, int D, int normalise) {
    const size_t row = static_cast<size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const int lane = threadIdx.x & 31;
    const float* r = rows + row * D;
    float4 v[NCH];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = (lane + 32 * i) * 4;
        v[i] = c < D ? ld_stream_f4(r + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);  // same order as row_norm_kernel: identical norms
    }
    s = warp_sum(s);
    const float nm = sqrtf(s);
    if (lane == 0 && norms) norms[row] = nm;
    const float sc = normalise ? (nm < 1e-9f ? 0.0f : 1.0f / nm) : 1.0f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = (lane + 32 * i) * 4;
        if (c < D) {
            uint2 o;
            o.x = pack_bf16(v[i].x * sc, v[i].y * sc);
            o.y = pack_bf16(v[i].z * sc, v[i].w * sc);
            *reinterpret_cast<uint2*>(rows16 + row * D + c) = o;
        }
    }
}

// Seed bound for the filter: the 32nd largest of L <= 320 group maxima of the sample (groups = the sample tiles of one
// CTA, or 32-row chunks for small shards).  The maxima belong to distinct rows, so at least 32 rows of the shard score
// >= that value: a valid lower bound on the 32nd best approximate score, which lets the filter drop ~99.99 % of the rows
// with one compare.  thr0[q] is exclusive (rows must score > thr0), hence the small margin below the selected value.
__global__ void __launch_bounds__(256) scan_seed_select_kernel(const float* __restrict__ seed_max, int L, int Q, float* __restrict__ thr0, int C) {
    const int qi = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (qi >= Q) return;
    const int lane = threadIdx.x & 31;
    if (L < C) {
        if (lane == 0) thr0[qi] = -INFINITY;
        return;
    }
    constexpr int kPer = kSgSeedGroupsMax / 32;
    float v[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) v[i] = lane + 32 * i < L ? seed_max[static_cast<size_t>(lane + 32 * i) * Q + qi] : -INFINITY;
    float sel = -INFINITY;
    for (int r = 0; r < C; ++r) {
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
    const float* cand_scores;  // [Q, C] approximate <q, r/|r|>
    uint64_t id_base;
    uint64_t* out_ids;         // [Q, k]
    float* out_scores;         // [Q, k]
    int32_t* out_counts;       // [Q] or nullptr
    const int32_t* overflow;   // [Q] candidate buffer overflowed
    const float* thr0;         // [Q] seed bound used by the filter (-inf: every row of the shard was a candidate)
    int32_t* flags;            // [Q] 1 = not proven exact, needs the exact scan
    int32_t* n_flagged;        // running count of flagged queries
    float eps;                 // bound on |approx cosine - exact cosine|
    int D, Q, k, mode;
    const int32_t* qmap;       // nullptr, or [Q] query indices: candidate list s belongs to query qmap[s] (escalation of unproven queries)
};

// One warp per query: exact fp32 cosine of the C candidates with EXACTLY the arithmetic of the exact scan (scan_t8_kernel, scan.cuh:
// chunk j of a row belongs to group j % 8; group partial = sequential packed FMA over its chunks, (x,y) then (z,w), lo + hi;
// numerator = ((((p0 + p1) + p2) + ...) + p7); one divide), so the filter path and the exact scan return bit-identical scores.
// Final order (score desc, id asc), proof check.  Lane l owns candidates l, l + 32, ... (CPL = C / 32 of them) and computes the
// whole dot product of each of them itself (dim % 32 == 0 on this path).
template <int CPL>
__global__ void __launch_bounds__(256) scan_rescore_kernel(RescoreParams p) {
    constexpr int C = 32 * CPL;
    const int slot = blockIdx.x * 8 + (threadIdx.x >> 5);  // candidate list
    if (slot >= p.Q) return;
    const int qi = p.qmap ? p.qmap[slot] : slot;            // query
    const int lane = threadIdx.x & 31;
    const float qn = p.qnorms[qi];
    uint64_t my_gid[CPL];
    float my_approx[CPL], my_s[CPL];
    int ncand = 0;
    const float* q = p.queries + static_cast<size_t>(qi) * p.D;
#pragma unroll
    for (int u = 0; u < CPL; ++u) {
        my_gid[u] = p.cand_ids[static_cast<size_t>(slot) * C + u * 32 + lane];
        my_approx[u] = p.cand_scores[static_cast<size_t>(slot) * C + u * 32 + lane];
        my_s[u] = -INFINITY;
        ncand += __popc(__ballot_sync(0xffffffffu, my_gid[u] != kNoId64));  // candidates are packed at the front (sorted lists)
        if (my_gid[u] != kNoId64) {
            const size_t r = static_cast<size_t>(my_gid[u] - p.id_base);
            const float* rp = p.rows + r * p.D;
            uint64_t a2[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) a2[g] = 0ull;
            for (int c0 = 0; c0 < p.D; c0 += 32) {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(rp + c0 + 4 * g);
                    const ulonglong2 f = __ldg(reinterpret_cast<const ulonglong2*>(q + c0 + 4 * g));
                    a2[g] = f2_fma(v.x, f.x, a2[g]);
                    a2[g] = f2_fma(v.y, f.y, a2[g]);
                }
            }
            float acc = 0.f;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float lo, hi;
                f2_unpack(a2[g], lo, hi);
                acc = g == 0 ? lo + hi : acc + (lo + hi);
            }
            const float rn = p.norms[r];
            if (p.mode == SCAN_SEGMENT) my_s[u] = rn < 1e-9f ? 0.0f : acc / (qn * rn);
            else my_s[u] = acc / fmaxf(qn * rn, 1e-9f);
        }
    }
    // rank among the candidates by (score desc, id asc)
    int rank[CPL];
#pragma unroll
    for (int u = 0; u < CPL; ++u) rank[u] = 0;
#pragma unroll 1
    for (int j = 0; j < ncand; ++j) {
        float sj = 0.f;
        uint64_t ij = 0;
#pragma unroll
        for (int u = 0; u < CPL; ++u) {
            const float ts = __shfl_sync(0xffffffffu, my_s[u], j & 31);
            const uint64_t ti = __shfl_sync(0xffffffffu, my_gid[u], j & 31);
            if ((j >> 5) == u) { sj = ts; ij = ti; }
        }
#pragma unroll
        for (int u = 0; u < CPL; ++u)
            if (sj > my_s[u] || (sj == my_s[u] && ij < my_gid[u])) ++rank[u];
    }
    const bool empty_query = p.mode == SCAN_SEGMENT && qn < 1e-9f;
    const int nres = empty_query ? 0 : min(p.k, ncand);
    uint64_t* oi = p.out_ids + static_cast<size_t>(qi) * p.k;
    float* os = p.out_scores + static_cast<size_t>(qi) * p.k;
    float kth = 0.f;
    bool have_kth = false;
#pragma unroll
    for (int u = 0; u < CPL; ++u) {
        const bool valid = u * 32 + lane < ncand;
        if (valid && rank[u] < nres) { oi[rank[u]] = my_gid[u]; os[rank[u]] = my_s[u]; }
        const uint32_t kth_mask = __ballot_sync(0xffffffffu, valid && rank[u] == p.k - 1);
        if (kth_mask) {
            kth = __shfl_sync(0xffffffffu, my_s[u], __ffs(kth_mask) - 1);
            have_kth = true;
        }
    }
    for (int j = nres + lane; j < p.k; j += 32) { oi[j] = kNoId64; os[j] = -INFINITY; }
    if (lane == 0 && p.out_counts) p.out_counts[qi] = nres;
    // proof: rows outside the list have approximate cosine <= m_C (the smallest kept one) => exact <= m_C + eps
    const float m_last = __shfl_sync(0xffffffffu, my_approx[CPL - 1], 31);  // lists are sorted descending: last = smallest
    if (lane == 0 && !empty_query) {
        // fewer than C candidates: the list holds EVERY row whose approximate score beat the seed bound thr0, so the rows outside
        // it have approximate cosine <= thr0 / |q| (thr0 = -inf, unseeded: there are no rows outside)
        const float thr0 = p.thr0[qi];
        bool proven = false;
        if (ncand < C) proven = thr0 == -INFINITY || (have_kth && qn >= 1e-9f && kth > thr0 / qn + p.eps);
        if (ncand == C && have_kth && qn >= 1e-9f) proven = kth > m_last / qn + p.eps;
        if (p.overflow[qi]) proven = false;
        p.flags[qi] = proven ? 0 : 1;
        if (!proven) atomicAdd(p.n_flagged, 1);
    } else if (lane == 0) {
        p.flags[qi] = 0;
    }
}

}  // namespace kj
