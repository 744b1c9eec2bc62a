// CTA-pair (cta_group::2) tcgen05 GEMM with the A tile resident in shared memory, for the encoder's K <= 384
// projections (QKV and FFN-up at hidden size 384):
//     C[M,N] (bf16) = act( A[M,K] (bf16, K-major) x W[N,K]^T (bf16, K-major) + bias )
// Same reference slot as gemm_tcgen05.cuh (LinearLayer::matmul*, kjarni-transformers/src/linear_layer/linear_layer.rs:160-282
// -> cpu/ops/matmul.rs:370-479 -> cpu/kernels/x86/f32.rs:9-124, bias + activation of cpu/feedforward/standard_new.rs:65-73).
//
// Why a second GEMM kernel: the 1-CTA kernel streams A and W for every 128 x BN tile and is bound by the L2 -> SM path
// (~43 B/clk/SM; a 128 x 192 tile asks for 104 B/clk at full tensor rate).  Here
//   * two CTAs on one TPC share every weight tile: each loads half of its rows (BN/2) and one tcgen05.mma.cta_group::2
//     (M = 256: 128 rows per CTA) reads both halves, so weight traffic per CTA halves;
//   * the CTA's 128 x K activation tile is loaded once (K/64 k-blocks of 16 KB) and reused for all N/BN column tiles.
// Shared-memory traffic per CTA drops from (16 + BN/8) KB to BN/16 KB per k-block, e.g. 40 KB -> 12 KB for QKV.
//
// Roles per CTA (384 threads): warp 0 TMA producer (own A rows, own half of W), warp 1 MMA issuer (leader CTA only),
// warp 2 TMEM allocator, warps 4-11 epilogue (TMEM -> registers -> bias/activation -> bf16 -> swizzled smem -> TMA store).
// Barriers: the leader's full barriers count the bytes of BOTH CTAs' loads; tcgen05.commit multicasts to both CTAs'
// empty / accumulator-full barriers; the peer's epilogue warps release the accumulator on the leader's barrier.
#pragma once
#include <cuda.h>

#include "gemm_tcgen05.cuh"

namespace kj {

constexpr int kPairMaxKB = 6;  // K <= 384

template <int BN>
struct PairCfg {
    static constexpr int kStages = 5;
    static constexpr int kABlockBytes = kGemmBlockM * kGemmBlockK * 2;  // 16 KB per k-block, resident
    static constexpr int kABytes = kPairMaxKB * kABlockBytes;           // 96 KB
    static constexpr int kBBytes = (BN / 2) * kGemmBlockK * 2;          // this CTA's half of a weight k-block
    static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
    // epilogue geometry as in GemmCfg: 192-column tiles -> 12 warps, each stages a 32 x 64 part and issues one TMA store per tile
    // 256-column tiles: 4 parts of 64 columns = 16 epilogue warps (the epilogue is a per-warp latency chain of ~1 us per 32-column
    // chunk under load, so the parallelism has to come from more warps), each staging one 32 x 32 chunk at a time
    static constexpr int kParts = (BN == 192) ? 3 : ((BN == 256) ? 4 : 2);
    static constexpr int kEpiWarpsN = 4 * kParts;
    static constexpr int kThreads = 128 + 32 * kEpiWarpsN;
    static constexpr int kColsPerPart = BN / kParts;
    static constexpr int kStoreCols = (BN == 192) ? 64 : kEpiChunkCols;
    static constexpr int kEpiBufs = (BN == 192 || BN == 256) ? 1 : 2;
    static constexpr int kEpiBufBytes = 32 * kStoreCols * 2;
    static constexpr int kEpiBytes = kEpiWarpsN * kEpiBufs * kEpiBufBytes;
    static constexpr int kBiasBytes = kEpiBiasMax * 4;
    static constexpr int kSmemBytes = kABytes + kStages * kBBytes + kEpiBytes + 256 + kBiasBytes;
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PairCfg<BN>::kThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, GemmParams p) {
    using Cfg = PairCfg<BN>;
    constexpr int kStages = Cfg::kStages;
    static_assert(BN % 32 == 0 && BN >= 64 && BN <= 256, "BN");
    static_assert(EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_ACT_BF16, "pair kernel stores bf16");

    extern __shared__ __align__(1024) uint8_t smem_pair[];
    if (threadIdx.x == 0) KJ_TRACE(0);
    if (smem_u32(smem_pair) & 1023) __trap();
    uint8_t* smem_a = smem_pair;
    uint8_t* smem_b = smem_a + Cfg::kABytes;
    uint8_t* smem_epi = smem_b + kStages * Cfg::kBBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::kEpiBytes);
    uint64_t* full_bar = bars;                        // [stages]   (leader's are used)
    uint64_t* empty_bar = full_bar + kStages;         // [stages]
    uint64_t* a_full = empty_bar + kStages;           // [6]        (leader's are used)
    uint64_t* a_empty = a_full + kPairMaxKB;          // [6]
    uint64_t* tmem_full = a_empty + kPairMaxKB;       // [2]
    uint64_t* tmem_empty = tmem_full + 2;             // [2]        (leader's are used)
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
    const bool bias_in_smem = p.bias != nullptr && p.N <= kEpiBiasMax;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (bias_in_smem)  // a weight: staged before pdl_wait()
        for (int i = threadIdx.x; i < p.N; i += Cfg::kThreads) s_bias[i] = __ldg(p.bias + i);
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = static_cast<int>(cluster_id_x());
    const int n_pairs = static_cast<int>(cluster_nctaid_x());

    const int m_tiles = (p.M + 2 * kGemmBlockM - 1) / (2 * kGemmBlockM);  // 256-row tiles
    const int n_tiles = (p.N + BN - 1) / BN;
    const int k_blocks = (p.K + kGemmBlockK - 1) / kGemmBlockK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        tma_prefetch_desc(&tmap_c);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < kPairMaxKB; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 2 * Cfg::kEpiWarpsN);  // the epilogue warps of both CTAs
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_2sm<Cfg::kTmemCols>(tmem_base_smem);
    tc_fence_before();
    cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / TMA completion
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    if (threadIdx.x == 0) KJ_TRACE(1);
    pdl_wait();               // everything above overlapped the previous kernel's tail; its outputs are visible from here on
    pdl_launch_dependents();  // the next kernel may begin its own prologue as soon as this CTA's resources are released
    if (threadIdx.x == 0) KJ_TRACE(2);

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, ai = 0;
            for (int mt = pair; mt < m_tiles; mt += n_pairs, ++ai) {
                const int row0 = mt * 2 * kGemmBlockM + static_cast<int>(rank) * kGemmBlockM;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&a_empty[kb], (ai & 1) ^ 1);
                    if (leader) mbar_arrive_expect_tx(&a_full[kb], 2 * Cfg::kABlockBytes);
                    tma_load_2d_2sm(smem_a + kb * Cfg::kABlockBytes, &tmap_a, mapa_shared(smem_u32(&a_full[kb]), 0), kb * kGemmBlockK, row0,
                                    kEvictFirst);
                }
                for (int nb = 0; nb < n_tiles; ++nb) {
                    const int wrow0 = nb * BN + static_cast<int>(rank) * (BN / 2);
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kBBytes);
                        tma_load_2d_2sm(smem_b + stage * Cfg::kBBytes, &tmap_b, mapa_shared(smem_u32(&full_bar[stage]), 0), kb * kGemmBlockK,
                                        wrow0, kEvictLast);
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------- MMA issuer (leader CTA only)
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = umma_idesc(1 /*bf16*/, 2 * kGemmBlockM, BN);
            int stage = 0, it = 0;
            uint32_t phase = 0, ai = 0;
            for (int mt = pair; mt < m_tiles; mt += n_pairs, ++ai) {
                for (int nb = 0; nb < n_tiles; ++nb, ++it) {
                    const int acc = it & 1;
                    mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * BN;
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        if (nb == 0) mbar_wait(&a_full[kb], ai & 1);
                        mbar_wait(&full_bar[stage], phase);
                        if (it == 0 && kb == 0) KJ_TRACE(3);
                        tc_fence_after();
                        const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + kb * Cfg::kABlockBytes));
                        const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
                        for (int k = 0; k < kGemmBlockK / 16; ++k) umma_f16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        umma_commit_2sm(&empty_bar[stage], 3);
                        if (nb == n_tiles - 1) umma_commit_2sm(&a_empty[kb], 3);
                        if (kb == k_blocks - 1) umma_commit_2sm(&tmem_full[acc], 3);
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp >= kGemmEpiWarp0) {
        // ---------------------------------------------------------- epilogue
        const int ew = warp - kGemmEpiWarp0;
        const int quad = warp & 3;            // TMEM lane quadrant this warp may access
        const int part = ew >> 2;             // column part handled by this warpgroup
        constexpr int kColsPerPart = Cfg::kColsPerPart;
        uint8_t* stage_buf = smem_epi + ew * Cfg::kEpiBufs * Cfg::kEpiBufBytes;
        const uint32_t leader_empty0 = mapa_shared(smem_u32(&tmem_empty[0]), 0);
        const uint32_t leader_empty1 = mapa_shared(smem_u32(&tmem_empty[1]), 0);
        int sbuf = 0, it = 0;
        for (int mt = pair; mt < m_tiles; mt += n_pairs) {
            const int row0 = mt * 2 * kGemmBlockM + static_cast<int>(rank) * kGemmBlockM + quad * 32;
            for (int nb = 0; nb < n_tiles; ++nb, ++it) {
                const int acc = it & 1;
                mbar_wait(&tmem_full[acc], (it >> 1) & 1);
                if (ew == 0 && lane == 0 && it < 6) KJ_TRACE(4 + 2 * it);
                tc_fence_after();
                const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + part * kColsPerPart;
                const int colp = nb * BN + part * kColsPerPart;
                constexpr int kChunks = kColsPerPart / 32;
                if constexpr (Cfg::kStoreCols == 64) {
                    if (lane == 0) bulk_wait_read<0>();  // the previous tile's store has read the staging tile
                    __syncwarp();
                }
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr0 + c * 32, v);
                    tmem_ld_wait();
                    if (c + 1 == kChunks) {  // last load landed: release the accumulator (on the leader's barrier)
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(acc ? leader_empty1 : leader_empty0);
                    }
                    const int col0 = colp + c * 32;
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    if (p.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (col0 + 4 * j < p.N) {
                                float4 b;
                                if (bias_in_smem) b = *reinterpret_cast<const float4*>(s_bias + col0 + 4 * j);
                                else b = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
                                f[4 * j + 0] += b.x;
                                f[4 * j + 1] += b.y;
                                f[4 * j + 2] += b.z;
                                f[4 * j + 3] += b.w;
                            }
                        }
                    }
                    if (EPI == EPI_BIAS_ACT_BF16) apply_act_tile(f, p.act);
                    if constexpr (Cfg::kStoreCols == 64) {
                        const uint32_t rbase = smem_u32(stage_buf) + lane * 128;
                        const uint32_t sw = lane & 7;  // 128B swizzle
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            st_shared_v4(rbase + ((static_cast<uint32_t>(c * 4 + j) ^ sw) << 4), pack_bf16(f[8 * j + 0], f[8 * j + 1]),
                                         pack_bf16(f[8 * j + 2], f[8 * j + 3]), pack_bf16(f[8 * j + 4], f[8 * j + 5]),
                                         pack_bf16(f[8 * j + 6], f[8 * j + 7]));
                        }
                    } else if (col0 < p.N) {
                        if (lane == 0) bulk_wait_read<Cfg::kEpiBufs - 1>();
                        __syncwarp();
                        uint8_t* buf = stage_buf + sbuf * Cfg::kEpiBufBytes;
                        const uint32_t rbase = smem_u32(buf) + lane * 64;
                        const uint32_t sw = (lane >> 1) & 3;  // 64B swizzle
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            st_shared_v4(rbase + ((j ^ sw) << 4), pack_bf16(f[8 * j + 0], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                         pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tmap_c, buf, col0, row0);
                            bulk_commit();
                        }
                        if (++sbuf == Cfg::kEpiBufs) sbuf = 0;
                    }
                }
                if constexpr (Cfg::kStoreCols == 64) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0 && colp < p.N) {
                        tma_store_2d(&tmap_c, stage_buf, colp, row0);
                        bulk_commit();
                    }
                }
                if (ew == 0 && lane == 0 && it < 6) KJ_TRACE(5 + 2 * it);
            }
        }
        if (lane == 0) bulk_wait_read<0>();
        if (ew == 0 && lane == 0) KJ_TRACE(16);
    }

    tc_fence_before();
    cluster_sync_all();  // peer smem / barriers stay valid until both CTAs are done
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
    }
    if (threadIdx.x == 64) KJ_TRACE(17);
}

}  // namespace kj
