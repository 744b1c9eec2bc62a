// Projection + bias + residual + LayerNorm in one kernel, for hidden size 384 (MiniLM) and 768 (DistilBERT / BERT-base):
//     x_out[M,H] (bf16) = LayerNorm( A[M,K] x W[H,K]^T + bias + x_res[M,H] ; gamma, beta, eps ),  H = 384 * kCluster
// Fuses  attention out-proj -> residual add -> LN1   (reference: cpu/encoder/encoder_layer.rs:120-147) and
//        FFN down-proj      -> residual add -> LN2   (cpu/feedforward/standard_new.rs:76-79, encoder_layer.rs:150-176),
// i.e. LinearLayer::matmul_noalloc + the scalar residual loop + LayerNorm::forward_noalloc
// (cpu/normalization/layer_norm.rs:37-134: per-row mean, biased variance, eps inside the sqrt).
//
// One CTA owns full rows: tile 128 x 384, fp32 accumulators in 384 TMEM columns (two 128x192x16 tcgen05.mma per
// k-step), so the row statistics never leave the SM.  Epilogue: 12 warps = 4 TMEM lane quadrants x 3 column parts of 128,
// thread = row (the epilogue is issue/latency-bound, so it gets three warps per scheduler), two passes:
//   pass A  v = acc + bias + residual (residual chunks TMA-loaded into swizzled smem), row sum and sum of squares;
//           v written back to TMEM
//   pass B  (v - mean) * rstd * gamma + beta -> bf16 -> swizzled smem -> TMA store
// Variance = E[v^2] - mean^2 in fp32 (clamped at 0): with |mean| <~ 10 sigma its rounding error is ~1e-5 relative,
// far below the bf16 rounding of the output.  The residual may alias the output (in place): every warp reads its whole
// region before it writes it.  bias / gamma / beta are staged in shared memory before griddepcontrol.wait.
//
// H = 768 (kCluster = 2): a thread-block cluster of two CTAs owns a 128-row tile, CTA `rank` computes columns
// [384 * rank, 384 * rank + 384) exactly as above (its own 512 TMEM columns on its own SM).  After pass A every epilogue thread
// stores its row's partial (sum, sum of squares) into the PEER's shared memory (st.shared::cluster) and arrives on the peer's
// mbarrier with release.cluster; the peer waits with acquire.cluster, so the full-row statistics cost one DSMEM round trip and
// nothing [M,768]-sized in fp32 ever goes to HBM (round 1 wrote fp32 sums and ran a separate layernorm_kernel).  The peer
// buffer and its mbarrier are double-buffered by tile parity: a CTA can be at most one exchange ahead of its peer.
#pragma once
#include <cuda.h>

#include "gemm_tcgen05.cuh"

namespace kj {

constexpr int kLnN = 384;
constexpr int kLnHalfN = 192;
#ifndef KJ_LN_STAGES
#define KJ_LN_STAGES 3
#endif
constexpr int kLnStages = KJ_LN_STAGES;
// With 3 operand stages (192 KB) the epilogue's staging buffers no longer fit beside the ring, so they ALIAS ring stage 0: the
// epilogue only touches them after tmem_full (every MMA of the tile has retired, the ring is idle), and the producer waits on
// `epi_free` before it loads the next tile's operands.  Two 64 KB stages left the K-loop TMA-latency-bound (one k-block of
// MMAs, 0.6 us, could not cover a load round trip under load): 30 us per FFN-down launch against a 14 us MMA floor.
constexpr bool kLnAliasEpi = kLnStages >= 3;
constexpr int kLnParts = 3;                                // column parts per lane quadrant
constexpr int kLnPartCols = kLnN / kLnParts;               // 128
constexpr int kLnEpiWarps = 4 * kLnParts;                  // 12
constexpr int kLnThreads = 128 + 32 * kLnEpiWarps;         // 512
constexpr int kLnABytes = kGemmBlockM * kGemmBlockK * 2;   // 16 KB
constexpr int kLnBBytes = kLnN * kGemmBlockK * 2;          // 48 KB (two TMA boxes of 192 rows)
constexpr int kLnStageBytes = kLnABytes + kLnBBytes;       // 64 KB
constexpr int kLnEpiBytesPerWarp = 2 * kEpiStageBytes;     // 2 x 2 KB: residual double buffer, then store double buffer
constexpr int kLnStatBytes = kLnParts * 128 * 8;           // [part][row] (sum, sum of squares)
constexpr int kLnVecBytes = 3 * kLnN * 4;                  // bias | gamma | beta
constexpr int kLnPeerStatBytes = 2 * kLnStatBytes;         // the peer CTA's partial sums, double-buffered by tile parity (kCluster = 2)
constexpr int kLnSmemBytes = kLnStages * kLnStageBytes + (kLnAliasEpi ? 0 : kLnEpiWarps * kLnEpiBytesPerWarp) + kLnStatBytes + kLnPeerStatBytes + kLnVecBytes + 512;
static_assert(kLnEpiWarps * kLnEpiBytesPerWarp <= kLnStageBytes, "aliased epilogue staging must fit in one ring stage");
static_assert(kLnSmemBytes <= 232448, "shared memory budget");

// The LayerNorm epilogue is issue-bound (12 warps x 128 columns per row tile, scripts/chain_trace.py): its arithmetic runs on the
// packed fp32x2 pipe.  Shared by gemm_ln_kernel and gemm_ln_gemm_kernel so that both produce the same bits.
// Pass A, one 32-column chunk: v = acc + residual + bias (per element exactly as in round 1: (acc + res) + bias), row sum and sum of
// squares accumulated as (even columns, odd columns) pairs -- the caller adds the two halves at the end.
__device__ __forceinline__ void ln_pass_a_chunk(uint32_t (&v)[32], const uint4 (&r4)[4], const float* bs, uint64_t& sum2, uint64_t& sq2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t w[4] = {r4[j].x, r4[j].y, r4[j].z, r4[j].w};
        const float4 b0 = *reinterpret_cast<const float4*>(bs + 8 * j);
        const float4 b1 = *reinterpret_cast<const float4*>(bs + 8 * j + 4);
        const uint64_t bb[4] = {f2_pack(b0.x, b0.y), f2_pack(b0.z, b0.w), f2_pack(b1.x, b1.y), f2_pack(b1.z, b1.w)};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint64_t res = f2_pack(__uint_as_float(w[e] << 16), __uint_as_float(w[e] & 0xffff0000u));  // two bf16 -> two fp32
            const uint64_t a = f2_add(f2_add(f2_pack(__uint_as_float(v[8 * j + 2 * e]), __uint_as_float(v[8 * j + 2 * e + 1])), res), bb[e]);
            sum2 = f2_add(sum2, a);
            sq2 = f2_fma(a, a, sq2);
            float a0, a1;
            f2_unpack(a, a0, a1);
            v[8 * j + 2 * e] = __float_as_uint(a0);
            v[8 * j + 2 * e + 1] = __float_as_uint(a1);
        }
    }
}
// Pass B, one 32-column chunk: ((v * rstd + nmr) * gamma + beta) -> 16 packed bf16 pairs (per element the same two fmas as round 1)
__device__ __forceinline__ void ln_pass_b_chunk(const uint32_t (&v)[32], float rstd, float nmr, const float* g, const float* bt, uint32_t (&pk)[16]) {
    const uint64_t r2 = f2_pack(rstd, rstd), n2 = f2_pack(nmr, nmr);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 g4 = *reinterpret_cast<const float4*>(g + 4 * j);
        const float4 b4 = *reinterpret_cast<const float4*>(bt + 4 * j);
        float y0, y1, y2, y3;
        f2_unpack(f2_fma(f2_fma(f2_pack(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), r2, n2), f2_pack(g4.x, g4.y), f2_pack(b4.x, b4.y)), y0, y1);
        f2_unpack(f2_fma(f2_fma(f2_pack(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), r2, n2), f2_pack(g4.z, g4.w), f2_pack(b4.z, b4.w)), y2, y3);
        pk[2 * j] = pack_bf16(y0, y1);
        pk[2 * j + 1] = pack_bf16(y2, y3);
    }
}

struct GemmLnParams {
    int M, K;
    const float* bias;   // [H] or nullptr
    const float* gamma;  // [H]
    const float* beta;   // [H]
    float eps;
};

__device__ __forceinline__ void st_cluster_f2(uint32_t cluster_addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}

template <int kCluster>
__global__ void __launch_bounds__(kLnThreads, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_out, GemmLnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_ln[];
    uint8_t* smem = smem_ln;
    if (smem_u32(smem) & 1023) __trap();
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kLnStages * kLnABytes;
    uint8_t* smem_tail = smem + kLnStages * kLnStageBytes;
    // aliased staging = stage 0 of the W array: 48 KB, exactly 12 warps x 4 KB
    uint8_t* smem_epi = kLnAliasEpi ? smem_b : smem_tail;
    float2* stat = reinterpret_cast<float2*>(smem_tail + (kLnAliasEpi ? 0 : kLnEpiWarps * kLnEpiBytesPerWarp));  // [3][128]
    float2* peer_stat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(stat) + kLnStatBytes);  // [2][3][128], written by the peer CTA
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(peer_stat) + kLnPeerStatBytes);
    float* s_gamma = s_bias + kLnN;
    float* s_beta = s_gamma + kLnN;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_beta + kLnN);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kLnStages;
    uint64_t* tmem_full = bars + 2 * kLnStages;
    uint64_t* tmem_empty = bars + 2 * kLnStages + 1;
    uint64_t* res_bar = bars + 2 * kLnStages + 2;  // [12 warps][2 buffers]
    uint64_t* epi_free = res_bar + 2 * kLnEpiWarps;  // epilogue staging (aliased on the ring) released for the next tile's loads
    uint64_t* xstat_bar = epi_free + 1;              // [2] by tile parity (kCluster = 2): the peer's 384 epilogue threads arrive once per tile
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(xstat_bar + 2);
    static_assert(kCluster == 1 || kCluster == 2, "one CTA or a CTA pair per row tile");
    const uint32_t rank = kCluster == 1 ? 0u : cluster_ctarank();
    const int ncol0 = static_cast<int>(rank) * kLnN;   // this CTA's first output column
    const int tile0 = kCluster == 1 ? static_cast<int>(blockIdx.x) : static_cast<int>(cluster_id_x());
    const int tile_step = kCluster == 1 ? static_cast<int>(gridDim.x) : static_cast<int>(cluster_nctaid_x());

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_tiles = (p.M + kGemmBlockM - 1) / kGemmBlockM;
    const int k_blocks = (p.K + kGemmBlockK - 1) / kGemmBlockK;

    // weights: independent of the predecessor kernel, staged before griddepcontrol.wait
    for (int i = threadIdx.x; i < kLnN; i += kLnThreads) {
        s_bias[i] = p.bias != nullptr ? __ldg(p.bias + ncol0 + i) : 0.0f;
        s_gamma[i] = __ldg(p.gamma + ncol0 + i);
        s_beta[i] = __ldg(p.beta + ncol0 + i);
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_res);
        tma_prefetch_desc(&tmap_out);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kLnStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, kLnEpiWarps);
        for (int i = 0; i < 2 * kLnEpiWarps; ++i) mbar_init(&res_bar[i], 1);
        mbar_init(epi_free, kLnEpiWarps);
        mbar_init(&xstat_bar[0], kLnEpiWarps * 32);
        mbar_init(&xstat_bar[1], kLnEpiWarps * 32);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    if (kCluster > 1) cluster_sync_all();  // the peer's xstat_bar is initialised before anyone can arrive on it
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();               // everything above overlapped the previous kernel's tail; its outputs are visible from here on
    pdl_launch_dependents();  // the next kernel may begin its own prologue as soon as this CTA's resources are released

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = tile0; tile < m_tiles; tile += tile_step, ++it) {
                if (kLnAliasEpi && it > 0) mbar_wait(epi_free, (it - 1) & 1);  // the previous tile's epilogue is done with the ring
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], kLnStageBytes);
                    tma_load_2d(smem_a + stage * kLnABytes, &tmap_a, &full_bar[stage], kb * kGemmBlockK, tile * kGemmBlockM, kEvictFirst);
                    tma_load_2d(smem_b + stage * kLnBBytes, &tmap_w, &full_bar[stage], kb * kGemmBlockK, ncol0, kEvictLast);
                    tma_load_2d(smem_b + stage * kLnBBytes + kLnHalfN * 128, &tmap_w, &full_bar[stage], kb * kGemmBlockK, ncol0 + kLnHalfN, kEvictLast);
                    if (++stage == kLnStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer
        if (KJ_MMA_UNIFORM != 0 || lane == 0) {
            constexpr uint32_t idesc = umma_idesc(1 /*bf16*/, kGemmBlockM, kLnHalfN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = tile0; tile < m_tiles; tile += tile_step, ++it) {
                mbar_wait(tmem_empty, (it & 1) ^ 1);  // single accumulator: the previous tile's epilogue must have drained it
                tc_fence_after();
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + stage * kLnABytes));
                    const uint64_t db0 = umma_desc_k_sw128(smem_u32(smem_b + stage * kLnBBytes));
                    const uint64_t db1 = umma_desc_k_sw128(smem_u32(smem_b + stage * kLnBBytes + kLnHalfN * 128));
                    if (mma_issuer_lane()) {
#pragma unroll
                        for (int k = 0; k < kGemmBlockK / 16; ++k) {
                            umma_f16(tmem_base, da + 2 * k, db0 + 2 * k, idesc, (kb | k) != 0);
                            umma_f16(tmem_base + kLnHalfN, da + 2 * k, db1 + 2 * k, idesc, (kb | k) != 0);
                        }
                        umma_commit(&empty_bar[stage]);
                        if (kb == k_blocks - 1) umma_commit(tmem_full);
                    }
                    mma_issuer_sync();
                    if (++stage == kLnStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp >= kGemmEpiWarp0) {
        // ---------------------------------------------------------- epilogue
        const int ew = warp - kGemmEpiWarp0;  // 0..11
        const int quad = warp & 3;
        const int part = ew >> 2;             // column part [part*128, part*128 + 128)
        constexpr int kChunks = kLnPartCols / kEpiChunkCols;  // 4
        uint8_t* ebuf = smem_epi + ew * kLnEpiBytesPerWarp;
        uint64_t* rbar = res_bar + 2 * ew;
        const uint32_t sw = (lane >> 1) & 3;  // 64B swizzle of this lane's row
        const int trow = quad * 32 + lane;    // row inside the tile
        uint32_t rphase = 0;                  // bit b = parity of rbar[b]
        int it = 0;
        for (int tile = tile0; tile < m_tiles; tile += tile_step, ++it) {
            const int row0 = tile * kGemmBlockM + quad * 32;
            const int col_base = part * kLnPartCols;
            // the store staging of the previous tile aliases the residual buffers: wait until TMA has read it
            auto prefetch_residual = [&] {
                if (lane == 0) {
                    bulk_wait_read<0>();
                    for (int c = 0; c < 2; ++c) {
                        mbar_arrive_expect_tx(&rbar[c], kEpiStageBytes);
                        tma_load_2d(ebuf + c * kEpiStageBytes, &tmap_res, &rbar[c], ncol0 + col_base + c * kEpiChunkCols, row0, kEvictFirst);
                    }
                }
                __syncwarp();
            };
            // separate staging: the first two residual chunks land while the MMAs of this tile run; aliased staging: the ring is
            // only free once the accumulator is complete
            if (!kLnAliasEpi) prefetch_residual();
            mbar_wait(tmem_full, it & 1);
            tc_fence_after();
            if (kLnAliasEpi) prefetch_residual();
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + col_base;

            // ---- pass A: v = acc + bias + residual ; row sum and sum of squares ; v -> TMEM
            uint64_t sum2 = f2_pack(0.0f, 0.0f), sq2 = f2_pack(0.0f, 0.0f);
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c) {
                const int b = c & 1;
                uint32_t v[32];
                tmem_ld_32x32(taddr0 + c * kEpiChunkCols, v);
                mbar_wait(&rbar[b], (rphase >> b) & 1);
                rphase ^= 1u << b;
                const uint32_t rbase = smem_u32(ebuf + b * kEpiStageBytes) + lane * 64;
                uint4 r4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) r4[j] = ld_shared_v4(rbase + ((j ^ sw) << 4));
                tmem_ld_wait();
                ln_pass_a_chunk(v, r4, s_bias + col_base + c * kEpiChunkCols, sum2, sq2);
                tmem_st_32x32(taddr0 + c * kEpiChunkCols, v);
                __syncwarp();  // every lane has read residual buffer b
                if (lane == 0 && c + 2 < kChunks) {
                    mbar_arrive_expect_tx(&rbar[b], kEpiStageBytes);
                    tma_load_2d(ebuf + b * kEpiStageBytes, &tmap_res, &rbar[b], ncol0 + col_base + (c + 2) * kEpiChunkCols, row0, kEvictFirst);
                }
            }
            tmem_st_wait();
            float s1, s2;
            {
                float e0, e1, q0, q1;
                f2_unpack(sum2, e0, e1);
                f2_unpack(sq2, q0, q1);
                s1 = e0 + e1;
                s2 = q0 + q1;
            }
            stat[part * 128 + trow] = make_float2(s1, s2);
            if (kCluster > 1) {  // the same partial sums into the peer's buffer for this tile parity, then a releasing remote arrive
                const uint32_t peer = rank ^ 1u;
                st_cluster_f2(mapa_shared(smem_u32(peer_stat + (it & 1) * (kLnParts * 128) + part * 128 + trow), peer), s1, s2);
                mbar_arrive_cluster(mapa_shared(smem_u32(&xstat_bar[it & 1]), peer));
            }
            named_bar_sync(1, kLnEpiWarps * 32);
            float t1 = 0.0f, t2 = 0.0f;
#pragma unroll
            for (int q = 0; q < kLnParts; ++q) {
                const float2 t = stat[q * 128 + trow];
                t1 += t.x;
                t2 += t.y;
            }
            if (kCluster > 1) {
                mbar_wait_cluster(&xstat_bar[it & 1], (it >> 1) & 1);
                float u1 = 0.0f, u2 = 0.0f;
#pragma unroll
                for (int q = 0; q < kLnParts; ++q) {
                    const float2 t = peer_stat[(it & 1) * (kLnParts * 128) + q * 128 + trow];
                    u1 += t.x;
                    u2 += t.y;
                }
                // rank 0 adds (own + peer), rank 1 (peer + own): both CTAs use the same column order, hence the same statistics
                t1 = rank == 0 ? t1 + u1 : u1 + t1;
                t2 = rank == 0 ? t2 + u2 : u2 + t2;
            }
            constexpr float kInvN = 1.0f / (kLnN * kCluster);
            const float mean = t1 * kInvN;
            const float var = fmaxf(t2 * kInvN - mean * mean, 0.0f);
            const float rstd = 1.0f / sqrtf(var + p.eps);
            const float nmr = -mean * rstd;

            // ---- pass B: normalise, bf16, swizzled smem, TMA store
            int sbuf = 0;
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(taddr0 + c * kEpiChunkCols, v);
                tmem_ld_wait();
                if (c + 1 == kChunks) {  // accumulator fully consumed: the MMA warp may start the next tile
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty);
                }
                const int col0 = col_base + c * kEpiChunkCols;  // within this CTA's 384 columns
                uint32_t pk[16];
                ln_pass_b_chunk(v, rstd, nmr, s_gamma + col0, s_beta + col0, pk);
                if (lane == 0) bulk_wait_read<1>();
                __syncwarp();
                uint8_t* buf = ebuf + sbuf * kEpiStageBytes;
                const uint32_t obase = smem_u32(buf) + lane * 64;
#pragma unroll
                for (int j = 0; j < 4; ++j) st_shared_v4(obase + ((j ^ sw) << 4), pk[4 * j + 0], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmap_out, buf, ncol0 + col0, row0);
                    bulk_commit();
                }
                sbuf ^= 1;
            }
            // the stat exchange of the next tile happens after its tmem_full wait, i.e. after every warp of this tile has
            // passed the barrier above and read stat[]: no extra sync is needed before stat[] is overwritten.
            if (kLnAliasEpi && tile + tile_step < m_tiles) {  // hand the ring back to the producer
                if (lane == 0) {
                    bulk_wait_read<0>();
                    mbar_arrive(epi_free);
                }
                __syncwarp();
            }
        }
        if (lane == 0) bulk_wait_read<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (kCluster > 1) cluster_sync_all();  // no CTA exits while its peer may still write into its shared memory
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace kj
