// Projection + bias + residual + LayerNorm in one kernel, for hidden size 384 (MiniLM):
//     x_out[M,384] (bf16) = LayerNorm( A[M,K] x W[384,K]^T + bias + x_res[M,384] ; gamma, beta, eps )
// Fuses  attention out-proj -> residual add -> LN1   (reference: cpu/encoder/encoder_layer.rs:120-147) and
//        FFN down-proj      -> residual add -> LN2   (cpu/feedforward/standard_new.rs:76-79, encoder_layer.rs:150-176),
// i.e. LinearLayer::matmul_noalloc + the scalar residual loop + LayerNorm::forward_noalloc
// (cpu/normalization/layer_norm.rs:37-134: per-row mean, biased variance, eps inside the sqrt).
//
// One CTA owns full rows: tile 128 x 384, fp32 accumulators in 384 TMEM columns (two 128x192x16 tcgen05.mma per
// k-step), so the row statistics never leave the SM.  Epilogue (8 warps, thread = row, 2 column halves):
//   pass 1  v = acc + bias + residual (residual tile TMA-loaded into swizzled smem), row sum; v written back to TMEM
//   pass 2  sum of (v - mean)^2   (exact two-pass variance, as the reference)
//   pass 3  (v - mean) * rstd * gamma + beta -> bf16 -> swizzled smem -> TMA store
// The residual may alias the output (in place): every warp reads its whole region before it writes it.
#pragma once
#include <cuda.h>

#include "gemm_tcgen05.cuh"

namespace kj {

constexpr int kLnN = 384;
constexpr int kLnHalfN = 192;
constexpr int kLnStages = 3;
constexpr int kLnABytes = kGemmBlockM * kGemmBlockK * 2;   // 16 KB
constexpr int kLnBBytes = kLnN * kGemmBlockK * 2;          // 48 KB (two TMA boxes of 192 rows)
constexpr int kLnStageBytes = kLnABytes + kLnBBytes;       // 64 KB
constexpr int kLnEpiBytesPerWarp = 2 * kEpiStageBytes;     // 2 x 2 KB: residual double buffer, then store double buffer
constexpr int kLnStatBytes = 2 * 2 * 128 * 4;              // [pass][half][row] partial sums
constexpr int kLnSmemBytes = kLnStages * kLnStageBytes + kEpiWarps * kLnEpiBytesPerWarp + kLnStatBytes + 256;  // 231,680 B of the 232,448 B limit

struct GemmLnParams {
    int M, K;
    const float* bias;   // [384] or nullptr
    const float* gamma;  // [384]
    const float* beta;   // [384]
    float eps;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_ln384_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_out, GemmLnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_ln[];  // no slack left for manual alignment: checked below
    uint8_t* smem = smem_ln;
    if (smem_u32(smem) & 1023) __trap();
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kLnStages * kLnABytes;
    uint8_t* smem_epi = smem + kLnStages * kLnStageBytes;
    float* stat = reinterpret_cast<float*>(smem_epi + kEpiWarps * kLnEpiBytesPerWarp);  // [2][2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stat) + kLnStatBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kLnStages;
    uint64_t* tmem_full = bars + 2 * kLnStages;
    uint64_t* tmem_empty = bars + 2 * kLnStages + 1;
    uint64_t* res_bar = bars + 2 * kLnStages + 2;  // [8 warps][2 buffers]
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(res_bar + 2 * kEpiWarps);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_tiles = (p.M + kGemmBlockM - 1) / kGemmBlockM;
    const int k_blocks = (p.K + kGemmBlockK - 1) / kGemmBlockK;

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
        mbar_init(tmem_empty, kEpiWarps);
        for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();               // everything above overlapped the previous kernel's tail; its outputs are visible from here on
    pdl_launch_dependents();  // the next kernel may begin its own prologue as soon as this CTA's resources are released

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], kLnStageBytes);
                    tma_load_2d(smem_a + stage * kLnABytes, &tmap_a, &full_bar[stage], kb * kGemmBlockK, tile * kGemmBlockM, kEvictFirst);
                    tma_load_2d(smem_b + stage * kLnBBytes, &tmap_w, &full_bar[stage], kb * kGemmBlockK, 0, kEvictLast);
                    tma_load_2d(smem_b + stage * kLnBBytes + kLnHalfN * 128, &tmap_w, &full_bar[stage], kb * kGemmBlockK, kLnHalfN, kEvictLast);
                    if (++stage == kLnStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(1 /*bf16*/, kGemmBlockM, kLnHalfN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
                mbar_wait(tmem_empty, (it & 1) ^ 1);  // single accumulator: the previous tile's epilogue must have drained it
                tc_fence_after();
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + stage * kLnABytes));
                    const uint64_t db0 = umma_desc_k_sw128(smem_u32(smem_b + stage * kLnBBytes));
                    const uint64_t db1 = umma_desc_k_sw128(smem_u32(smem_b + stage * kLnBBytes + kLnHalfN * 128));
#pragma unroll
                    for (int k = 0; k < kGemmBlockK / 16; ++k) {
                        umma_f16(tmem_base, da + 2 * k, db0 + 2 * k, idesc, (kb | k) != 0);
                        umma_f16(tmem_base + kLnHalfN, da + 2 * k, db1 + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kb == k_blocks - 1) umma_commit(tmem_full);
                    if (++stage == kLnStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp >= kGemmEpiWarp0) {
        // ---------------------------------------------------------- epilogue
        const int ew = warp - kGemmEpiWarp0;
        const int quad = warp & 3;
        const int half = ew >> 2;
        constexpr int kChunks = kLnHalfN / kEpiChunkCols;  // 6
        uint8_t* ebuf = smem_epi + ew * kLnEpiBytesPerWarp;
        uint64_t* rbar = res_bar + 2 * ew;
        const uint32_t sw = (lane >> 1) & 3;  // 64B swizzle of this lane's row
        const int trow = quad * 32 + lane;    // row inside the tile
        uint32_t rphase = 0;                  // bit b = parity of rbar[b]
        int it = 0;
        for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
            const int row0 = tile * kGemmBlockM + quad * 32;
            const int col_base = half * kLnHalfN;
            // the store staging of the previous tile aliases the residual buffers: wait until TMA has read it
            if (lane == 0) {
                bulk_wait_read<0>();
                // prefetch the first two residual chunks; they land while the MMAs of this tile run
                for (int c = 0; c < 2; ++c) {
                    mbar_arrive_expect_tx(&rbar[c], kEpiStageBytes);
                    tma_load_2d(ebuf + c * kEpiStageBytes, &tmap_res, &rbar[c], col_base + c * kEpiChunkCols, row0, kEvictFirst);
                }
            }
            __syncwarp();
            mbar_wait(tmem_full, it & 1);
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + col_base;

            // ---- pass 1: v = acc + bias + residual ; row sum ; v -> TMEM
            float sum = 0.0f;
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
                const int col0 = col_base + c * kEpiChunkCols;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t w[4] = {r4[j].x, r4[j].y, r4[j].z, r4[j].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
                        float a0 = __uint_as_float(v[8 * j + 2 * e]) + lo;
                        float a1 = __uint_as_float(v[8 * j + 2 * e + 1]) + hi;
                        if (p.bias != nullptr) {
                            const float2 bb = __ldg(reinterpret_cast<const float2*>(p.bias + col0 + 8 * j + 2 * e));
                            a0 += bb.x;
                            a1 += bb.y;
                        }
                        sum += a0 + a1;
                        v[8 * j + 2 * e] = __float_as_uint(a0);
                        v[8 * j + 2 * e + 1] = __float_as_uint(a1);
                    }
                }
                tmem_st_32x32(taddr0 + c * kEpiChunkCols, v);
                __syncwarp();  // every lane has read residual buffer b
                if (lane == 0 && c + 2 < kChunks) {
                    mbar_arrive_expect_tx(&rbar[b], kEpiStageBytes);
                    tma_load_2d(ebuf + b * kEpiStageBytes, &tmap_res, &rbar[b], col_base + (c + 2) * kEpiChunkCols, row0, kEvictFirst);
                }
            }
            tmem_st_wait();
            stat[(0 * 2 + half) * 128 + trow] = sum;
            named_bar_sync(1, kEpiWarps * 32);
            const float mean = (stat[(0 * 2 + 0) * 128 + trow] + stat[(0 * 2 + 1) * 128 + trow]) * (1.0f / kLnN);

            // ---- pass 2: sum of squared deviations
            float sq = 0.0f;
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(taddr0 + c * kEpiChunkCols, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float d = __uint_as_float(v[j]) - mean;
                    sq = fmaf(d, d, sq);
                }
            }
            stat[(1 * 2 + half) * 128 + trow] = sq;
            named_bar_sync(1, kEpiWarps * 32);
            const float var = (stat[(1 * 2 + 0) * 128 + trow] + stat[(1 * 2 + 1) * 128 + trow]) * (1.0f / kLnN);
            const float rstd = 1.0f / sqrtf(var + p.eps);

            // ---- pass 3: normalise, bf16, swizzled smem, TMA store
            int sbuf = 0;
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(taddr0 + c * kEpiChunkCols, v);
                tmem_ld_wait();
                const int col0 = col_base + c * kEpiChunkCols;
                float f[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + col0) + j);
                    const float4 bt = __ldg(reinterpret_cast<const float4*>(p.beta + col0) + j);
                    f[4 * j + 0] = (__uint_as_float(v[4 * j + 0]) - mean) * rstd * g.x + bt.x;
                    f[4 * j + 1] = (__uint_as_float(v[4 * j + 1]) - mean) * rstd * g.y + bt.y;
                    f[4 * j + 2] = (__uint_as_float(v[4 * j + 2]) - mean) * rstd * g.z + bt.z;
                    f[4 * j + 3] = (__uint_as_float(v[4 * j + 3]) - mean) * rstd * g.w + bt.w;
                }
                if (lane == 0) bulk_wait_read<1>();
                __syncwarp();
                uint8_t* buf = ebuf + sbuf * kEpiStageBytes;
                const uint32_t obase = smem_u32(buf) + lane * 64;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    st_shared_v4(obase + ((j ^ sw) << 4), pack_bf16(f[8 * j + 0], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                 pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmap_out, buf, col0, row0);
                    bulk_commit();
                }
                sbuf ^= 1;
            }
            // accumulator fully consumed: the MMA warp may start the next tile
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
            // the second stat exchange of this tile and the first of the next use different slots; a slow warp still
            // reading stat[1] cannot be overtaken by a write to stat[1] (that needs two more barriers), so no extra sync.
        }
        if (lane == 0) bulk_wait_read<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace kj
