// Whole feed-forward block + residual + LayerNorm in one kernel, for hidden size 384 (MiniLM):
//     x[M,384] (bf16, in place) = LayerNorm( x + act(x W1^T + b1) W2^T + b2 ; gamma, beta, eps )
// Fuses StdFeedForwardNew::forward_noalloc (reference: kjarni-transformers/src/cpu/feedforward/standard_new.rs:47-80,
// erf-GELU activations.rs:57-59), the residual add and LN2 of EncoderLayer::forward_postnorm_noalloc
// (cpu/encoder/encoder_layer.rs:150-176) and LayerNorm::forward_noalloc (cpu/normalization/layer_norm.rs:37-134).
//
// Why: as two kernels the [M, I] intermediate costs 2 x M x I x 2 bytes of L2 write + read traffic per layer, and the up
// projection was bound by its TMA store path (the stores of a 148 x 128 x 1536 tile set drain at ~18 GB/s per SM).  Here the
// intermediate never leaves the SM:
//   for each 64-wide chunk c of the intermediate dimension (I / 64 chunks):
//     G1  acc1[128 x 64]   = x_tile[128 x 384] . W1[c*64 .. c*64+64, :]^T           tcgen05.mma 128 x 64 x 16, x resident in smem
//     E1  h_c = act(acc1 + b1)  -> bf16 -> swizzled smem tile [128 x 64] (double-buffered)    8 epilogue warps
//     G2  acc2[128 x 384] += h_c . W2[:, c*64 .. c*64+64]^T                          3 x tcgen05.mma 128 x 128 x 16 per k-step
//   LN  v = acc2 + b2 + x_tile (residual read from the resident tile), two-pass LayerNorm, result written back into the
//       resident tile and stored with 6 TMA boxes.                                            12 epilogue warps
// acc1 is double-buffered and G1 runs one chunk ahead of G2 (issue order G1(c+1), G2(c)), so the G1 -> E1 -> G2 latency chain
// of one chunk is covered by tensor work of its neighbours.
// Weights stream through a 96 KB ring of TMA boxes in consumption order.
//
// Two variants.  kPair = false: one CTA per 128-row tile.  kPair = true (default): a CTA pair (cluster of 2, one TPC) runs
// tcgen05.mma.cta_group::2 with M = 256 -- the leader CTA issues every MMA for both tiles, each CTA keeps its own x tile,
// h buffers and accumulators, and holds only HALF of every weight tile (32 of the 64 W1 rows, 64 of each 128 W2 rows).  That
// halves the MMA instruction count per tile (the single issuing thread was the bottleneck: ~56 ns per 128x64x16 MMA against
// 16 ns of tensor work), halves the L2 -> SM weight traffic per CTA and makes the same 96 KB ring two chunks deep instead of
// one (with a one-chunk ring every slot's refill latency was exposed once per chunk).
// TMEM: acc2 = columns [0, 384), acc1 = two buffers at [384, 448) and [448, 512).
#pragma once
#include <cuda.h>

#include "gemm_tcgen05.cuh"

namespace kj {

constexpr int kFfH = 384;
constexpr int kFfChunk = 64;                              // intermediate units per chunk
constexpr int kFfRingBytes = 6 * 16384;                   // 96 KB weight ring
constexpr int kFfXBytes = 6 * 16384;                      // resident x tile: 6 k-blocks of [128 x 64] bf16
constexpr int kFfHBufBytes = 128 * kFfChunk * 2;          // 16 KB
constexpr int kFfActWarps = 8;                            // warps 4..11: activation epilogue
constexpr int kFfLnWarps = 12;                            // warps 4..15: LayerNorm epilogue
constexpr int kFfThreads = 128 + 32 * kFfLnWarps;         // 512
constexpr int kFfSmemBytes = kFfXBytes + 2 * kFfHBufBytes + kFfRingBytes + 512;  // 229,888 B
static_assert(kFfSmemBytes <= 232448, "shared memory budget");
constexpr uint32_t kFfAcc1Col = 384;

struct FfnParams {
    int M, I;            // rows, intermediate size (multiple of 64)
    const float* b1;     // [I] or nullptr
    const float* b2;     // [384] or nullptr
    const float* gamma;  // [384]
    const float* beta;   // [384]
    float eps;
    int act;             // Activation
};

template <bool kPair>
__global__ void __launch_bounds__(kFfThreads, 1)
ffn_ln384_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                 const __grid_constant__ CUtensorMap tmap_w2, FfnParams p) {
    constexpr int kSlots = kPair ? 12 : 6;
    constexpr int kSlotBytes = kFfRingBytes / kSlots;     // 8 KB (this CTA's half) or 16 KB
    constexpr int kW1Rows = kPair ? 32 : 64;              // W1 rows of a chunk held by this CTA
    constexpr int kNCta = kPair ? 2 : 1;
    extern __shared__ __align__(1024) uint8_t smem_ff[];
    if (smem_u32(smem_ff) & 1023) __trap();
    uint8_t* smem_x = smem_ff;
    uint8_t* smem_h = smem_x + kFfXBytes;                 // 2 x [128 x 64] bf16; LayerNorm statistics alias it at the end
    uint8_t* smem_w = smem_h + 2 * kFfHBufBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_w + kFfRingBytes);
    uint64_t* full_bar = bars;                   // [kSlots]  (pair: the leader's count both CTAs' bytes)
    uint64_t* empty_bar = full_bar + kSlots;     // [kSlots]
    uint64_t* x_local = empty_bar + kSlots;      // this CTA's x tile landed
    uint64_t* x_ready = x_local + 1;             // pair, leader's: both CTAs' x tiles landed (count 2)
    uint64_t* x_empty = x_ready + 1;             // LayerNorm output stored, tile may be replaced
    uint64_t* acc1_full = x_empty + 1;           // [2]
    uint64_t* acc1_empty = acc1_full + 2;        // [2] leader's: count 8 per CTA
    uint64_t* h_full = acc1_empty + 2;           // [2] leader's: count 8 per CTA
    uint64_t* h_empty = h_full + 2;              // [2]
    uint64_t* acc2_full = h_empty + 2;
    uint64_t* acc2_empty = acc2_full + 1;        // leader's: count 12 per CTA
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(acc2_empty + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_tiles = (p.M + kGemmBlockM - 1) / kGemmBlockM;
    const int n_chunks = p.I / kFfChunk;
    // tile schedule: CTA (or CTA `rank` of pair `clusterid`) handles tiles first, first + stride, ...; in pair mode a CTA may get a
    // tile index >= m_tiles (odd tile count): it runs on zero-filled rows and its stores are clipped.
    const uint32_t rank = kPair ? cluster_ctarank() : 0;
    const bool leader = rank == 0;
    const int tile_first = kPair ? 2 * static_cast<int>(cluster_id_x()) + static_cast<int>(rank) : static_cast<int>(blockIdx.x);
    const int tile_stride = kPair ? 2 * static_cast<int>(cluster_nctaid_x()) : static_cast<int>(gridDim.x);
    const int tile_end = kPair ? 2 * ((m_tiles + 1) / 2) : m_tiles;

    // arrive on a barrier that lives in the leader CTA (pair) / in this CTA
    auto arrive_leader = [&](uint64_t* bar) {
        if constexpr (kPair) mbar_arrive_cluster(mapa_shared(smem_u32(bar), 0));
        else mbar_arrive(bar);
    };
    // completion of all MMAs issued so far -> barrier (pair: in both CTAs)
    auto commit = [&](uint64_t* bar) {
        if constexpr (kPair) umma_commit_2sm(bar, 3);
        else umma_commit(bar);
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w1);
        tma_prefetch_desc(&tmap_w2);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kSlots; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(x_local, 1);
        mbar_init(x_ready, 2);
        mbar_init(x_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc1_full[i], 1);
            mbar_init(&acc1_empty[i], kFfActWarps * kNCta);
            mbar_init(&h_full[i], kFfActWarps * kNCta);
            mbar_init(&h_empty[i], 1);
        }
        mbar_init(acc2_full, 1);
        mbar_init(acc2_empty, kFfLnWarps * kNCta);
        fence_mbar_init();
    }
    if (warp == 2) {
        if constexpr (kPair) tmem_alloc_2sm<512>(tmem_base_smem);
        else tmem_alloc<512>(tmem_base_smem);
    }
    tc_fence_before();
    if constexpr (kPair) cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / load completion targets them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        // ------------------------------------------------ TMA producer: x tile, then the weight ring in consumption order
        if (lane == 0) {
            int slot = 0;
            uint32_t phase = 0;
            auto next_slot = [&]() {
                if (++slot == kSlots) { slot = 0; phase ^= 1; }
            };
            // one weight box into this CTA's slot; pair: the bytes are counted on the leader's barrier
            auto load_w = [&](uint8_t* dst, const CUtensorMap* map, int c0, int c1) {
                if constexpr (kPair) tma_load_2d_2sm(dst, map, mapa_shared(smem_u32(&full_bar[slot]), 0), c0, c1, kEvictLast);
                else tma_load_2d(dst, map, &full_bar[slot], c0, c1, kEvictLast);
            };
            int it = 0;
            for (int tile = tile_first; tile < tile_end; tile += tile_stride, ++it) {
                mbar_wait(x_empty, (it & 1) ^ 1);
                mbar_arrive_expect_tx(x_local, kFfXBytes);
                for (int kb = 0; kb < 6; ++kb) tma_load_2d(smem_x + kb * 16384, &tmap_x, x_local, kb * 64, tile * kGemmBlockM, kEvictFirst);
                auto load_w1 = [&](int c) {  // this CTA's W1 rows of chunk c: 3 slots of two k-blocks [kW1Rows x 64]
                    for (int s = 0; s < 3; ++s) {
                        mbar_wait(&empty_bar[slot], phase ^ 1);
                        if (leader) mbar_arrive_expect_tx(&full_bar[slot], kSlotBytes * kNCta);
                        uint8_t* dst = smem_w + slot * kSlotBytes;
                        const int row = c * kFfChunk + static_cast<int>(rank) * kW1Rows;
                        load_w(dst, &tmap_w1, (2 * s) * 64, row);
                        load_w(dst + kSlotBytes / 2, &tmap_w1, (2 * s + 1) * 64, row);
                        next_slot();
                    }
                };
                load_w1(0);
                for (int c = 0; c < n_chunks; ++c) {
                    if (c + 1 < n_chunks) load_w1(c + 1);
                    for (int r = 0; r < 3; ++r) {  // W2 output rows [r*128, +128), columns [c*64, +64): this CTA's 64 (pair) / 128 rows
                        mbar_wait(&empty_bar[slot], phase ^ 1);
                        if (leader) mbar_arrive_expect_tx(&full_bar[slot], kSlotBytes * kNCta);
                        uint8_t* dst = smem_w + slot * kSlotBytes;
                        if constexpr (kPair) {
                            load_w(dst, &tmap_w2, c * kFfChunk, r * 128 + static_cast<int>(rank) * 64);
                        } else {
                            load_w(dst, &tmap_w2, c * kFfChunk, r * 128);
                            load_w(dst + 8192, &tmap_w2, c * kFfChunk, r * 128 + 64);
                        }
                        next_slot();
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------ pair: tell the leader that this CTA's x tile has landed
        if constexpr (kPair) {
            if (lane == 0) {
                int it = 0;
                for (int tile = tile_first; tile < tile_end; tile += tile_stride, ++it) {
                    mbar_wait(x_local, it & 1);
                    arrive_leader(x_ready);
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer (pair: leader CTA only, for both tiles)
        if (lane == 0 && leader) {
            constexpr uint32_t idesc1 = umma_idesc(1 /*bf16*/, 128 * kNCta, kFfChunk);
            constexpr uint32_t idesc2 = umma_idesc(1 /*bf16*/, 128 * kNCta, 128);
            auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
                if constexpr (kPair) umma_f16_2sm(d, a, b, idesc, accumulate);
                else umma_f16(d, a, b, idesc, accumulate);
            };
            int slot = 0;
            uint32_t phase = 0;
            auto next_slot = [&]() {
                if (++slot == kSlots) { slot = 0; phase ^= 1; }
            };
            uint32_t cc = 0;  // running chunk counter (barrier parities continue across tiles)
            int it = 0;
            for (int tile = tile_first; tile < tile_end; tile += tile_stride, ++it) {
                if constexpr (kPair) mbar_wait(x_ready, it & 1);
                else mbar_wait(x_local, it & 1);
                mbar_wait(acc2_empty, (it & 1) ^ 1);
                tc_fence_after();
                auto issue_g1 = [&](int c) {  // acc1[j & 1] = x . W1_c^T
                    const uint32_t j = cc + c;
                    const int ab = j & 1;
                    mbar_wait(&acc1_empty[ab], ((j >> 1) & 1) ^ 1);  // E1 of chunk j-2 has read this buffer
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + kFfAcc1Col + ab * kFfChunk;
                    for (int s = 0; s < 3; ++s) {
                        mbar_wait(&full_bar[slot], phase);
                        tc_fence_after();
#pragma unroll
                        for (int hb = 0; hb < 2; ++hb) {
                            const uint64_t da = umma_desc_k_sw128(smem_u32(smem_x + (2 * s + hb) * 16384));
                            const uint64_t db = umma_desc_k_sw128(smem_u32(smem_w + slot * kSlotBytes + hb * (kSlotBytes / 2)));
#pragma unroll
                            for (int k = 0; k < 4; ++k) mma(tmem_d, da + 2 * k, db + 2 * k, idesc1, (s | hb | k) != 0);
                        }
                        commit(&empty_bar[slot]);
                        next_slot();
                    }
                    commit(&acc1_full[ab]);
                };
                issue_g1(0);
                for (int c = 0; c < n_chunks; ++c) {
                    if (c + 1 < n_chunks) issue_g1(c + 1);
                    // ---- G2(c): acc2 += h_c . W2_c^T
                    const uint32_t j = cc + c;
                    const int hb = j & 1;
                    mbar_wait(&h_full[hb], (j >> 1) & 1);
                    tc_fence_after();
                    const uint64_t da = umma_desc_k_sw128(smem_u32(smem_h + hb * kFfHBufBytes));
                    for (int r = 0; r < 3; ++r) {
                        mbar_wait(&full_bar[slot], phase);
                        tc_fence_after();
                        const uint64_t db = umma_desc_k_sw128(smem_u32(smem_w + slot * kSlotBytes));
#pragma unroll
                        for (int k = 0; k < 4; ++k) mma(tmem_base + r * 128, da + 2 * k, db + 2 * k, idesc2, (c | k) != 0);
                        commit(&empty_bar[slot]);
                        next_slot();
                    }
                    commit(&h_empty[hb]);
                }
                commit(acc2_full);
                cc += n_chunks;
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;  // 0..11
        const int quad = warp & 3;
        const int trow = quad * 32 + lane;
        uint32_t cc = 0;
        int it = 0;
        for (int tile = tile_first; tile < tile_end; tile += tile_stride, ++it) {
            // ------------------------------------------------ activation epilogue (warps 4..11)
            if (ew < kFfActWarps) {
                const int half = ew >> 2;  // columns [half*32, +32) of the 64-wide chunk
                const uint32_t taddr_q = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + kFfAcc1Col + half * 32;
                const uint32_t sw = trow & 7;
                for (int c = 0; c < n_chunks; ++c) {
                    const uint32_t j = cc + c;
                    float4 b[8];  // bias of this chunk, fetched before the wait so its latency is hidden
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        b[q] = p.b1 != nullptr ? __ldg(reinterpret_cast<const float4*>(p.b1 + c * kFfChunk + half * 32) + q)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
                    const int ab = j & 1;
                    mbar_wait(&acc1_full[ab], (j >> 1) & 1);
                    tc_fence_after();
                    uint32_t v[32];
                    tmem_ld_32x32(taddr_q + ab * kFfChunk, v);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_leader(&acc1_empty[ab]);  // this acc1 buffer may be overwritten by G1 of chunk j+2
                    float f[32];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        f[4 * q + 0] = __uint_as_float(v[4 * q + 0]) + b[q].x;
                        f[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + b[q].y;
                        f[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + b[q].z;
                        f[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + b[q].w;
                    }
                    apply_act_tile(f, p.act);
                    const int hb = j & 1;
                    mbar_wait(&h_empty[hb], ((j >> 1) & 1) ^ 1);  // G2 of chunk j-2 has read this buffer
                    const uint32_t rbase = smem_u32(smem_h + hb * kFfHBufBytes) + trow * 128;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        st_shared_v4(rbase + ((static_cast<uint32_t>(half * 4 + q) ^ sw) << 4), pack_bf16(f[8 * q + 0], f[8 * q + 1]),
                                     pack_bf16(f[8 * q + 2], f[8 * q + 3]), pack_bf16(f[8 * q + 4], f[8 * q + 5]),
                                     pack_bf16(f[8 * q + 6], f[8 * q + 7]));
                    }
                    fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
                    __syncwarp();
                    if (lane == 0) arrive_leader(&h_full[hb]);
                }
            }
            cc += n_chunks;
            // ------------------------------------------------ LayerNorm epilogue (warps 4..15)
            const int part = ew >> 2;  // columns [part*128, +128)
            float2* stat = reinterpret_cast<float2*>(smem_h);  // [3][128]: the h buffers are idle once acc2 is complete
            mbar_wait(x_local, it & 1);  // the residual is read from the TMA-written tile: observe its barrier directly
            mbar_wait(acc2_full, it & 1);
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + part * 128;
            const uint32_t swx = trow & 7;
            // pass A: v = acc2 + b2 + residual (resident x tile) ; sum, sum of squares ; v -> TMEM
            float s1 = 0.0f, s2 = 0.0f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int col0 = part * 128 + c * 32;
                uint32_t v[32];
                tmem_ld_32x32(taddr0 + c * 32, v);
                const uint32_t xrow = smem_u32(smem_x + (col0 >> 6) * 16384) + trow * 128;
                const uint32_t ch0 = (col0 & 63) >> 3;
                uint4 r4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) r4[q] = ld_shared_v4(xrow + (((ch0 + q) ^ swx) << 4));
                float4 bb[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    bb[q] = p.b2 != nullptr ? __ldg(reinterpret_cast<const float4*>(p.b2 + col0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t w[4] = {r4[q].x, r4[q].y, r4[q].z, r4[q].w};
                    const float bv[8] = {bb[2 * q].x, bb[2 * q].y, bb[2 * q].z, bb[2 * q].w, bb[2 * q + 1].x, bb[2 * q + 1].y, bb[2 * q + 1].z, bb[2 * q + 1].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
                        const float a0 = __uint_as_float(v[8 * q + 2 * e]) + lo + bv[2 * e];
                        const float a1 = __uint_as_float(v[8 * q + 2 * e + 1]) + hi + bv[2 * e + 1];
                        s1 += a0 + a1;
                        s2 = fmaf(a0, a0, s2);
                        s2 = fmaf(a1, a1, s2);
                        v[8 * q + 2 * e] = __float_as_uint(a0);
                        v[8 * q + 2 * e + 1] = __float_as_uint(a1);
                    }
                }
                tmem_st_32x32(taddr0 + c * 32, v);
            }
            tmem_st_wait();
            stat[part * 128 + trow] = make_float2(s1, s2);
            named_bar_sync(1, kFfLnWarps * 32);
            float t1 = 0.0f, t2 = 0.0f;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float2 t = stat[q * 128 + trow];
                t1 += t.x;
                t2 += t.y;
            }
            const float mean = t1 * (1.0f / kFfH);
            const float var = fmaxf(t2 * (1.0f / kFfH) - mean * mean, 0.0f);
            const float rstd = 1.0f / sqrtf(var + p.eps);
            const float nmr = -mean * rstd;
            // pass B: normalise -> bf16 -> back into the resident tile (own row, own columns)
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int col0 = part * 128 + c * 32;
                uint32_t v[32];
                tmem_ld_32x32(taddr0 + c * 32, v);
                tmem_ld_wait();
                if (c == 3) {  // accumulator consumed
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_leader(acc2_empty);
                }
                float f[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + col0) + q);
                    const float4 bt = __ldg(reinterpret_cast<const float4*>(p.beta + col0) + q);
                    f[4 * q + 0] = fmaf(fmaf(__uint_as_float(v[4 * q + 0]), rstd, nmr), g.x, bt.x);
                    f[4 * q + 1] = fmaf(fmaf(__uint_as_float(v[4 * q + 1]), rstd, nmr), g.y, bt.y);
                    f[4 * q + 2] = fmaf(fmaf(__uint_as_float(v[4 * q + 2]), rstd, nmr), g.z, bt.z);
                    f[4 * q + 3] = fmaf(fmaf(__uint_as_float(v[4 * q + 3]), rstd, nmr), g.w, bt.w);
                }
                const uint32_t xrow = smem_u32(smem_x + (col0 >> 6) * 16384) + trow * 128;
                const uint32_t ch0 = (col0 & 63) >> 3;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    st_shared_v4(xrow + (((ch0 + q) ^ swx) << 4), pack_bf16(f[8 * q + 0], f[8 * q + 1]), pack_bf16(f[8 * q + 2], f[8 * q + 3]),
                                 pack_bf16(f[8 * q + 4], f[8 * q + 5]), pack_bf16(f[8 * q + 6], f[8 * q + 7]));
                }
            }
            fence_proxy_async_smem();
            named_bar_sync(1, kFfLnWarps * 32);  // the whole tile is normalised (and stat[] fully read)
            if (ew == 0 && lane == 0) {
                for (int kb = 0; kb < 6; ++kb) tma_store_2d(&tmap_x, smem_x + kb * 16384, kb * 64, tile * kGemmBlockM);
                bulk_commit();
                bulk_wait_read<0>();   // smem read by the stores: the tile may be replaced
                mbar_arrive(x_empty);
            }
        }
    }

    tc_fence_before();
    if constexpr (kPair) cluster_sync_all();  // the peer may still signal this CTA's barriers / read its smem until it is done too
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (kPair) tmem_dealloc_2sm<512>(tmem_base);
        else tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace kj
