// tcgen05 / TMEM self-attention for short sequences (S <= 128, head_dim 32 or 64): the B200-native form of
// EncoderSelfAttention between the QKV projection and the output projection (reference:
// kjarni-transformers/src/cpu/encoder/encoder_self_attention.rs:213-298; mask :311-325 and utils/masks.rs:7-36;
// softmax activations.rs:223-279).  One work unit = (sequence, head):
//     S = Q K^T          tcgen05.mma 128 x 128 x d, Q/K head slices TMA-loaded (K-major, swizzled), fp32 scores in TMEM
//     P = softmax(S)     one thread per query row reads its row straight out of TMEM (no shuffles), scale + key-padding
//                        mask + exp2 in fp32 registers, un-normalised P written as bf16 into a swizzled smem tile
//     O = P V            tcgen05.mma 128 x d x 128, V consumed in place as an MN-major operand, fp32 O in TMEM
//     ctx = O / rowsum   -> bf16 -> swizzled smem -> TMA store into the merged-head [B,S,H] layout
// Persistent CTAs; two or three warpgroups rotate over as many (smem, TMEM) slots so the tensor core, the TMA engine and the MUFU-bound
// softmax of consecutive units overlap.  Nothing [S,S]-sized ever reaches HBM.
#pragma once
#include <cuda.h>

#include "attention.cuh"

namespace kj {

// warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4.. softmax+epilogue warpgroups (AtcCfg<D>::kThreads in total)
constexpr int kAtcS = 128;        // padded sequence tile

// Milestone stamps of the softmax warpgroups (scripts/attn_trace.py) are compiled in only with
// `make EXTRA=-DKJ_ATTN_TRACE_BUILD=1`: even untaken, their per-unit checks cost ~0.3 us per launch.
#ifndef KJ_ATTN_TRACE_BUILD
#define KJ_ATTN_TRACE_BUILD 0
#endif

template <int D>
struct AtcCfg {
    static constexpr int kRowBytes = D * 2;                       // 64 or 128: one swizzle atom wide
    static constexpr int kTileBytes = kAtcS * kRowBytes;          // Q / K / V head slice
    static constexpr int kInBytes = 3 * kTileBytes;               // one input stage (Q, K, V of a unit)
    static constexpr int kWarpGroups = D == 32 ? 3 : 2;           // softmax warpgroups = (smem P, TMEM S/O) slots in flight
    static constexpr int kThreads = 128 + kWarpGroups * 128;
    static constexpr int kInStages = D == 32 ? 5 : 3;             // TMA runs this many units ahead of the tensor core
    // A separate output staging tile per slot (instead of aliasing the P tile) was measured: no change (18.4 us per launch,
    // scripts/attn_trace.py), so the staging stays aliased and the TMA ring keeps its fifth stage
    static constexpr bool kSepOut = false;
    static constexpr int kTmemO = kWarpGroups * 128;              // TMEM columns: S[w] at w*128, O[w] at kTmemO + w*D
    static constexpr int kPBytes = kAtcS * kAtcS * 2;             // 32 KB: two K-blocks of [128 rows x 128 B]; the output
                                                                  // staging of the same unit aliases it (P is dead by then)
    static constexpr int kStageOutBytes = 32 * kRowBytes;         // per-warp output staging
    static constexpr int kSlotBytes = kPBytes + 1024 + (kSepOut ? 4 * kStageOutBytes : 0);  // + mask codes / flags (+ output staging)
    static constexpr int kSmemBytes = kInStages * kInBytes + kWarpGroups * kSlotBytes + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr uint32_t kSwizzleLayout = D == 32 ? 4u : 2u;  // UMMA layout type: SWIZZLE_64B / SWIZZLE_128B
    static constexpr uint32_t kSbo = 8 * kRowBytes;                // bytes between 8-row groups
};

// K-major operand whose rows are exactly one swizzle atom (64 B or 128 B) wide, or MN-major operand one atom wide:
// in both cases consecutive rows are kRowBytes apart and 8-row groups are SBO apart.
__device__ __forceinline__ uint64_t umma_desc_atom(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 16;  // LBO: only used when the tile is several atoms wide; keep it sane
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout_type) << 61;
    return d;
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* smem_src, int32_t c0, int32_t c1, int32_t c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :
                 : "l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

template <int D>
__global__ void __launch_bounds__(AtcCfg<D>::kThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_ctx, AttnParams p) {
    using Cfg = AtcCfg<D>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int NIN = Cfg::kInStages;
    constexpr int NWG = Cfg::kWarpGroups;
    uint8_t* smem_in = smem;
    uint8_t* smem_slots = smem + NIN * Cfg::kInBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_slots + NWG * Cfg::kSlotBytes);
    uint64_t* in_full = bars;                     // [NIN] TMA -> MMA
    uint64_t* in_empty = bars + NIN;              // [NIN] MMA (PV done) -> TMA
    uint64_t* s_full = bars + 2 * NIN;            // [NWG] MMA (QK done) -> warpgroup
    uint64_t* p_full = bars + 2 * NIN + NWG;      // [NWG] warpgroup (P written, S consumed) -> MMA
    uint64_t* o_full = bars + 2 * NIN + 2 * NWG;  // [NWG] MMA (PV done) -> warpgroup
    uint64_t* o_empty = bars + 2 * NIN + 3 * NWG; // [NWG] warpgroup (O consumed) -> MMA
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 2 * NIN + 4 * NWG);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_units = p.B * p.heads;
    // Work-unit order: with enough sequences every CTA walks whole sequences (all heads back to back: the mask codes are
    // built once per sequence and consecutive Q/K/V slices are adjacent in memory); small batches fall back to striding units.
    const bool seq_major = p.B >= static_cast<int>(gridDim.x);
    auto get_unit = [&](int i, int& b, int& h) -> bool {
        if (seq_major) {
            b = blockIdx.x + (i / p.heads) * gridDim.x;
            h = i % p.heads;
            return b < p.B;
        }
        const int u = blockIdx.x + i * gridDim.x;
        b = u / p.heads;
        h = u % p.heads;
        return u < n_units;
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_qkv);
        tma_prefetch_desc(&tmap_ctx);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NIN; ++i) {
            mbar_init(&in_full[i], 1);
            mbar_init(&in_empty[i], 1);
        }
        for (int i = 0; i < NWG; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);   // one arrive per warp of the warpgroup
            mbar_init(&o_full[i], 1);
            mbar_init(&o_empty[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();               // everything above overlapped the previous kernel's tail; its outputs are visible from here on
    pdl_launch_dependents();  // the next kernel may begin its own prologue as soon as this CTA's resources are released

    auto slot_base = [&](int slot) { return smem_slots + slot * Cfg::kSlotBytes; };  // P tile (+ aliased output staging), codes
    auto in_base = [&](int stage) { return smem_in + stage * Cfg::kInBytes; };      // Q | K | V

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int b, h;
            for (int i = 0; get_unit(i, b, h); ++i) {
                const int stage = i % NIN;
                const uint32_t par = (i / NIN) & 1;
                mbar_wait(&in_empty[stage], par ^ 1);
                uint8_t* sb = in_base(stage);
                mbar_arrive_expect_tx(&in_full[stage], Cfg::kInBytes);
                tma_load_3d(sb, &tmap_qkv, &in_full[stage], h * D, 0, b);                                   // Q
                tma_load_3d(sb + Cfg::kTileBytes, &tmap_qkv, &in_full[stage], p.H + h * D, 0, b);           // K
                tma_load_3d(sb + 2 * Cfg::kTileBytes, &tmap_qkv, &in_full[stage], 2 * p.H + h * D, 0, b);   // V
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc_qk = umma_idesc(1, 128, kAtcS);
            constexpr uint32_t idesc_pv = umma_idesc(1, 128, D) | (1u << 16);  // B (= V) is MN-major
            int n_mine = 0;
            {
                int b, h;
                while (get_unit(n_mine, b, h)) ++n_mine;
            }
            // Greedy issue: QK of unit q as soon as its inputs landed and its S slot is free (the softmax of unit q - NWG has
            // published its P, i.e. PV of that unit was already issued), PV of unit v as soon as its P is published.
            int qk_next = 0, pv_next = 0;
            while (pv_next < n_mine) {
                if (qk_next < n_mine && qk_next < pv_next + NWG && mbar_try_wait(&in_full[qk_next % NIN], (qk_next / NIN) & 1)) {
                    const int slot = qk_next % NWG, stage = qk_next % NIN;
                    tc_fence_after();
                    uint8_t* sb = in_base(stage);
                    const uint64_t dq = umma_desc_atom(smem_u32(sb), Cfg::kSbo, Cfg::kSwizzleLayout);
                    const uint64_t dk = umma_desc_atom(smem_u32(sb + Cfg::kTileBytes), Cfg::kSbo, Cfg::kSwizzleLayout);
#pragma unroll
                    for (int k = 0; k < D / 16; ++k) umma_f16(tmem_base + slot * 128, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
                    umma_commit(&s_full[slot]);
                    ++qk_next;
                }
                if (pv_next < qk_next) {
                    const int j = pv_next, slot = j % NWG;
                    const uint32_t par = (j / NWG) & 1;
                    if (mbar_try_wait(&p_full[slot], par) && mbar_try_wait(&o_empty[slot], par ^ 1)) {
                        tc_fence_after();
                        const int stage = j % NIN;
                        const uint64_t dv = umma_desc_atom(smem_u32(in_base(stage) + 2 * Cfg::kTileBytes), Cfg::kSbo, Cfg::kSwizzleLayout);
                        uint8_t* sp = slot_base(slot);
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb) {
                            const uint64_t dp = umma_desc_k_sw128(smem_u32(sp + kb * (kAtcS * 128)));
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                // V advances 16 keys = 16 rows per k-step
                                umma_f16(tmem_base + Cfg::kTmemO + slot * D, dp + 2 * k, dv + ((kb * 4 + k) * 16 * Cfg::kRowBytes >> 4),
                                         idesc_pv, (kb | k) != 0);
                            }
                        }
                        umma_commit(&o_full[slot]);
                        umma_commit(&in_empty[stage]);
                        ++pv_next;
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------- softmax + epilogue warpgroups
        const int wg = (warp - 4) >> 2;  // warpgroup index = slot
        const int quad = warp & 3;
        const int row = quad * 32 + lane;  // query row = TMEM lane
        const int slot = wg;
        uint8_t* sp = slot_base(slot);
        uint8_t* sout = Cfg::kSepOut ? sp + Cfg::kPBytes + 1024 + quad * Cfg::kStageOutBytes
                                     : sp + quad * Cfg::kStageOutBytes;        // D = 64: aliases the P tile (dead once o_full fires)
        float* codes = reinterpret_cast<float*>(sp + Cfg::kPBytes);            // [128] + flags
        int* wvalid = reinterpret_cast<int*>(codes + kAtcS);                   // [4] any key kept, per 32-key chunk
        int* wfull = wvalid + 4;                                               // [4] all 32 keys of the chunk kept
        constexpr float kMaskedLog2 = -1.0e9f * 1.4426950408889634f;
        const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + slot * 128;
        const uint32_t t_o = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + Cfg::kTmemO + slot * D;
        const int bar_id = 1 + wg;

        int n = 0, last_b = -1, b = 0, h = 0;
        bool poison = false;
        int cfull = 0;
        for (int i = 0; get_unit(i, b, h); ++i) {
            if (i % NWG != wg) continue;
            const uint32_t par = n & 1;
            ++n;
            const bool tr = KJ_ATTN_TRACE_BUILD && p.trace != nullptr && quad == 0 && lane == 0 && n >= 2 && n <= 4;
            unsigned long long* tp = tr ? p.trace + blockIdx.x * 64 + wg * 20 + (n - 2) * 6 : nullptr;
            auto stamp = [&](int k) { if (tr) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); tp[k] = t; } };
            stamp(0);
            // the previous unit's TMA stores have left the staging area (it aliases the P tile every thread is about to
            // write) and its readers of codes[] are done
            if constexpr (!Cfg::kSepOut) {
                if (lane == 0) bulk_wait_read<0>();
                named_bar_sync(bar_id, 128);
            }
            if (b != last_b) {
                if constexpr (Cfg::kSepOut) named_bar_sync(bar_id, 128);  // every reader of the previous sequence's codes[] is done
                // mask codes of this sequence: 0 = keep, else the value the score is replaced by
                last_b = b;
                const int j = quad * 32 + lane;
                float code;
                bool keep = false;
                if (j >= p.S) code = -INFINITY;
                else {
                    keep = (p.mask == nullptr) || (p.mask[static_cast<size_t>(b) * p.S + j] != 0.0f);
                    code = keep ? 0.0f : kMaskedLog2;
                }
                codes[j] = code;
                const uint32_t any = __ballot_sync(0xffffffffu, keep);
                if (lane == 0) {
                    wvalid[quad] = any != 0;
                    wfull[quad] = any == 0xffffffffu;
                }
                named_bar_sync(bar_id, 128);
                poison = p.nan_if_all_masked && !(wvalid[0] | wvalid[1] | wvalid[2] | wvalid[3]);
                cfull = (wfull[0] ? 1 : 0) | (wfull[1] ? 2 : 0) | (wfull[2] ? 4 : 0) | (wfull[3] ? 8 : 0);
            }

            mbar_wait(&s_full[slot], par);
            stamp(1);
            tc_fence_after();
            // pass 1: row max of the scaled + masked scores (next chunk's TMEM load overlaps this chunk's math)
            float mx = -INFINITY;
            uint32_t va[32], vb[32];
            auto max_chunk = [&](const uint32_t (&v)[32], int c) {
                if ((cfull >> c) & 1) {  // warp-uniform fast path: no padding in this 32-key chunk
                    float m = __uint_as_float(v[0]);
#pragma unroll
                    for (int j = 1; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
                    mx = fmaxf(mx, m * p.scale_log2e);  // scale > 0 commutes with max
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float cd = codes[c * 32 + j];
                        const float s = cd == 0.0f ? __uint_as_float(v[j]) * p.scale_log2e : cd;
                        mx = fmaxf(mx, s);
                    }
                }
            };
            tmem_ld_32x32(t_s, va);
            tmem_ld_wait();
            tmem_ld_32x32(t_s + 32, vb);
            max_chunk(va, 0);
            tmem_ld_wait();
            tmem_ld_32x32(t_s + 64, va);
            max_chunk(vb, 1);
            tmem_ld_wait();
            tmem_ld_32x32(t_s + 96, vb);
            max_chunk(va, 2);
            tmem_ld_wait();
            tmem_ld_32x32(t_s, va);  // first chunk of pass 2
            max_chunk(vb, 3);
            stamp(2);
            // pass 2: p = exp2(s - max), row sum, bf16 P into the swizzled K-major smem tile
            float sum = 0.0f;
            const uint32_t prow = smem_u32(sp) + row * 128;
            auto exp_chunk = [&](const uint32_t (&v)[32], int c) {
                float f[32];
                if ((cfull >> c) & 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        f[j] = ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2e, -mx));
                        sum += f[j];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float cd = codes[c * 32 + j];
                        const float s = cd == 0.0f ? __uint_as_float(v[j]) * p.scale_log2e : cd;
                        f[j] = ex2_approx(s - mx);
                        sum += f[j];
                    }
                }
                const uint32_t blk = prow + (c >> 1) * (kAtcS * 128);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + j) ^ static_cast<uint32_t>(row & 7);
                    st_shared_v4(blk + (chunk << 4), pack_bf16(f[8 * j + 0], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                 pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
                }
            };
            tmem_ld_wait();
            tmem_ld_32x32(t_s + 32, vb);
            exp_chunk(va, 0);
            tmem_ld_wait();
            tmem_ld_32x32(t_s + 64, va);
            exp_chunk(vb, 1);
            tmem_ld_wait();
            tmem_ld_32x32(t_s + 96, vb);
            exp_chunk(va, 2);
            tmem_ld_wait();
            exp_chunk(vb, 3);
            fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core's async-proxy reads
            tc_fence_before();         // S fully read before the MMA warp may overwrite it
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[slot]);
            stamp(3);

            // epilogue: O / rowsum -> bf16 -> swizzled staging -> TMA store
            float inv = 1.0f / sum;
            if (poison) inv = __int_as_float(0x7fc00000);
            mbar_wait(&o_full[slot], par);
            stamp(4);
            tc_fence_after();
            uint32_t o[D];
            if constexpr (D == 32) {
                tmem_ld_32x32(t_o, reinterpret_cast<uint32_t(&)[32]>(o));
            } else {
                tmem_ld_32x32(t_o, reinterpret_cast<uint32_t(&)[32]>(o[0]));
                tmem_ld_32x32(t_o + 32, reinterpret_cast<uint32_t(&)[32]>(o[32]));
            }
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_empty[slot]);
            if constexpr (Cfg::kSepOut) {  // this warp's previous ctx store (one unit ago) has read its staging tile
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
            }
            const uint32_t obase = smem_u32(sout) + lane * Cfg::kRowBytes;
            const uint32_t sw = D == 32 ? ((lane >> 1) & 3) : (lane & 7);
#pragma unroll
            for (int j = 0; j < D / 8; ++j) {
                st_shared_v4(obase + ((static_cast<uint32_t>(j) ^ sw) << 4),
                             pack_bf16(__uint_as_float(o[8 * j + 0]) * inv, __uint_as_float(o[8 * j + 1]) * inv),
                             pack_bf16(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv),
                             pack_bf16(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv),
                             pack_bf16(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && quad * 32 < p.S) {
                tma_store_3d(&tmap_ctx, sout, h * D, quad * 32, b);  // rows >= S are clipped by the tensor map
                bulk_commit();
            }
            stamp(5);
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
