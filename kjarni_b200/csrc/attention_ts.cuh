// tcgen05 / TMEM self-attention for S <= 512, head_dim 32 or 64 -- the default attention kernel of the encoder (round 2).
// B200-native form of EncoderSelfAttention between the QKV projection and the output projection (reference:
// kjarni-transformers/src/cpu/encoder/encoder_self_attention.rs:213-298; key-padding mask :311-325 and utils/masks.rs:7-36;
// softmax activations.rs:223-279).
//
// What changed against attention_tc.cuh (round 1, S <= 128 only):
//   * P never touches shared memory.  The softmax threads write the un-normalised bf16 P straight back into the TMEM columns the
//     scores came from (tcgen05.st, P aliases the first half of its S block) and the P.V product takes its A operand FROM TMEM
//     (tcgen05.mma [d], [a_tmem], b_desc): no 32 KB P tile, no generic->async proxy fence, no P-tile/out-staging alias, and the
//     shared memory that frees holds six input stages.  On an idle SM a TS-mode MMA of N = 32 retires every 16-20 clk against 40 for
//     the SS form (scripts/ubench/mma_issue.cu).
//   * O accumulates in the dead upper half of the slot's first S block ([64, 64+D)), so a (query block) slot is exactly
//     ceil(S/128) x 128 TMEM columns: four units in flight at S <= 128 (round 1: three), two at S <= 256, one at S <= 512.
//   * Sequences longer than 128 tokens: a unit is still (sequence, head); its Q, K, V slices are loaded once and stay resident
//     while its ceil(S/128) query blocks run.  The key blocks of one query block are split over warpgroups (thread = query row x
//     128 keys), which exchange the row max / row sum through shared memory -- exact two-pass softmax, no online rescaling.
//   * The MMA issuer is a whole warp running warp-uniform code: descriptors and TMEM addresses stay in uniform registers and ONE
//     elected lane issues each group of tcgen05 instructions back to back.  Under `if (lane == 0)` the compiler wraps every
//     UTCHMMA in an R2UR / ELECT / BRA.U.ANY waterfall (~10 dependent instructions), and on a scheduler shared with four busy
//     softmax warps every dependent instruction waits for an issue slot again: issuing the eight P.V steps of a unit cost 0.6 us
//     (profiles/r02_attention_analysis.md).  It polls with mbarrier.test_wait (try_wait may suspend the thread on ONE barrier
//     while another one completes).
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-19 four softmax/epilogue warpgroups.
#pragma once
#include <cuda.h>

#include "attention_tc.cuh"

namespace kj {

// 2^x for x <= 0 on the FMA pipe, two values per call (packed fp32x2): x = n + f with n = round(x) taken from the mantissa of
// x + 1.5 * 2^23, 2^f by a cubic on [-0.5, 0.5] (max relative error 1.9e-4, a tenth of the bf16 rounding of P), 2^n added to the
// exponent field.  x is clamped at -125 (2^-125 is 0 for every purpose here).  Pairs of every 8 that use it: KJ_ATTN_POLY_PAIRS.
#ifndef KJ_ATTN_POLY_PAIRS
#define KJ_ATTN_POLY_PAIRS 0
#endif
constexpr int kPolyPairs = KJ_ATTN_POLY_PAIRS;
#ifndef KJ_ATTN_TWO_ISSUERS
#define KJ_ATTN_TWO_ISSUERS 1
#endif
#ifndef KJ_ATTN_PACKED_SOFTMAX
#define KJ_ATTN_PACKED_SOFTMAX 1
#endif
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& r0, float& r1) {
    const uint64_t X = f2_pack(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
    const uint64_t T = f2_add(X, f2_pack(12582912.0f, 12582912.0f));
    const uint64_t Nf = f2_add(T, f2_pack(-12582912.0f, -12582912.0f));
    const uint64_t F = f2_fma(Nf, f2_pack(-1.0f, -1.0f), X);
    uint64_t P = f2_fma(F, f2_pack(0.0558755f, 0.0558755f), f2_pack(0.24229444f, 0.24229444f));
    P = f2_fma(P, F, f2_pack(0.69312727f, 0.69312727f));
    P = f2_fma(P, F, f2_pack(0.99994824f, 0.99994824f));
    float p0, p1, t0, t1;
    f2_unpack(P, p0, p1);
    f2_unpack(T, t0, t1);
    r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
    r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}


#ifndef KJ_ATS_MAX_STAGES
#define KJ_ATS_MAX_STAGES 6
#endif
template <int D, int NKB>
struct AtsCfg {
    static_assert(NKB == 1 || NKB == 2 || NKB == 4, "key blocks per slot");
    static constexpr int kSlots = 4 / NKB;                      // query blocks in flight (TMEM: kSlots x NKB x 128 = 512 columns)
    static constexpr int kThreads = 128 + 4 * 128;              // 640
    static constexpr int kRowBytes = D * 2;                     // one swizzle atom wide (64 B / 128 B)
    static constexpr int kTileBytes = 128 * kRowBytes;          // one 128-row block of Q, K or V
    static constexpr int kStageBytes = 3 * NKB * kTileBytes;    // Q | K | V of one (sequence, head)
    static constexpr int kOutBufs = NKB == 1 ? (D == 32 ? 2 : 1) : 0;  // per-warp output staging (S <= 128: TMA store)
    static constexpr int kOutWarpBytes = 32 * kRowBytes;
    static constexpr int kOutBytes = 16 * kOutBufs * kOutWarpBytes;
    static constexpr int kMiscBytes = 3 * 4 * 128 * 4 + 256 + 512;  // codes | xmax | xsum, flags, barriers
    static constexpr int kBudget = 232448 - 1024 - kOutBytes - kMiscBytes;
    static constexpr int kMaxStages = KJ_ATS_MAX_STAGES;
    static constexpr int kInStages = kBudget / kStageBytes > kMaxStages ? kMaxStages : kBudget / kStageBytes;
    static_assert(kInStages >= 1, "shared memory budget");
    static constexpr int kSmemBytes = kInStages * kStageBytes + kOutBytes + kMiscBytes + 1024;
    static constexpr uint32_t kSwizzleLayout = D == 32 ? 4u : 2u;  // UMMA layout type: SWIZZLE_64B / SWIZZLE_128B
    static constexpr uint32_t kSbo = 8 * kRowBytes;
    static constexpr int kOutCols = D / NKB;                    // S > 128: O columns each warpgroup of a slot stores
    static constexpr int kOCol = 64;                            // O accumulator columns inside the slot (dead half of S block 0)
};

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

template <int D, int NKB>
__global__ void __launch_bounds__(AtsCfg<D, NKB>::kThreads, 1)
attention_ts_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_ctx, AttnParams p) {
    using Cfg = AtsCfg<D, NKB>;
    constexpr int NIN = Cfg::kInStages;
    constexpr int NSL = Cfg::kSlots;
    constexpr bool kTwoIssuers = KJ_ATTN_TWO_ISSUERS != 0 && NKB == 1;  // S <= 128 only: measured slower at S = 256 (32.4 vs 30.9 us)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_in = smem;
    uint8_t* smem_out = smem + NIN * Cfg::kStageBytes;
    float* s_codes = reinterpret_cast<float*>(smem_out + Cfg::kOutBytes);  // [4 warpgroups][128 keys]
    float* s_xmax = s_codes + 4 * 128;                                     // [4][128 rows]
    float* s_xsum = s_xmax + 4 * 128;                                      // [4][128 rows]
    int* s_any = reinterpret_cast<int*>(s_xsum + 4 * 128);                 // [4][4] any key of the 32-key chunk kept
    int* s_all = s_any + 16;                                               // [4][4] every key of the chunk kept
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_all + 48);
    uint64_t* full_qk = bars;                  // [NIN] TMA -> MMA (Q and K blocks landed)
    uint64_t* full_v = bars + NIN;             // [NIN] TMA -> MMA (V blocks landed)
    uint64_t* empty_qk = bars + 2 * NIN;       // [NIN] MMA (last Q.K^T of the unit retired) -> TMA
    uint64_t* empty_v = bars + 3 * NIN;        // [NIN] MMA (last P.V of the unit retired) -> TMA
    uint64_t* s_full = bars + 4 * NIN;         // [4]   MMA (scores of key block j of the slot) -> warpgroup
    uint64_t* p_full = s_full + 4;             // [4]   warpgroup (P written, S consumed) -> MMA
    uint64_t* o_full = p_full + 4;             // [NSL] MMA (P.V done) -> warpgroups of the slot
    uint64_t* o_empty = o_full + NSL;          // [NSL] warpgroups (O consumed) -> MMA
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(o_empty + NSL);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto gstamp = [&](int slot_idx) {  // milestone stamps, compiled in with -DKJ_ATTN_TRACE_BUILD=1 (scripts/attn_trace.py)
        if (KJ_ATTN_TRACE_BUILD && p.trace != nullptr) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
            p.trace[blockIdx.x * 64 + slot_idx] = t;
        }
    };
    if (threadIdx.x == 0) gstamp(60);
    const int n_units = p.B * p.heads;
    const int nkb = (p.S + 127) >> 7;  // live key blocks = query blocks of a unit (<= NKB)
    // unit order: with enough sequences every CTA walks whole sequences (mask codes built once per sequence, adjacent slices)
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
            mbar_init(&full_qk[i], 1);
            mbar_init(&full_v[i], 1);
            mbar_init(&empty_qk[i], 1);
            mbar_init(&empty_v[i], 1);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);  // one arrive per warp of the warpgroup
        }
        for (int i = 0; i < NSL; ++i) {
            mbar_init(&o_full[i], 1);
            mbar_init(&o_empty[i], 4 * NKB);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x == 0) gstamp(61);

    auto stage_q = [&](int stage, int blk) { return smem_in + stage * Cfg::kStageBytes + blk * Cfg::kTileBytes; };
    auto stage_k = [&](int stage, int blk) { return smem_in + stage * Cfg::kStageBytes + (NKB + blk) * Cfg::kTileBytes; };
    auto stage_v = [&](int stage, int blk) { return smem_in + stage * Cfg::kStageBytes + (2 * NKB + blk) * Cfg::kTileBytes; };

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int b, h;
            for (int i = 0; get_unit(i, b, h); ++i) {
                const int stage = i % NIN;
                const uint32_t par = (i / NIN) & 1;
                mbar_wait(&empty_qk[stage], par ^ 1);
                mbar_arrive_expect_tx(&full_qk[stage], 2 * nkb * Cfg::kTileBytes);
                for (int k = 0; k < nkb; ++k) {
                    tma_load_3d(stage_q(stage, k), &tmap_qkv, &full_qk[stage], h * D, k * 128, b);
                    tma_load_3d(stage_k(stage, k), &tmap_qkv, &full_qk[stage], p.H + h * D, k * 128, b);
                }
                mbar_wait(&empty_v[stage], par ^ 1);
                mbar_arrive_expect_tx(&full_v[stage], nkb * Cfg::kTileBytes);
                for (int k = 0; k < nkb; ++k) tma_load_3d(stage_v(stage, k), &tmap_qkv, &full_v[stage], 2 * p.H + h * D, k * 128, b);
            }
        }
    } else if (warp == 1 || (kTwoIssuers && warp == 2)) {
        // -------------------------------------------------------------- MMA issuers (whole warps, warp-uniform code)
        // kTwoIssuers (S <= 128): warp 1 issues every Q.K^T, warp 2 (idle after the TMEM allocation) every P.V, each in unit order with
        // blocking waits -- a P.V no longer queues behind the polling and the Q.K^T issue of other slots (the softmax warps waited
        // 0.9 us per unit for O, profiles/r02_attention_analysis.md).  Otherwise warp 1 issues both from one greedy polling loop.
        constexpr uint32_t idesc_qk = umma_idesc(1, 128, 128);
        constexpr uint32_t idesc_pv = umma_idesc(1, 128, D) | (1u << 16);  // B (= V) is MN-major; A (= P) comes from TMEM
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t smem_in_u = __shfl_sync(0xffffffffu, smem_u32(smem_in), 0);
        int n_mine = 0;
        {
            int b, h;
            while (get_unit(n_mine, b, h)) ++n_mine;
        }
        if ((p.dbg & 1) && warp == 2) n_mine = 0;
        if ((p.dbg & 1) && warp == 1) {  // probe: consume the input stages without any tensor or softmax work
            for (int i = 0; i < n_mine; ++i) {
                mbar_wait(&full_qk[i % NIN], (i / NIN) & 1);
                mbar_wait(&full_v[i % NIN], (i / NIN) & 1);
                if (lane == 0) {
                    mbar_arrive(&empty_qk[i % NIN]);
                    mbar_arrive(&empty_v[i % NIN]);
                }
            }
            n_mine = 0;
        }
        const int total = n_mine * nkb;  // query blocks of this CTA, in issue order g = unit * nkb + query block
        // S[j] = Q[qb] K[j]^T for every live key block of query block g
        auto issue_qk = [&](int g) {
            const int iu = g / nkb, qb = g - iu * nkb;
            const int stage = iu % NIN, slot = g % NSL;
            tc_fence_after();
            const uint32_t sbase = smem_in_u + stage * Cfg::kStageBytes;
            const uint64_t dq = umma_desc_atom(sbase + qb * Cfg::kTileBytes, Cfg::kSbo, Cfg::kSwizzleLayout);
            const uint64_t dk0 = umma_desc_atom(sbase + NKB * Cfg::kTileBytes, Cfg::kSbo, Cfg::kSwizzleLayout);
            const uint32_t t_sl = tmem_u + slot * NKB * 128;
            if (elect_one()) {
                if (g < 8) gstamp(24 + g);
#pragma unroll
                for (int jj = 0; jj < NKB; ++jj) {
                    if (jj < nkb) {
#pragma unroll
                        for (int k = 0; k < D / 16; ++k)
                            umma_f16(t_sl + jj * 128, dq + 2 * k, dk0 + ((jj * Cfg::kTileBytes) >> 4) + 2 * k, idesc_qk, k != 0);
                    }
                    umma_commit(&s_full[slot * NKB + jj]);  // dead key blocks (jj >= nkb): arrives at once
                }
                if (qb == nkb - 1) umma_commit(&empty_qk[stage]);
            }
            __syncwarp();
        };
        // O = P V over every live key block of the slot (P read from TMEM, V consumed in place as an MN-major operand)
        auto issue_pv = [&](int g) {
            const int iu = g / nkb, qb = g - iu * nkb;
            const int stage = iu % NIN, slot = g % NSL;
            tc_fence_after();
            const uint32_t t_sl = tmem_u + slot * NKB * 128;
            const uint64_t dv0 = umma_desc_atom(smem_in_u + stage * Cfg::kStageBytes + 2 * NKB * Cfg::kTileBytes, Cfg::kSbo, Cfg::kSwizzleLayout);
            if (elect_one()) {
                if (g < 8) gstamp(32 + g);
#pragma unroll
                for (int jj = 0; jj < NKB; ++jj) {
                    if (jj < nkb) {
#pragma unroll
                        for (int k = 0; k < 8; ++k)  // 16 keys per step: 8 packed P columns, 16 V rows
                            umma_f16_ts(t_sl + Cfg::kOCol, t_sl + jj * 128 + 8 * k, dv0 + ((jj * Cfg::kTileBytes + k * 16 * Cfg::kRowBytes) >> 4), idesc_pv, (jj | k) != 0);
                    }
                }
                umma_commit(&o_full[slot]);
                if (qb == nkb - 1) umma_commit(&empty_v[stage]);
            }
            __syncwarp();
        };
        // Greedy issue: Q.K^T of query block q as soon as its inputs landed and its slot is free (O of the previous occupant read),
        // P.V of query block v as soon as every warpgroup of its slot has published its P block.
        int qk_next = 0, pv_next = 0;
        if (kTwoIssuers) {
            if (warp == 1) {
                for (int g = 0; g < total; ++g) {
                    const int iu = g / nkb;
                    const int stage = iu % NIN, slot = g % NSL, use = g / NSL;
                    mbar_wait(&full_qk[stage], (iu / NIN) & 1);
                    if (use > 0) mbar_wait(&o_empty[slot], (use - 1) & 1);
                    issue_qk(g);
                }
            } else {
                for (int g = 0; g < total; ++g) {
                    const int iu = g / nkb;
                    const int stage = iu % NIN, slot = g % NSL;
                    const uint32_t par = (g / NSL) & 1;
                    mbar_wait(&full_v[stage], (iu / NIN) & 1);
#pragma unroll
                    for (int j = 0; j < NKB; ++j) mbar_wait(&p_full[slot * NKB + j], par);
                    issue_pv(g);
                }
            }
            pv_next = total;
        }
        while (pv_next < total) {
            if (qk_next < total && qk_next < pv_next + NSL) {
                const int iu = qk_next / nkb;
                const int stage = iu % NIN, slot = qk_next % NSL, use = qk_next / NSL;
                if (mbar_test_wait(&full_qk[stage], (iu / NIN) & 1) && (use == 0 || mbar_test_wait(&o_empty[slot], (use - 1) & 1))) issue_qk(qk_next++);
            }
            if (pv_next < qk_next) {
                const int iu = pv_next / nkb;
                const int stage = iu % NIN, slot = pv_next % NSL;
                const uint32_t par = (pv_next / NSL) & 1;
                bool ready = mbar_test_wait(&full_v[stage], (iu / NIN) & 1);
#pragma unroll
                for (int j = 0; j < NKB; ++j) ready = ready && mbar_test_wait(&p_full[slot * NKB + j], par);
                if (ready) issue_pv(pv_next++);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------- softmax + epilogue warpgroups
        const int w = (warp - 4) >> 2;     // warpgroup 0..3
        const int quad = warp & 3;         // TMEM lane quadrant of this warp
        const int r = quad * 32 + lane;    // query row inside the block = TMEM lane
        const int slot = w / NKB, j = w % NKB;  // this warpgroup: key block j of the query blocks that go through `slot`
        const bool live = j < nkb;
        const uint32_t t_slot = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + slot * NKB * 128;
        const uint32_t t_s = t_slot + j * 128;
        float* codes = s_codes + w * 128;
        const int bar_id = 1 + slot;
        constexpr int kGroupThreads = NKB * 128;
        constexpr float kMaskedLog2 = -1.0e9f * 1.4426950408889634f;
        uint8_t* sout = smem_out + ((warp - 4) * (Cfg::kOutBufs > 0 ? Cfg::kOutBufs : 1)) * Cfg::kOutWarpBytes;

        int last_b = -1, b = 0, h = 0, n = 0;
        bool poison = false;
        int cfull = 0;
        for (int g = slot;; g += NSL, ++n) {
            const int iu = g / nkb, qb = g - iu * nkb;
            if (!get_unit(iu, b, h) || (p.dbg & 1)) break;
            const uint32_t par = n & 1;
            const bool tr = KJ_ATTN_TRACE_BUILD && quad == 0 && lane == 0 && n == (p.dbg >> 4);
            auto stamp = [&](int k) { if (tr) gstamp(w * 6 + k); };
            stamp(0);
            if (b != last_b) {
                // mask codes of key block j of this sequence: 0 = keep, else the value the score is replaced by.  Every reader of
                // the previous codes / flags has passed its p_full arrive, which precedes the o_full this thread has waited for.
                last_b = b;
                const int key = j * 128 + r;
                float code = -INFINITY;
                bool keep = false;
                if (key < p.S) {
                    keep = (p.mask == nullptr) || (p.mask[static_cast<size_t>(b) * p.S + key] != 0.0f);
                    code = keep ? 0.0f : kMaskedLog2;
                }
                codes[r] = code;
                const uint32_t any = __ballot_sync(0xffffffffu, keep);
                if (lane == 0) {
                    s_any[w * 4 + quad] = any != 0;
                    s_all[w * 4 + quad] = any == 0xffffffffu;
                }
                named_bar_sync(bar_id, kGroupThreads);
                int any_seq = 0;
#pragma unroll
                for (int i = 0; i < 4 * NKB; ++i) any_seq |= s_any[slot * NKB * 4 + i];
                poison = p.nan_if_all_masked && !any_seq;
                cfull = (s_all[w * 4] ? 1 : 0) | (s_all[w * 4 + 1] ? 2 : 0) | (s_all[w * 4 + 2] ? 4 : 0) | (s_all[w * 4 + 3] ? 8 : 0);
            }

            mbar_wait(&s_full[w], par);
            stamp(1);
            tc_fence_after();
            // pass 1: row max of the scaled + masked scores over this warpgroup's 128 keys
            float mx = -INFINITY;
            if (live) {
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_s + c * 32, v);
                    tmem_ld_wait();
                    if ((cfull >> c) & 1) {  // warp-uniform fast path: no padding in this 32-key chunk
#if KJ_ATTN_PACKED_SOFTMAX
                        // three-input maxima (FMNMX3): 16 instead of 31 instructions on a half-rate pipe, four independent chains
                        float m4[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            m4[q] = fmax3(__uint_as_float(v[8 * q]), __uint_as_float(v[8 * q + 1]), __uint_as_float(v[8 * q + 2]));
                            m4[q] = fmax3(m4[q], __uint_as_float(v[8 * q + 3]), __uint_as_float(v[8 * q + 4]));
                            m4[q] = fmax3(m4[q], __uint_as_float(v[8 * q + 5]), __uint_as_float(v[8 * q + 6]));
                        }
                        float m = fmax3(fmax3(m4[0], m4[1], __uint_as_float(v[7])), fmax3(m4[2], m4[3], __uint_as_float(v[15])),
                                        fmaxf(__uint_as_float(v[23]), __uint_as_float(v[31])));
#else
                        float m = __uint_as_float(v[0]);
#pragma unroll
                        for (int i = 1; i < 32; ++i) m = fmaxf(m, __uint_as_float(v[i]));
#endif
                        mx = fmaxf(mx, m * p.scale_log2e);  // scale > 0 commutes with max
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float cd = codes[c * 32 + i];
                            mx = fmaxf(mx, cd == 0.0f ? __uint_as_float(v[i]) * p.scale_log2e : cd);
                        }
                    }
                }
            }
            if constexpr (NKB > 1) {  // row max over the key blocks of the slot
                s_xmax[w * 128 + r] = mx;
                named_bar_sync(bar_id, kGroupThreads);
#pragma unroll
                for (int i = 0; i < NKB; ++i) mx = fmaxf(mx, s_xmax[(slot * NKB + i) * 128 + r]);
            }
            stamp(2);
            // pass 2: p = exp2(s - max) -> row sum, bf16 P back into TMEM (two keys per 32-bit column, over the first half of S)
            float sum = 0.0f;
            if (live) {
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_s + c * 32, v);
                    tmem_ld_wait();
                    uint32_t pk[16];
                    if (KJ_ATTN_PACKED_SOFTMAX && NKB == 1 && kPolyPairs == 0 && ((cfull >> c) & 1)) {
                        // scale + shift and the row sum on the packed fp32x2 pipe: per key 1 MUFU.EX2 + half an FFMA2, FADD2 and
                        // F2FP each (2.5 issue slots instead of 3.5 -- the pass is bound by issue slots, not by the MUFU pipe)
                        const uint64_t sc2 = f2_pack(p.scale_log2e, p.scale_log2e), nm2 = f2_pack(-mx, -mx);
                        uint64_t acc2[2] = {f2_pack(0.0f, 0.0f), f2_pack(0.0f, 0.0f)};
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float x0, x1;
                            f2_unpack(f2_fma(f2_pack(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2), x0, x1);
                            const float f0 = ex2_approx(x0), f1 = ex2_approx(x1);
                            acc2[i & 1] = f2_add(acc2[i & 1], f2_pack(f0, f1));
                            pk[i] = pack_bf16(f0, f1);
                        }
                        float a0, a1;
                        f2_unpack(f2_add(acc2[0], acc2[1]), a0, a1);
                        sum += a0 + a1;
                    } else if ((cfull >> c) & 1) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float f0, f1;
                            if (kPolyPairs > 0 && (i & 7) < kPolyPairs) {
                                // these pairs take their 2^x from the FMA pipe (MUFU.EX2 is the busiest pipe of this pass, 4 lanes/clk/SMSP)
                                exp2_poly2(fmaf(__uint_as_float(v[2 * i]), p.scale_log2e, -mx), fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2e, -mx), f0, f1);
                            } else {
                                f0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), p.scale_log2e, -mx));
                                f1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2e, -mx));
                            }
                            sum += f0 + f1;
                            pk[i] = pack_bf16(f0, f1);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float c0 = codes[c * 32 + 2 * i], c1 = codes[c * 32 + 2 * i + 1];
                            const float f0 = ex2_approx((c0 == 0.0f ? __uint_as_float(v[2 * i]) * p.scale_log2e : c0) - mx);
                            const float f1 = ex2_approx((c1 == 0.0f ? __uint_as_float(v[2 * i + 1]) * p.scale_log2e : c1) - mx);
                            sum += f0 + f1;
                            pk[i] = pack_bf16(f0, f1);
                        }
                    }
                    tmem_st_32x16(t_s + c * 16, pk);  // chunk c lands in columns already read (S chunk c / 2)
                }
                tmem_st_wait();
            }
            if constexpr (NKB > 1) s_xsum[w * 128 + r] = sum;
            tc_fence_before();  // S fully read, P written: the MMA warp may consume P and, later, overwrite the slot
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[w]);
            stamp(3);

            // epilogue: O / rowsum -> bf16 -> ctx (merged-head layout)
            mbar_wait(&o_full[slot], par);
            stamp(4);
            tc_fence_after();
            if constexpr (NKB > 1) {
                sum = 0.0f;
#pragma unroll
                for (int i = 0; i < NKB; ++i) sum += s_xsum[(slot * NKB + i) * 128 + r];
            }
            float inv = 1.0f / sum;
            if (poison) inv = __int_as_float(0x7fc00000);
            if constexpr (NKB == 1) {
                uint32_t o[D];
                tmem_ld_32x32(t_slot + Cfg::kOCol, reinterpret_cast<uint32_t(&)[32]>(o[0]));
                if constexpr (D == 64) tmem_ld_32x32(t_slot + Cfg::kOCol + 32, reinterpret_cast<uint32_t(&)[32]>(o[32]));
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_empty[slot]);
                uint8_t* buf = sout + (Cfg::kOutBufs == 2 ? (n & 1) : 0) * Cfg::kOutWarpBytes;
                if (lane == 0) {  // the store that last used this staging buffer has read it
                    if constexpr (Cfg::kOutBufs == 2) bulk_wait_read<1>();
                    else bulk_wait_read<0>();
                }
                __syncwarp();
                const uint32_t obase = smem_u32(buf) + lane * Cfg::kRowBytes;
                const uint32_t sw = D == 32 ? ((lane >> 1) & 3) : (lane & 7);
#pragma unroll
                for (int i = 0; i < D / 8; ++i) {
                    st_shared_v4(obase + ((static_cast<uint32_t>(i) ^ sw) << 4),
                                 pack_bf16(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv),
                                 pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv),
                                 pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv),
                                 pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv));
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && quad * 32 < p.S) {
                    tma_store_3d(&tmap_ctx, buf, h * D, quad * 32, b);  // rows >= S are clipped by the tensor map
                    bulk_commit();
                }
                stamp(5);
            } else {
                // each warpgroup of the slot stores D / NKB columns of every row, straight from registers (16-byte pieces)
                constexpr int CW = Cfg::kOutCols;
                uint32_t o[CW];
                if constexpr (CW == 8) tmem_ld_32x8(t_slot + Cfg::kOCol + j * CW, reinterpret_cast<uint32_t(&)[8]>(o[0]));
                else if constexpr (CW == 16) tmem_ld_32x16(t_slot + Cfg::kOCol + j * CW, reinterpret_cast<uint32_t(&)[16]>(o[0]));
                else tmem_ld_32x32(t_slot + Cfg::kOCol + j * CW, reinterpret_cast<uint32_t(&)[32]>(o[0]));
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_empty[slot]);
                const int row = qb * 128 + r;
                if (row < p.S) {
                    __nv_bfloat16* dst = p.ctx + (static_cast<size_t>(b) * p.S + row) * p.H + h * D + j * CW;
#pragma unroll
                    for (int i = 0; i < CW / 8; ++i) {
                        uint4 q;
                        q.x = pack_bf16(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
                        q.y = pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
                        q.z = pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
                        q.w = pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
                        reinterpret_cast<uint4*>(dst)[i] = q;
                    }
                }
                stamp(5);
            }
        }
        if (Cfg::kOutBufs > 0 && lane == 0) bulk_wait_read<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
    if (threadIdx.x == 0) gstamp(62);
}

}  // namespace kj
