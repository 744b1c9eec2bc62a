// Persistent warp-specialised tcgen05 GEMM for the encoder projections:
//     C[M,N] = epilogue( A[M,K] (bf16, K-major) x W[N,K]^T (bf16, K-major) )
// This is the B200 replacement for LinearLayer::matmul / matmul_noalloc
// (reference: kjarni-transformers/src/linear_layer/linear_layer.rs:160-282 ->
// cpu/ops/matmul.rs:370-479 -> cpu/kernels/x86/f32.rs:9-124), with the bias,
// activation (cpu/feedforward/standard_new.rs:65-73) and residual add
// (cpu/encoder/encoder_layer.rs:129-136,156-163) fused into the epilogue.
//
// Structure (one CTA per SM, 384 threads):
//   warp 0      TMA producer: A tile [128 x 64] + W tile [BN x 64] per stage, 128B swizzle
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma 128 x BN x 16, fp32 accum in TMEM
//   warp 2      TMEM allocator (2 accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 4-11  epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> bias/GELU/residual -> global
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace kj {

enum GemmEpilogue : int {
    EPI_BIAS_BF16 = 0,      // out bf16 = acc + bias
    EPI_BIAS_ACT_BF16 = 1,  // out bf16 = act(acc + bias)
    EPI_BIAS_RES_F32 = 2,   // out f32  = acc + bias + residual(f32)
    EPI_BIAS_F32 = 3,       // out f32  = acc + bias
};
enum Activation : int { ACT_GELU_ERF = 0, ACT_GELU_TANH = 1, ACT_RELU = 2, ACT_NONE = 3 };

struct GemmParams {
    int M, N, K;
    const float* bias;      // [N] or nullptr
    const __nv_bfloat16* residual;  // [M, ldr] bf16 residual stream (EPI_BIAS_RES_F32)
    const float* residual32;        // [M, ldr] fp32 residual stream: used instead of `residual` when non-null (fp32-residual mode)
    void* out;              // [M, ldo] bf16 or f32
    int ldo, ldr;
    int act;
    int dbg;  // microbenchmark switches (kjc_dbg_gemm_time): 1 = no epilogue work, 2 = no MMA issue, 4 = no TMA loads
    unsigned long long* trace;  // optional [gridDim.x][32] %globaltimer stamps of CTA milestones (kjc_dbg_gemm_time flag 8)
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define KJ_TRACE(slot) do { if (p.trace) { p.trace[blockIdx.x * 32 + (slot)] = gtimer(); if ((slot) == 2 || (slot) == 16) p.trace[blockIdx.x * 32 + 20 + ((slot) == 16)] = clock64(); } } while (0)

constexpr int kGemmBlockM = 128;
constexpr int kGemmBlockK = 64;
constexpr int kGemmThreads = 384;
constexpr int kGemmEpiWarp0 = 4;

constexpr int kEpiWarps = 8;
constexpr int kEpiChunkCols = 32;                         // columns per tcgen05.ld / per TMA-store box
constexpr int kEpiStageBytes = 32 * kEpiChunkCols * 2;    // one warp's bf16 staging tile: 32 rows x 64 B (64B swizzle)
constexpr int kEpiBiasMax = 3072;                         // bias columns staged in shared memory (BN <= 192 configurations)

#ifndef KJ_GEMM_PAIR_DEFAULT
#define KJ_GEMM_PAIR_DEFAULT 3  // encoder: QKV (bit 0) and FFN-up (bit 1) as CTA pairs when their tiles are 192 / 256 columns wide
#endif
#ifndef KJ_GEMM_PARTS192
#define KJ_GEMM_PARTS192 3
#endif
#ifndef KJ_GEMM_WIDE256
#define KJ_GEMM_WIDE256 1
#endif
constexpr bool kGemm192WideStore = KJ_GEMM_PARTS192 == 3;
// Tiles whose bf16 epilogue stages 32 x 64 parts (128B swizzle) and issues one TMA store per warp per tile; the host builds the
// output tensor map with a 32 x 64 box for these widths.
constexpr bool gemm_wide_store(int bn) { return (bn == 192 && kGemm192WideStore) || (bn == 256 && KJ_GEMM_WIDE256 != 0); }

// kPair (gemm_tcgen05_kernel<BN, EPI, true>): two CTAs of a cluster run two row tiles against the same weight tile as one
// tcgen05.mma.cta_group::2 (M = 256), each CTA loading its own A rows and HALF of the weight rows.  A K-streaming 128 x BN tile
// needs (16 + BN / 8) KB from L2 per k-block of BN / 2 clk x 4 MMAs -- 87 flop/B at BN = 256, and L2 delivers ~10 TB/s to 148 SMs
// streaming at once: ~870 TFLOP/s, which is what the one-CTA kernel measures on the hidden-768 projections (0.63-0.65 of the
// sustained roof).  Half the weight bytes per SM is 131 flop/B, and the smaller stages make the ring deeper in MMA time.
template <int BN, bool kPair = false>
struct GemmCfg {
    static constexpr bool kWide = gemm_wide_store(BN);
    // 256-column tiles: a tcgen05.mma of M = 128 costs ~128 clk whatever N <= 256 is (measured, scripts/mma_rate.py), so the widest
    // tile wastes no tensor time; its 16 epilogue warps need 64 KB of staging, hence 3 operand stages of 48 KB
    static constexpr int kABytes = kGemmBlockM * kGemmBlockK * 2;
    static constexpr int kBBytes = (kPair ? BN / 2 : BN) * kGemmBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
    // epilogue warps: 4 TMEM lane quadrants x kParts column parts.  Wide-store tiles use parts of 64 columns (12 / 16 warps, 3 / 4 per
    // scheduler: the epilogue is a latency chain per warp, more resident warps hide its waits); the other widths use 2 parts.
    static constexpr int kParts = kWide ? BN / 64 : 2;
    static constexpr int kEpiWarpsN = 4 * kParts;
    static constexpr int kThreads = 128 + 32 * kEpiWarpsN;
    static constexpr int kColsPerPart = BN / kParts;
    // output staging per epilogue warp: wide-store tiles stage the warp's whole 32 x 64 part (4 KB, 128-byte rows, one TMA store
    // per tile: the TMA store path is charged per row segment, 128-byte segments halve its load); others 2-3 x (32 x 32) chunks
    static constexpr int kStoreCols = kWide ? 64 : kEpiChunkCols;
    static constexpr int kEpiBufs = kWide ? 1 : ((BN <= 192) ? 3 : 2);
    static constexpr int kEpiBufBytes = 32 * kStoreCols * 2;
    static constexpr int kEpiBytes = kEpiWarpsN * kEpiBufs * kEpiBufBytes;
    static constexpr int kBiasBytes = (BN <= 192 || kWide) ? kEpiBiasMax * 4 : 0;  // bias staged in smem where it fits
    static constexpr int kPairStages = (232448 - 1024 - 256 - kBiasBytes - kEpiBytes) / kStageBytes;  // as many as fit (BN 256: 4 x 32 KB, 192: 5 x 28 KB)
    static constexpr int kStages = kPair ? (kPairStages > 8 ? 8 : kPairStages) : ((BN <= 64) ? 6 : ((BN <= 128) ? 5 : ((BN == 256 && kWide) ? 3 : 4)));
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kBiasBytes;
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// erf-GELU, 0.5*x*(1+erf(x/sqrt2)) (reference activations.rs:57-59), with erf(t) = 1 - 2^-q(t) for t >= 0,
// q a degree-5 polynomial without constant term fitted to -log2(erfc(t)) (monotone, so no clamp is needed;
// max |erf| error 7e-6, max |GELU| error 8.5e-6 -- far below the bf16 rounding of the stored activation).  The polynomial is
// evaluated in |x| directly (the 1/sqrt2 of t = |x|/sqrt2 is folded into the coefficients: same accuracy, one multiply less).
// Two elements per call on the packed fp32x2 pipe: 8 FFMA2/FMUL2 + 2 FABS + 2 MUFU.EX2 per pair.
__device__ __forceinline__ void gelu_erf_fast2(float& x0, float& x1) {
    const uint64_t X = f2_pack(x0, x1);
    const uint64_t A = f2_pack(fabsf(x0), fabsf(x1));
    uint64_t Q = f2_fma(A, f2_pack(-0.0005145793f, -0.0005145793f), f2_pack(0.0074325f, 0.0074325f));  // coefficients negated: Q = -q(|x|/sqrt2)
    Q = f2_fma(Q, A, f2_pack(-0.052670788f, -0.052670788f));
    Q = f2_fma(Q, A, f2_pack(-0.4591722f, -0.4591722f));
    Q = f2_fma(Q, A, f2_pack(-1.1511078f, -1.1511078f));
    Q = f2_mul(Q, A);
    float q0, q1, e0, e1;
    f2_unpack(Q, q0, q1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(q0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q1));
    const uint64_t W = f2_fma(f2_pack(e0, e1), f2_pack(-1.0f, -1.0f), f2_pack(1.0f, 1.0f));  // erf(|x|/sqrt2)
    const uint64_t R = f2_mul(f2_fma(A, W, X), f2_pack(0.5f, 0.5f));                          // 0.5*(x + |x|*erf)
    f2_unpack(R, x0, x1);
}
// f[0..4) += b on the packed pipe (two FADD2 instead of four FADD; the same IEEE additions)
__device__ __forceinline__ void add_bias4(float* f, const float4& b) {
    f2_unpack(f2_add(f2_pack(f[0], f[1]), f2_pack(b.x, b.y)), f[0], f[1]);
    f2_unpack(f2_add(f2_pack(f[2], f[3]), f2_pack(b.z, b.w)), f[2], f[3]);
}
__device__ __forceinline__ float gelu_tanh_fast(float x) {
    // gelu_new_scalar, activations.rs:62-66
    const float inner = 0.7978845608f * fmaf(0.044715f * x * x, x, x);
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(inner));
    return 0.5f * x * (1.0f + th);
}
// Activation over a register tile; the (warp-uniform) switch sits outside the element loop so the
// 32 independent polynomial chains interleave.
template <int NELEM>
__device__ __forceinline__ void apply_act_tile(float (&f)[NELEM], int act) {
    if (act == ACT_GELU_ERF) {
#pragma unroll
        for (int j = 0; j < NELEM; j += 2) gelu_erf_fast2(f[j], f[j + 1]);
    } else if (act == ACT_GELU_TANH) {
#pragma unroll
        for (int j = 0; j < NELEM; ++j) f[j] = gelu_tanh_fast(f[j]);
    } else if (act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < NELEM; ++j) f[j] = fmaxf(f[j], 0.0f);
    }
}

template <int BN, int EPI, bool kPair = false>
__global__ void __launch_bounds__(GemmCfg<BN, kPair>::kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, GemmParams p) {
    using Cfg = GemmCfg<BN, kPair>;
    static_assert(!kPair || (BN % 32 == 0 && BN >= 128 && 2 * Cfg::kStages + 5 <= 32), "pair tiles: 128-256 columns");
    constexpr int kStages = Cfg::kStages;
    static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");

    extern __shared__ uint8_t smem_raw[];
    if (threadIdx.x == 0) KJ_TRACE(0);  // kernel entry
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * Cfg::kABytes;
    uint8_t* smem_epi = smem + kStages * Cfg::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::kEpiBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tmem_full = bars + 2 * kStages;
    uint64_t* tmem_empty = bars + 2 * kStages + 2;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
    const bool bias_in_smem = Cfg::kBiasBytes > 0 && p.bias != nullptr && p.N <= kEpiBiasMax;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // the bias is a weight: it does not depend on the predecessor kernel, so it is staged before pdl_wait()
    if (bias_in_smem)
        for (int i = threadIdx.x; i < p.N; i += Cfg::kThreads) s_bias[i] = __ldg(p.bias + i);

    const int m_tiles = (p.M + kGemmBlockM - 1) / kGemmBlockM;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int k_blocks = (p.K + kGemmBlockK - 1) / kGemmBlockK;
    // work units: (row tile, column tile), or for pairs (two consecutive row tiles, column tile) with this CTA on row tile 2 u + rank
    // (a row tile beyond M is all padding: loads zero-filled, stores clipped)
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int worker = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int n_workers = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int num_tiles = (kPair ? (m_tiles + 1) / 2 : m_tiles) * n_tiles;
    auto row_tile_of = [&](int tile) { return kPair ? 2 * (tile / n_tiles) + static_cast<int>(rank) : tile / n_tiles; };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_ACT_BF16) tma_prefetch_desc(&tmap_c);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], (kPair ? 2 : 1) * Cfg::kEpiWarpsN);  // one arrive per epilogue warp (pairs: of both CTAs, on the leader's)
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        if constexpr (kPair) tmem_alloc_2sm<Cfg::kTmemCols>(tmem_base_smem);
        else tmem_alloc<Cfg::kTmemCols>(tmem_base_smem);
    }
    tc_fence_before();
    if constexpr (kPair) cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / TMA completion
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    if (threadIdx.x == 0) KJ_TRACE(1);  // prologue done
    pdl_wait();               // everything above overlapped the previous kernel's tail; its outputs are visible from here on
    pdl_launch_dependents();  // the next kernel may begin its own prologue as soon as this CTA's resources are released
    if (threadIdx.x == 0) KJ_TRACE(2);  // predecessor complete

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = worker; tile < num_tiles && !(p.dbg & 64); tile += n_workers) {
                const int m_blk = row_tile_of(tile), n_blk = tile % n_tiles;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if constexpr (kPair) {
                        // this CTA's A rows and its half of the weight tile; all bytes are counted on the LEADER's barrier
                        const uint32_t lbar = mapa_shared(smem_u32(&full_bar[stage]), 0);
                        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
                        tma_load_2d_2sm(smem_a + stage * Cfg::kABytes, &tmap_a, lbar, kb * kGemmBlockK, m_blk * kGemmBlockM, kEvictFirst);
                        tma_load_2d_2sm(smem_b + stage * Cfg::kBBytes, &tmap_b, lbar, kb * kGemmBlockK, n_blk * BN + static_cast<int>(rank) * (BN / 2), kEvictLast);
                    } else if (p.dbg & 4) {
                        mbar_arrive(&full_bar[stage]);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                        tma_load_2d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * kGemmBlockK, m_blk * kGemmBlockM,
                                    kEvictFirst);
                        tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * kGemmBlockK, n_blk * BN, kEvictLast);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer
        if ((KJ_MMA_UNIFORM != 0 || lane == 0) && leader) {  // pairs: the leader issues for both CTAs
            constexpr uint32_t idesc = umma_idesc(1 /*bf16*/, kPair ? 2 * kGemmBlockM : kGemmBlockM, BN);
            auto commit = [&](uint64_t* bar) {
                if constexpr (kPair) umma_commit_2sm(bar, 3);  // the barrier at this offset in both CTAs
                else umma_commit(bar);
            };
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = worker; tile < num_tiles; tile += n_workers, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                if constexpr (kPair) mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);  // the peer's arrives are remote
                else mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    if (!(p.dbg & 64)) mbar_wait(&full_bar[stage], phase);  // 64: raw MMA issue rate (no operand handshake)
                    if (it == 0 && kb == 0 && lane == 0) KJ_TRACE(3);  // first operands landed
                    tc_fence_after();
                    const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
                    const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
                    if (mma_issuer_lane()) {
                        if (!(p.dbg & 2)) {
#pragma unroll
                            for (int k = 0; k < kGemmBlockK / 16; ++k) {
                                // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in (addr >> 4) units
                                if constexpr (kPair) umma_f16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                                else umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                            }
                        }
                        if (!(p.dbg & 64)) commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                        if (kb == k_blocks - 1) commit(&tmem_full[acc]);
                    }
                    mma_issuer_sync();
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp >= kGemmEpiWarp0) {
        // ---------------------------------------------------------- epilogue
        const int ew = warp - kGemmEpiWarp0;  // 0 .. kEpiWarpsN-1
        const int quad = warp & 3;            // TMEM lane quadrant this warp may access
        const int half = ew >> 2;             // column part handled by this warpgroup
        constexpr int kColsPerHalf = Cfg::kColsPerPart;
        constexpr bool kStaged = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_ACT_BF16);
        uint8_t* stage_buf = smem_epi + ew * Cfg::kEpiBufs * Cfg::kEpiBufBytes;
        int sbuf = 0;
        int it = 0;
        // "this warp has read the accumulator": pairs arrive on the leader's barrier, without the cluster-scope release (the reads are
        // complete; the release form waits for the warp's earlier shared-memory / TMA traffic, ~1400 clk per tile in the chained kernels)
        auto release_acc = [&](int a) {
            if constexpr (kPair) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tmem_empty[a]), 0));
            else mbar_arrive(&tmem_empty[a]);
        };
        for (int tile = worker; tile < num_tiles; tile += n_workers, ++it) {
            const int m_blk = row_tile_of(tile), n_blk = tile % n_tiles;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tmem_full[acc], acc_phase);
            if (ew == 0 && lane == 0 && it < 6) KJ_TRACE(4 + 2 * it);  // accumulator of tile `it` ready
            tc_fence_after();
            const int row0 = m_blk * kGemmBlockM + quad * 32;
            const int row = row0 + lane;
            const bool row_ok = row < p.M;
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + half * kColsPerHalf;
            if (p.dbg & 1) {
                // microbenchmark: accumulator released untouched
            } else if constexpr (kStaged && Cfg::kStoreCols == 64) {
                // 192-column tiles: this warp owns a 32-row x 64-column part.  TMEM -> registers -> bias/activation -> bf16 ->
                // 128B-swizzled 4 KB smem tile -> ONE TMA store per tile.  The accumulator is released as soon as the last
                // tcgen05.ld has landed.
                static_assert(kColsPerHalf == 64, "part width");
                const int colp = n_blk * BN + half * kColsPerHalf;
                const bool tr = p.trace != nullptr && ew == 5 && lane == 0 && it == 2;  // epilogue sub-steps of one warp, third tile
                if (tr) p.trace[blockIdx.x * 32 + 22] = gtimer();
                if (lane == 0) bulk_wait_read<0>();  // the previous tile's store has read the staging tile
                __syncwarp();
                if (tr) p.trace[blockIdx.x * 32 + 23] = gtimer();
                const uint32_t rbase = smem_u32(stage_buf) + lane * 128;
                const uint32_t sw = lane & 7;  // 128B swizzle: 16-byte chunk index ^= row & 7
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    if (p.dbg & 128) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = lane + j;
                    } else {
                        tmem_ld_32x32(taddr0 + c * 32, v);
                        tmem_ld_wait();
                    }
                    if (tr) p.trace[blockIdx.x * 32 + 24 + c] = gtimer();
                    if (c == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) release_acc(acc);
                    }
                    const int col0 = colp + c * 32;
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    if (bias_in_smem && col0 + 32 <= p.N && !(p.dbg & 512)) {  // whole chunk inside N: no per-column checks
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            add_bias4(f + 4 * j, *reinterpret_cast<const float4*>(s_bias + col0 + 4 * j));
                        }
                    } else if (p.bias != nullptr && !(p.dbg & 512)) {
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
                    if (EPI == EPI_BIAS_ACT_BF16 && !(p.dbg & 512)) apply_act_tile(f, p.act);
                    if (!(p.dbg & 256))
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        st_shared_v4(rbase + ((static_cast<uint32_t>(c * 4 + j) ^ sw) << 4), pack_bf16(f[8 * j + 0], f[8 * j + 1]),
                                     pack_bf16(f[8 * j + 2], f[8 * j + 3]), pack_bf16(f[8 * j + 4], f[8 * j + 5]),
                                     pack_bf16(f[8 * j + 6], f[8 * j + 7]));
                    }
                }
                if (tr) p.trace[blockIdx.x * 32 + 26] = gtimer();
                if (!(p.dbg & 32)) fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && colp < p.N && !(p.dbg & 16)) {  // columns >= N are clipped by the tensor map
                    tma_store_2d(&tmap_c, stage_buf, colp, row0);
                    bulk_commit();
                }
                if (tr) p.trace[blockIdx.x * 32 + 27] = gtimer();
            } else if constexpr (kStaged) {
                // TMEM -> registers -> bias/activation -> bf16 -> swizzled smem tile -> TMA store (coalesced, async).
                // The tcgen05.ld of chunk c+1 is in flight while chunk c is processed; the accumulator is released as soon
                // as the last load has landed, before the math and stores of the last chunk.
                constexpr int kChunks = kColsPerHalf / kEpiChunkCols;
                constexpr int kBufs = Cfg::kEpiBufs;
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr0 + c * kEpiChunkCols, v);
                    tmem_ld_wait();
                    if (c + 1 == kChunks) {  // last load landed: release the accumulator before the math and stores of this chunk
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) release_acc(acc);
                    }
                    const int col0 = n_blk * BN + half * kColsPerHalf + c * kEpiChunkCols;
                    if (col0 < p.N) {
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                        if (p.bias != nullptr) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                if (col0 + 4 * j < p.N) {
                                    const float4 b = bias_in_smem ? *reinterpret_cast<const float4*>(s_bias + col0 + 4 * j)
                                                                  : __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
                                    f[4 * j + 0] += b.x;
                                    f[4 * j + 1] += b.y;
                                    f[4 * j + 2] += b.z;
                                    f[4 * j + 3] += b.w;
                                }
                            }
                        }
                        if (EPI == EPI_BIAS_ACT_BF16) apply_act_tile(f, p.act);
                        // the staging buffer we are about to overwrite must have been read by its previous TMA store
                        if (lane == 0) bulk_wait_read<kBufs - 1>();
                        __syncwarp();
                        uint8_t* buf = stage_buf + sbuf * kEpiStageBytes;
                        const uint32_t rbase = smem_u32(buf) + lane * 64;
                        const uint32_t sw = (lane >> 1) & 3;  // 64B swizzle: 16-byte chunk index ^= (row >> 1) & 3
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            st_shared_v4(rbase + ((j ^ sw) << 4), pack_bf16(f[8 * j + 0], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                         pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
                        }
                        if (!(p.dbg & 32)) fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0 && !(p.dbg & 16)) {
                            tma_store_2d(&tmap_c, buf, col0, row0);
                            bulk_commit();
                        }
                        if (++sbuf == kBufs) sbuf = 0;
                    }
                }
            } else {
                constexpr int kChunks = kColsPerHalf / 16;
#pragma unroll 1
                for (int c = 0; c < kChunks; ++c) {
                    uint32_t v[16];
                    tmem_ld_32x16(taddr0 + c * 16, v);
                    tmem_ld_wait();
                    const int col0 = n_blk * BN + half * kColsPerHalf + c * 16;
                    if (col0 < p.N) {  // N tail (N is a multiple of 16)
                        float f[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
                        if (p.bias != nullptr) {
                            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 b = __ldg(b4 + j);
                                f[4 * j + 0] += b.x;
                                f[4 * j + 1] += b.y;
                                f[4 * j + 2] += b.z;
                                f[4 * j + 3] += b.w;
                            }
                        }
                        if (row_ok) {
                            float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col0;
                            if (EPI == EPI_BIAS_RES_F32 && p.residual32 != nullptr) {
                                const float4* r4 = reinterpret_cast<const float4*>(p.residual32 + static_cast<size_t>(row) * p.ldr + col0);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float4 r = __ldg(r4 + j);
                                    f[4 * j + 0] += r.x;
                                    f[4 * j + 1] += r.y;
                                    f[4 * j + 2] += r.z;
                                    f[4 * j + 3] += r.w;
                                }
                            } else if (EPI == EPI_BIAS_RES_F32) {
                                const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + static_cast<size_t>(row) * p.ldr + col0);
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    const uint4 r = __ldg(r4 + j);
                                    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        f[8 * j + 2 * e] += __uint_as_float(w[e] << 16);
                                        f[8 * j + 2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
                                    }
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                reinterpret_cast<float4*>(o)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                        }
                    }
                }
            }
            // all tcgen05.ld of this warp have completed (wait::ld above): release the accumulator stage
            // (the staged path released it right after its last load)
            if (!kStaged || (p.dbg & 1)) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(acc);
            }
            if (ew == 0 && lane == 0 && it < 6) KJ_TRACE(5 + 2 * it);  // epilogue of tile `it` done (stores issued)
        }
        if (kStaged && lane == 0) bulk_wait_read<0>();  // smem must stay valid until the last store has read it
        if (ew == 0 && lane == 0) KJ_TRACE(16);  // stores drained
    }

    tc_fence_before();
    if constexpr (kPair) cluster_sync_all();  // peer shared memory / barriers stay valid until both CTAs are done
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (kPair) tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
        else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
    if (threadIdx.x == 64) KJ_TRACE(17);  // exit
}

}  // namespace kj
