// Two dependent projections of the encoder layer in ONE launch, for hidden size 384 (MiniLM):
//     x'  [M,384]  (bf16) = LayerNorm( A[M,K1] x W1[384,K1]^T + b1 + x[M,384] ; gamma, beta, eps )     (phase 1)
//     out [M,N2]   (bf16) = act( x' x W2[N2,384]^T + b2 )                                                (phase 2)
// i.e.  attention out-proj + residual + LN1  ->  FFN-up (+ erf-GELU)            (encoder_layer.rs:120-147 + standard_new.rs:47-73)
// and   FFN-down + residual + LN2            ->  the NEXT layer's fused QKV     (encoder_layer.rs:150-176 + qkv_projection.rs:93-138).
//
// Why: every op of the layer is local to a 128-token tile, so the CTA that produced tile b of x' can run the next projection of
// tile b straight away.  Against two launches this removes one launch's fill and drain bubbles (~5 us of a ~20 us launch,
// DESIGN.md section 5) and the re-load of x' as the A operand (it never leaves shared memory: 6 k-blocks of 16 KB written by the
// LayerNorm epilogue directly in the 128B-swizzled K-major layout the tensor core reads).
//
// One tile per CTA (the host falls back to the two separate kernels when there are more 128-row tiles than SMs).
// Shared memory (bytes), one CTA per tile -- phase 2 lives entirely inside phase 1's operand ring:
//     [0, 48K)       phase 1: A ring, 3 x 16 KB               | phase 2: W2 stages 0, 1 (2 x 24 KB)
//     [48K, 96K)     phase 1: W ring stage 0 (= LN staging)   | phase 2: W2 stage 2 (24 KB) + output staging of warps 0-5 (6 x 4 KB)
//     [96K, 192K)    phase 1: W ring stages 1, 2              | phase 2: x' tile, 6 x 16 KB (A operand, resident)
//     [192K, 206K)   LayerNorm statistics, bias / gamma / beta, phase-2 bias, barriers (warp 6 stages its output on the dead statistics)
//     [206K, 226K)   output staging of warps 7-11
// TMEM: phase 1 accumulates the full 128 x 384 rows in columns [0, 384); phase 2 double-buffers 128 x 192 accumulators in
// [0, 192) and [192, 384).  Warp roles as in gemm_ln.cuh (warp 0 TMA, warp 1 MMA, warp 2 TMEM, warps 4-15 epilogue).
// Results are bit-identical to gemm_ln384_kernel followed by gemm_tcgen05_kernel<192, EPI>: same MMA shapes, same k order.
//
// kPair = true (the default of the encoder, KJC_CHAIN_PAIR): two CTAs of a cluster (one TPC) run their two row tiles together with
// tcgen05.mma.cta_group::2 (M = 256: 128 rows per CTA).  Each CTA loads only HALF of every weight tile (96 of the 192 rows of
// W1 / W2 per MMA) and the tensor core reads both halves.  What this buys (scripts/chain_trace.py, profiles/r02_chain_*): with one
// CTA per tile every SM streams the whole weight matrix out of L2 (148 x 1.2 MB per launch, ~18-24 TB/s at the MMA rate against
// the ~10 TB/s L2 delivers), a TMA round trip takes ~1500 clk under that load, and a three-stage ring holds 384-768 clk of MMAs
// per stage: phase 2 and FFN-down's phase 1 ran latency-bound at 180 clk per MMA.  Half-size weight stages make the same shared
// memory a ring twice as deep in MMA time -- four 40 KB phase-1 stages (A 16 KB | two 12 KB weight halves) and six 12 KB W2 stages
// -- and halve the L2 traffic: FFN-down's phase 1 now runs at 98 clk per MMA (floor 96), phase 2 at ~120.  Rows are still owned
// by one CTA (its 128 TMEM lanes), so LayerNorm, the x' tile and both epilogues are unchanged; only the leader CTA issues MMAs, the
// leader's barriers count both CTAs' TMA bytes, commits are multicast to both CTAs and the peer's epilogue warps release
// accumulators / publish x' on the leader's barriers.  The accumulator release is a RELAXED cluster arrive: the release form cost
// every epilogue warp ~1400 clk per tile (it waits for the warp's earlier shared-memory / TMA traffic to drain), which is what made
// the first pair variant of this kernel slower than the one-CTA form.  The residual staging has memory of its own behind the
// 160 KB ring ([160K, 192K) + 16 KB of the output staging), so all six W2 stages are prefetched under the LayerNorm epilogue.
//
// kTS = true (opt-in, KJC_CHAIN_TS, one CTA per tile): phase 2 reads x' from TENSOR MEMORY instead of shared memory.  The LayerNorm
// epilogue ALSO writes x' as packed bf16 into TMEM (tcgen05.st; columns [0,64) in place for column part 0, [384,512) for parts 1
// and 2) and phase 2 issues tcgen05.mma with the A operand in TMEM ("TS" form).  TMEM then has room for two 128-column accumulators
// ([64,192), [192,320)), so phase-2 tiles are 128 wide with five 16 KB W2 stages; the x' tile in shared memory remains only as the
// source of the TMA store to global memory.  Bit-identical.  Measured equal to the shared-memory form (33.1 vs 33.0 us): the A
// operand reads were not what bounded phase 2 (the W2 stream was), so it is not the default.
#pragma once
#include <cuda.h>

#include "gemm_ln.cuh"

namespace kj {

constexpr int kLg2BN = 192;                                 // phase-2 tile width
constexpr int kLg2WBytes = kLg2BN * kGemmBlockK * 2;        // 24 KB per W2 stage
constexpr int kLg2Stages = 3;
constexpr int kLg2KB = kLnN / kGemmBlockK;                  // 6 k-blocks of the resident x' tile
constexpr int kLg2BiasMax = 1536;                           // phase-2 bias columns staged in shared memory
constexpr int kLg2RingBytes = 3 * kLnStageBytes;            // 192 KB: phase 1 always runs a 3-stage ring here
constexpr int kLg2BaseBytes = kLg2RingBytes + kLnStatBytes + kLnVecBytes + kLg2BiasMax * 4 + 512;
// Phase-2 output staging: every epilogue warp stages its whole 32 x 64 part of a 192-column tile (4 KB, 128-byte rows) and issues ONE
// TMA store per tile (one proxy fence, one store, no wait for the first chunk's store inside the tile) -- the same form as the
// stand-alone GEMM's 192-column tiles.  Measured neutral against two 32 x 32 chunks through one 2 KB buffer (35.1 vs 34.6 us): the
// store chain was not what paced phase 2 either.  Warps 0-5 stage in [72K, 96K) (the rest of the LayerNorm staging), warp 6 on the
// LayerNorm statistics / bias1 (dead once x' is published), warps 7-11 in 20 KB behind the barriers.
constexpr int kLg2WideBytes = 32 * 64 * 2;                  // 4 KB per warp
constexpr int kLg2SmemBytes = kLg2BaseBytes + 5 * kLg2WideBytes;
static_assert(kLg2BaseBytes % 1024 == 0 && kLg2SmemBytes <= 232448, "shared memory budget");
static_assert(kLg2WBytes + 6 * kLg2WideBytes <= kLnBBytes && kLg2WideBytes <= kLnStatBytes + kLnN * 4, "wide staging aliases");
static_assert(2 * kLg2WBytes <= 3 * kLnABytes && kLg2WBytes + kLnEpiWarps * kEpiStageBytes <= kLnBBytes, "phase-2 aliasing");
// kTS (phase 2 takes its A operand from TENSOR MEMORY): tiles of 128 columns, five 16 KB W2 stages
#ifndef KJ_LG_EARLY
#define KJ_LG_EARLY 0  // kPair: bit 0 = weight halves of the first stages before griddepcontrol.wait, bit 1 = residual chunks before the accumulator is
                       // complete.  Measured 1.2 % SLOWER in the whole step (262-263 vs 266 k emb/s): the early loads compete with the
                       // predecessor's tail and with phase 1 for the same L2 bandwidth.  Compiled out.
#endif
#ifndef KJ_CHAIN_PAIR_DEFAULT
#define KJ_CHAIN_PAIR_DEFAULT 3
#endif
#ifndef KJ_CHAIN_TS_DEFAULT
#define KJ_CHAIN_TS_DEFAULT 0
#endif
constexpr int kLgTBN = 128;
constexpr int kLgTWBytes = kLgTBN * kGemmBlockK * 2;        // 16 KB per W2 stage
constexpr int kLgTStages = 5;                               // stages 0-2: phase-1 A ring, 3: own region behind the tail, 4: LayerNorm staging
constexpr int kLgTSmemBytes = kLg2BaseBytes + kLgTWBytes;
static_assert(kLgTSmemBytes <= 232448, "kTS shared memory budget");
static_assert(3 * kLgTWBytes <= 3 * kLnABytes && kLgTWBytes + kLnEpiWarps * kEpiStageBytes <= kLnBBytes, "kTS phase-2 aliasing");
constexpr int kLgTAcc0 = 64;                                // TMEM: x' in [0,64) | [384,512), accumulators [64,192) and [192,320)
__host__ __device__ constexpr int lg_phase2_bn(bool ts) { return ts ? kLgTBN : kLg2BN; }
__host__ __device__ constexpr int lg_smem_bytes(bool ts) { return ts ? kLgTSmemBytes : kLg2SmemBytes; }

// kTS: first TMEM column of the packed bf16 x' of column part `part` (128 columns of x' = 64 TMEM columns)
__host__ __device__ constexpr uint32_t lg_x_tmem_col(int part) { return part == 0 ? 0u : 384u + 64u * static_cast<uint32_t>(part - 1); }

#ifndef KJ_LG_TRACE
#define KJ_LG_TRACE 0
#endif
#define KJ_LGT(slot) do { if (KJ_LG_TRACE && p.trace != nullptr) p.trace[blockIdx.x * 256 + (slot)] = clock64(); } while (0)

struct GemmLnGemmParams {
    int M, K1;
    int tile_base;       // first 128-row tile of this launch (even for CTA pairs): a micro-batch wider than one tile per SM runs as
                         // several launches over consecutive row chunks of the same tensors
    const float* bias1;  // [384] or nullptr
    const float* gamma;  // [384]
    const float* beta;   // [384]
    float eps;
    int N2;              // phase-2 output columns (multiple of 8; tiles of 192, the last one may be partial)
    const float* bias2;  // [N2] or nullptr
    int act;             // Activation (EPI_BIAS_ACT_BF16)
    unsigned long long* trace;  // KJ_LG_TRACE builds: [gridDim.x][256] clock64 stamps (scripts/chain_micro.py TRACE=1)
    int dbg;             // phase-2 knock-outs of the micro benchmark (kjc_dbg_gemm_ln_gemm): 1 = no W2 loads (barrier arrives only),
                         // 2 = no phase-2 epilogue work (accumulator read + release only), 4 = no phase-2 MMAs
    // P1 == 1 (embedding front end): phase 1 = Embeddings::forward + embed LayerNorm instead of a GEMM (rowwise.cuh EmbedParams)
    const uint32_t* ids;       // [M]
    const uint32_t* type_ids;  // [M] or nullptr (row 0 of the type table for every token)
    const float* word;         // [vocab, 384]
    const float* pos;          // [max_pos, 384] or nullptr
    const float* type;         // [type_vocab, 384] or nullptr
    int* err_flag;             // set to 1 on a token-type id out of range (the reference panics)
    int S, vocab, max_pos, type_vocab, pos_offset;
};

// P1 = 0: phase 1 is the GEMM + residual + LayerNorm described above.
// P1 = 1: phase 1 is the embedding front end -- word[id] (+ pos[offset + s]) (+ type[tt]) -> embed LayerNorm (reference:
//         cpu/embeddings/mod.rs:181-326, transformer_encoder.rs:303-305), gathered by the epilogue warps (thread = token, 128
//         columns each) straight into TMEM for the same two-pass LayerNorm; phase 2 is then layer 0's QKV projection.
template <int EPI2, int P1 = 0, bool kPair = false, bool kTS = false>
__global__ void __launch_bounds__(kLnThreads, 1)
gemm_ln_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_x,
                    const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_out2, GemmLnGemmParams p) {
    static_assert(EPI2 == EPI_BIAS_BF16 || EPI2 == EPI_BIAS_ACT_BF16, "phase 2 stores bf16");
    static_assert(!(kPair && P1 == 1), "the embedding front end runs one CTA per tile");
    static_assert(!(kTS && (kPair || P1 == 1)), "x' in tensor memory: one CTA per tile, GEMM front end");
    constexpr int kBN2 = kTS ? kLgTBN : kLg2BN;             // phase-2 tile width
    // kPair: every CTA holds HALF of each weight tile, so the same shared memory holds rings twice as deep in MMA time: four phase-1
    // stages of 40 KB (A 16 KB | two 12 KB weight halves) and six 12 KB W2 stages.  A TMA round trip under load is ~1500 clk
    // (scripts/chain_trace.py) against 384 clk of MMAs per stage: three stages left both phases latency-bound.
    constexpr int kStages1 = kPair ? 4 : 3;
    constexpr int kStages2 = kTS ? kLgTStages : (kPair ? 6 : kLg2Stages);
    constexpr uint32_t kW2Bytes = kTS ? kLgTWBytes : kLg2WBytes;
    extern __shared__ __align__(1024) uint8_t smem_lg[];
    uint8_t* smem = smem_lg;
    if (smem_u32(smem) & 1023) __trap();
    // phase 1 views
    uint8_t* smem_a = smem;                         // 3 x 16 KB
    uint8_t* smem_b = smem + 3 * kLnABytes;         // 3 x 48 KB
    uint8_t* smem_epi1 = smem_b;                    // LN residual staging (12 x 4 KB), aliased on W stage 0 (kPair: see ln_stage)
    // phase 2 views
    uint8_t* smem_x = smem_b + kLnBBytes;           // x' tile: 6 x 16 KB
    auto w2_stage = [&](int s) -> uint8_t* {
        if constexpr (kTS) return s < 3 ? smem + s * kLgTWBytes : (s == 3 ? smem + kLg2BaseBytes : smem_b);
        else if constexpr (kPair) return smem + s * (kLg2WBytes / 2);  // [0, 72K): stages 0-3 over the phase-1 A ring, 4-5 over the LayerNorm staging
        else return s < 2 ? smem + s * kLg2WBytes : smem_b;
    };
    // phase-1 stage views (kPair: packed 40 KB stages)
    constexpr int kP1PairStage = kLnABytes + kLnBBytes / 2;
    static_assert(4 * kP1PairStage + 8 * kLnEpiBytesPerWarp <= kLg2RingBytes && 4 * kLnEpiBytesPerWarp <= 5 * kLg2WideBytes, "kPair phase-1 ring + residual staging");
    auto p1_a = [&](int st) -> uint8_t* { return kPair ? smem + st * kP1PairStage : smem_a + st * kLnABytes; };
    auto p1_w0 = [&](int st) -> uint8_t* { return kPair ? smem + st * kP1PairStage + kLnABytes : smem_b + st * kLnBBytes; };
    auto p1_w1 = [&](int st) -> uint8_t* { return kPair ? smem + st * kP1PairStage + kLnABytes + kLnBBytes / 4 : smem_b + st * kLnBBytes + kLnHalfN * 128; };
    uint8_t* smem_epi2 = smem_b + kW2Bytes;         // 12 x 2 KB
    uint8_t* tail = smem + kLg2RingBytes;
    float2* stat = reinterpret_cast<float2*>(tail);  // [3][128]
    float* s_bias = reinterpret_cast<float*>(tail + kLnStatBytes);
    float* s_gamma = s_bias + kLnN;
    float* s_beta = s_gamma + kLnN;
    float* s_bias2 = s_beta + kLnN;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias2 + kLg2BiasMax);
    uint64_t* full1 = kPair ? bars + 56 : bars;        // [3] (kPair: [4])
    uint64_t* empty1 = kPair ? bars + 60 : bars + 3;   // [3] (kPair: [4])
    uint64_t* tmem_full1 = bars + 6;      // phase-1 accumulator complete (all phase-1 MMAs retired)
    uint64_t* res_bar = bars + 7;         // [12 warps][2 buffers]
    uint64_t* x_ready = bars + 31;        // x' tile written, LN accumulator consumed, staging region free (12 arrivals)
    uint64_t* full2 = (kTS || kPair) ? bars + 44 : bars + 32;                   // [3] (kTS: [5], kPair: [6])
    uint64_t* empty2 = kTS ? bars + 49 : (kPair ? bars + 50 : bars + 35);       // [3] (kTS: [5], kPair: [6])
    uint64_t* acc_full = bars + 38;       // [2]
    uint64_t* acc_empty = bars + 40;      // [2] (12 arrivals each; kPair: the leader's, 24 arrivals = both CTAs' epilogue warps)
    uint64_t* x_pair = bars + 42;         // kPair, leader's: both CTAs' x' tiles are written (24 arrivals)
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 43);
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    constexpr int kWRows = kPair ? kLnHalfN / 2 : kLnHalfN;          // weight rows this CTA loads per 192-column MMA tile
    constexpr uint32_t kWBoxBytes = kWRows * kGemmBlockK * 2;        // 24 KB, or 12 KB per CTA of a pair
    constexpr int kEpiArrivals = kPair ? 2 * kLnEpiWarps : kLnEpiWarps;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tile = p.tile_base + blockIdx.x;  // one 128-row tile per CTA; a launch covers tiles [tile_base, tile_base + gridDim.x)
    const int k_blocks1 = (p.K1 + kGemmBlockK - 1) / kGemmBlockK;
    const int n2_tiles = (p.N2 + kBN2 - 1) / kBN2;
    const bool bias2_in_smem = p.bias2 != nullptr && p.N2 <= kLg2BiasMax;

    // weights: independent of the predecessor kernel, staged before griddepcontrol.wait
    for (int i = threadIdx.x; i < kLnN; i += kLnThreads) {
        s_bias[i] = p.bias1 != nullptr ? __ldg(p.bias1 + i) : 0.0f;
        s_gamma[i] = __ldg(p.gamma + i);
        s_beta[i] = __ldg(p.beta + i);
    }
    if (bias2_in_smem)
        for (int i = threadIdx.x; i < p.N2; i += kLnThreads) s_bias2[i] = __ldg(p.bias2 + i);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_res);
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w2);
        tma_prefetch_desc(&tmap_out2);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages1; ++i) {
            mbar_init(&full1[i], 1);
            mbar_init(&empty1[i], 1);
        }
        for (int i = 0; i < kStages2; ++i) {
            mbar_init(&full2[i], 1);
            mbar_init(&empty2[i], 1);
        }
        mbar_init(tmem_full1, 1);
        for (int i = 0; i < 2 * kLnEpiWarps; ++i) mbar_init(&res_bar[i], 1);
        mbar_init(x_ready, kLnEpiWarps);
        mbar_init(x_pair, kEpiArrivals);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], kEpiArrivals);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        if constexpr (kPair) tmem_alloc_2sm<512>(tmem_base_smem);
        else tmem_alloc<512>(tmem_base_smem);
    }
    tc_fence_before();
    if constexpr (kPair) cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / TMA completion
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    if (threadIdx.x == 0) KJ_LGT(0);
    // kPair: the producer thread requests the WEIGHT halves of the first phase-1 stages before it waits for the predecessor grid (they
    // do not depend on it; the activations of the same stages follow after the wait and complete the same barriers)
    constexpr bool kEarlyW = kPair && P1 == 0 && (KJ_LG_EARLY & 1) != 0;
    const int early_w = kEarlyW ? (k_blocks1 < kStages1 ? k_blocks1 : kStages1) : 0;
    if (kEarlyW && threadIdx.x == 0) {
        for (int st = 0; st < early_w; ++st) {
            const uint32_t lbar = mapa_shared(smem_u32(&full1[st]), 0);
            if (leader) mbar_arrive_expect_tx(&full1[st], 2 * (kLnABytes + 2 * kWBoxBytes));
            tma_load_2d_2sm(p1_w0(st), &tmap_w, lbar, st * kGemmBlockK, static_cast<int>(rank) * kWRows, kEvictLast);
            tma_load_2d_2sm(p1_w1(st), &tmap_w, lbar, st * kGemmBlockK, kLnHalfN + static_cast<int>(rank) * kWRows, kEvictLast);
        }
    }
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x == 0) KJ_LGT(1);

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < (P1 == 0 ? k_blocks1 : 0); ++kb) {
                mbar_wait(&empty1[stage], phase ^ 1);
                if constexpr (kPair) {
                    // this CTA's A rows and its 96-row half of each 192-row weight tile; all bytes are counted on the LEADER's barrier
                    const uint32_t lbar = mapa_shared(smem_u32(&full1[stage]), 0);
                    const bool w_requested = kb < early_w;  // weight halves (and the expected byte count) already issued above
                    if (leader && !w_requested) mbar_arrive_expect_tx(&full1[stage], 2 * (kLnABytes + 2 * kWBoxBytes));
                    tma_load_2d_2sm(p1_a(stage), &tmap_a, lbar, kb * kGemmBlockK, tile * kGemmBlockM, kEvictFirst);
                    if (!w_requested) {
                        tma_load_2d_2sm(p1_w0(stage), &tmap_w, lbar, kb * kGemmBlockK, static_cast<int>(rank) * kWRows, kEvictLast);
                        tma_load_2d_2sm(p1_w1(stage), &tmap_w, lbar, kb * kGemmBlockK, kLnHalfN + static_cast<int>(rank) * kWRows, kEvictLast);
                    }
                } else {
                    mbar_arrive_expect_tx(&full1[stage], kLnStageBytes);
                    tma_load_2d(smem_a + stage * kLnABytes, &tmap_a, &full1[stage], kb * kGemmBlockK, tile * kGemmBlockM, kEvictFirst);
                    tma_load_2d(smem_b + stage * kLnBBytes, &tmap_w, &full1[stage], kb * kGemmBlockK, 0, kEvictLast);
                    tma_load_2d(smem_b + stage * kLnBBytes + kLnHalfN * 128, &tmap_w, &full1[stage], kb * kGemmBlockK, kLnHalfN, kEvictLast);
                }
                if (++stage == kStages1) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            // phase 2: W2 tiles.  Stages 0 and 1 alias the phase-1 A ring (free once every phase-1 MMA has retired), stage 2 aliases
            // the LayerNorm staging (free once the LayerNorm epilogue is done): the first two stages are prefetched under the epilogue.
            if (P1 == 0) mbar_wait(tmem_full1, 0);
            int it = 0;
            for (int nb = 0; nb < n2_tiles; ++nb) {
                for (int kb = 0; kb < kLg2KB; ++kb, ++it) {
                    const int s = it % kStages2;
                    if (!kPair && it == kStages2 - 1) mbar_wait(x_ready, 0);  // the last stage aliases the LayerNorm staging (kPair: it has memory of its own)
                    mbar_wait(&empty2[s], ((it / kStages2) & 1) ^ 1);
                    if (it < 48) KJ_LGT(200 + it);
                    if constexpr (kPair) {
                        if (leader) mbar_arrive_expect_tx(&full2[s], 2 * kWBoxBytes);
                        tma_load_2d_2sm(w2_stage(s), &tmap_w2, mapa_shared(smem_u32(&full2[s]), 0), kb * kGemmBlockK,
                                        nb * kLg2BN + static_cast<int>(rank) * kWRows, kEvictLast);
                    } else if (p.dbg & 1) {
                        mbar_arrive(&full2[s]);
                    } else {
                        mbar_arrive_expect_tx(&full2[s], kW2Bytes);
                        tma_load_2d(w2_stage(s), &tmap_w2, &full2[s], kb * kGemmBlockK, nb * kBN2, kEvictLast);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer (kPair: the leader CTA issues for both)
        auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
            if constexpr (kPair) umma_f16_2sm(d, da, db, idesc, accum);
            else umma_f16(d, da, db, idesc, accum);
        };
        auto commit = [&](uint64_t* bar) {
            if constexpr (kPair) umma_commit_2sm(bar, 3);  // the barrier at this offset in both CTAs
            else umma_commit(bar);
        };
        if ((KJ_MMA_UNIFORM != 0 || lane == 0) && leader) {
            constexpr uint32_t idesc = umma_idesc(1 /*bf16*/, kPair ? 2 * kGemmBlockM : kGemmBlockM, kLnHalfN);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < (P1 == 0 ? k_blocks1 : 0); ++kb) {
                mbar_wait(&full1[stage], phase);
                if (kb == 0 && lane == 0) KJ_LGT(2);
                tc_fence_after();
                const uint64_t da = umma_desc_k_sw128(smem_u32(p1_a(stage)));
                const uint64_t db0 = umma_desc_k_sw128(smem_u32(p1_w0(stage)));
                const uint64_t db1 = umma_desc_k_sw128(smem_u32(p1_w1(stage)));
                if (mma_issuer_lane()) {
#pragma unroll
                    for (int k = 0; k < kGemmBlockK / 16; ++k) {
                        mma(tmem_base, da + 2 * k, db0 + 2 * k, idesc, (kb | k) != 0);
                        mma(tmem_base + kLnHalfN, da + 2 * k, db1 + 2 * k, idesc, (kb | k) != 0);
                    }
                    commit(&empty1[stage]);
                    if (kb == k_blocks1 - 1) commit(tmem_full1);
                }
                mma_issuer_sync();
                if (++stage == kStages1) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (lane == 0) KJ_LGT(3);
            // phase 2: A = the resident x' tile, W2 streamed; accumulators [0,192) / [192,384) alternate
            if constexpr (kPair) mbar_wait_cluster(x_pair, 0);  // both CTAs' x' tiles (the peer's arrives are remote)
            else mbar_wait(x_ready, 0);
            tc_fence_after();
            if (lane == 0) KJ_LGT(4);
            int it = 0;
            for (int nb = 0; nb < n2_tiles; ++nb) {
                const int acc = nb & 1;
                if constexpr (kPair) mbar_wait_cluster(&acc_empty[acc], ((nb >> 1) & 1) ^ 1);
                else mbar_wait(&acc_empty[acc], ((nb >> 1) & 1) ^ 1);
                if (lane == 0 && nb < 12) KJ_LGT(16 + 2 * nb);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (kTS ? kLgTAcc0 : 0) + acc * kBN2;
                [[maybe_unused]] long long tw_wait = 0;  // KJ_LG_TRACE: clocks this tile spent waiting for W2 stages
                for (int kb = 0; kb < kLg2KB; ++kb, ++it) {
                    const int s = it % kStages2;
                    [[maybe_unused]] const long long tw0 = KJ_LG_TRACE ? clock64() : 0;
                    mbar_wait(&full2[s], (it / kStages2) & 1);
                    if (KJ_LG_TRACE) tw_wait += clock64() - tw0;
                    tc_fence_after();
                    const uint64_t db = umma_desc_k_sw128(smem_u32(w2_stage(s)));
                    const uint64_t da = umma_desc_k_sw128(smem_u32(smem_x + kb * kLnABytes));
                    // kTS: x' columns [64 kb, 64 kb + 64) as packed bf16 = 32 TMEM columns of column part kb / 2
                    const uint32_t xa = tmem_base + lg_x_tmem_col((kb >> 1)) + (kb & 1) * 32;
                    if (mma_issuer_lane()) {
                        if (p.dbg & 4) {
                        } else if constexpr (kTS) {
                            constexpr uint32_t idesc_ts = umma_idesc(1 /*bf16*/, kGemmBlockM, kLgTBN);
#pragma unroll
                            for (int k = 0; k < kGemmBlockK / 16; ++k) umma_f16_ts(tmem_d, xa + 8 * k, db + 2 * k, idesc_ts, (kb | k) != 0);
                        } else {
#pragma unroll
                            for (int k = 0; k < kGemmBlockK / 16; ++k) mma(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        }
                        commit(&empty2[s]);
                        if (kb == kLg2KB - 1) commit(&acc_full[acc]);
                    }
                    mma_issuer_sync();
                    if (kb == kLg2KB - 1 && lane == 0 && nb < 12) {
                        KJ_LGT(17 + 2 * nb);
                        if (KJ_LG_TRACE && p.trace != nullptr) p.trace[blockIdx.x * 256 + 112 + nb] = p.trace[blockIdx.x * 256] + static_cast<unsigned long long>(tw_wait);  // printed relative to slot 0
                    }
                }
            }
        }
    } else if (warp >= kGemmEpiWarp0) {
        // ---------------------------------------------------------- epilogues
        const int ew = warp - kGemmEpiWarp0;  // 0..11
        const int quad = warp & 3;
        const int part = ew >> 2;
        constexpr int kChunks = kLnPartCols / kEpiChunkCols;  // 4
        // kPair: the packed phase-1 ring ends at 160 KB, so the residual staging has memory of its own -- 8 warps in [160K, 192K) (x' k-blocks 4
        // and 5, written only in pass B, after the statistics barrier that ends every warp's residual reads), 4 warps in the wide-store
        // tiles behind the barriers (phase 2) -- and the first two residual chunks are requested before the accumulator is complete
        uint8_t* ebuf = !kPair ? smem_epi1 + ew * kLnEpiBytesPerWarp
                               : (ew < 8 ? smem + 4 * kP1PairStage + ew * kLnEpiBytesPerWarp : smem + kLg2BaseBytes + (ew - 8) * kLnEpiBytesPerWarp);
        uint64_t* rbar = res_bar + 2 * ew;
        const uint32_t sw = (lane >> 1) & 3;  // 64B swizzle of this lane's row (residual / phase-2 staging tiles)
        const int trow = quad * 32 + lane;    // row inside the tile
        const int row0 = tile * kGemmBlockM + quad * 32;
        uint32_t rphase = 0;
        if constexpr (P1 == 1) {
            // ===== phase 1 (embedding front end): one warp per token, lanes across the 384 columns (coalesced 1536-byte rows, three
            // tokens in flight per warp), the arithmetic of embed_layernorm_kernel / warp_layernorm_store (rowwise.cuh) bit for bit,
            // rows written straight into the swizzled x' operand tile
            const bool has_type = p.type != nullptr && p.type_vocab > 0;
            constexpr int kU = 3;
            for (int r0 = ew; r0 < kGemmBlockM; r0 += kLnEpiWarps * kU) {
                float4 v[kU][3];
                bool live[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int r = r0 + u * kLnEpiWarps;
                    const int row = tile * kGemmBlockM + r;
                    live[u] = r < kGemmBlockM && row < p.M;
                    uint32_t id = 0, tt = 0;
                    bool has_word = false, has_pos = false;
                    int pidx = 0;
                    if (live[u]) {
                        id = __ldg(p.ids + row);
                        has_word = id < static_cast<uint32_t>(p.vocab);   // ids >= vocab contribute a zero row (embeddings/mod.rs:227-246)
                        pidx = p.pos_offset + row % p.S;
                        has_pos = p.pos != nullptr && pidx < p.max_pos;   // positions beyond the table add nothing (:199-214)
                        if (has_type && p.type_ids != nullptr) {
                            tt = __ldg(p.type_ids + row);
                            if (tt >= static_cast<uint32_t>(p.type_vocab)) {  // the reference panics (:312-317)
                                if (lane == 0) *p.err_flag = 1;
                                tt = 0;
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const int c = (lane + 32 * i) * 4;
                        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (live[u]) {
                            if (has_word) a = __ldg(reinterpret_cast<const float4*>(p.word + static_cast<size_t>(id) * kLnN + c));
                            if (has_pos) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(pidx) * kLnN + c));
                                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                            }
                            if (has_type) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.type + static_cast<size_t>(tt) * kLnN + c));
                                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                            }
                        }
                        v[u][i] = a;
                    }
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int r = r0 + u * kLnEpiWarps;
                    if (r >= kGemmBlockM) continue;  // warp-uniform
                    float sm = 0.0f;
#pragma unroll
                    for (int i = 0; i < 3; ++i) sm += (v[u][i].x + v[u][i].y) + (v[u][i].z + v[u][i].w);
                    const float mean = warp_sum(sm) / static_cast<float>(kLnN);
                    float q = 0.0f;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const float a = v[u][i].x - mean, b = v[u][i].y - mean, c = v[u][i].z - mean, d = v[u][i].w - mean;
                        q += (a * a + b * b) + (c * c + d * d);
                    }
                    const float inv_std = 1.0f / sqrtf(warp_sum(q) / static_cast<float>(kLnN) + p.eps);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const int c = (lane + 32 * i) * 4;
                        const float4 g = *reinterpret_cast<const float4*>(s_gamma + c);
                        const float4 bt = *reinterpret_cast<const float4*>(s_beta + c);
                        float y0 = (v[u][i].x - mean) * inv_std * g.x + bt.x;
                        float y1 = (v[u][i].y - mean) * inv_std * g.y + bt.y;
                        float y2 = (v[u][i].z - mean) * inv_std * g.z + bt.z;
                        float y3 = (v[u][i].w - mean) * inv_std * g.w + bt.w;
                        if (!live[u]) y0 = y1 = y2 = y3 = 0.0f;  // rows beyond M: a defined (zero) operand row
                        // x' tile: k-block c / 64, row r (128 B), 16-byte chunk ((c % 64) / 8) ^ (r & 7), 8 bytes at (c % 8) * 2
                        const uint32_t addr = smem_u32(smem_x) + (c >> 6) * kLnABytes + r * 128 + (((static_cast<uint32_t>(c & 63) >> 3) ^ static_cast<uint32_t>(r & 7)) << 4) +
                                              static_cast<uint32_t>(c & 7) * 2;
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(pack_bf16(y0, y1)), "r"(pack_bf16(y2, y3)) : "memory");
                    }
                }
            }
            fence_proxy_async_smem();  // x' (generic-proxy stores) -> visible to the tensor core and to the TMA store below
            __syncwarp();
            if (lane == 0) mbar_arrive(x_ready);
        } else {
            // ===== phase 1: bias + residual + LayerNorm, output into the resident x' tile
            const int col_base = part * kLnPartCols;
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + col_base;
            uint64_t sum2 = f2_pack(0.0f, 0.0f), sq2 = f2_pack(0.0f, 0.0f);
            if constexpr (P1 == 0) {
                auto request_residual = [&] {
                    if (lane == 0) {
                        for (int c = 0; c < 2; ++c) {
                            mbar_arrive_expect_tx(&rbar[c], kEpiStageBytes);
                            tma_load_2d(ebuf + c * kEpiStageBytes, &tmap_res, &rbar[c], col_base + c * kEpiChunkCols, row0, kEvictFirst);
                        }
                    }
                    __syncwarp();
                };
                if constexpr (kPair && (KJ_LG_EARLY & 2) != 0) request_residual();  // staging of its own: in flight while phase 1 runs
                mbar_wait(tmem_full1, 0);
                if (ew == 0 && lane == 0) KJ_LGT(5);
                tc_fence_after();
                if constexpr (!(kPair && (KJ_LG_EARLY & 2) != 0)) request_residual();  // staging aliased on the ring: free once every phase-1 MMA has retired
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
                    __syncwarp();
                    if (lane == 0 && c + 2 < kChunks) {
                        mbar_arrive_expect_tx(&rbar[b], kEpiStageBytes);
                        tma_load_2d(ebuf + b * kEpiStageBytes, &tmap_res, &rbar[b], col_base + (c + 2) * kEpiChunkCols, row0, kEvictFirst);
                    }
                }
            }
            tmem_st_wait();
            if (ew == 0 && lane == 0) KJ_LGT(6);
            float s1, s2;
            {
                float e0, e1, q0, q1;
                f2_unpack(sum2, e0, e1);
                f2_unpack(sq2, q0, q1);
                s1 = e0 + e1;
                s2 = q0 + q1;
            }
            stat[part * 128 + trow] = make_float2(s1, s2);
            named_bar_sync(1, kLnEpiWarps * 32);
            if (ew == 0 && lane == 0) KJ_LGT(7);
            float t1 = 0.0f, t2 = 0.0f;
#pragma unroll
            for (int q = 0; q < kLnParts; ++q) {
                const float2 t = stat[q * 128 + trow];
                t1 += t.x;
                t2 += t.y;
            }
            const float mean = t1 * (1.0f / kLnN);
            const float var = fmaxf(t2 * (1.0f / kLnN) - mean * mean, 0.0f);
            const float rstd = 1.0f / sqrtf(var + p.eps);
            const float nmr = -mean * rstd;
            // pass B: normalise -> bf16 -> the x' tile in the K-major 128B-swizzled layout (k-block = 64 columns, row = 128 B,
            // 16-byte chunk index ^= row & 7): exactly what a TMA load of x' with the A-operand tensor map would have produced
            const uint32_t xsw = static_cast<uint32_t>(trow & 7);
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(taddr0 + c * kEpiChunkCols, v);
                tmem_ld_wait();
                const int col0 = col_base + c * kEpiChunkCols;
                uint32_t pk[16];
                ln_pass_b_chunk(v, rstd, nmr, s_gamma + col0, s_beta + col0, pk);
                const uint32_t xrow = smem_u32(smem_x) + (col0 >> 6) * kLnABytes + trow * 128;
                const uint32_t chunk0 = static_cast<uint32_t>((col0 & 63) >> 3);  // 0 or 4
#pragma unroll
                for (int j = 0; j < 4; ++j) st_shared_v4(xrow + (((chunk0 + j) ^ xsw) << 4), pk[4 * j + 0], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                if constexpr (kTS) {
                    // the phase-2 A operand: row = lane, two bf16 per 32-bit column.  Part 0 compacts in place (columns [16c, 16c + 16)
                    // lie inside what this warp has already read), parts 1 and 2 go to the 128 columns phase 1 never used.
                    tmem_st_32x16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + lg_x_tmem_col(part) + c * 16, pk);
                }
            }
            if constexpr (kTS) tmem_st_wait();
            fence_proxy_async_smem();  // x' (generic-proxy stores) -> visible to the tensor core and to the TMA store below
            tc_fence_before();         // the LayerNorm accumulator is fully read: phase-2 MMAs may overwrite it
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(x_ready);
                if constexpr (kPair) mbar_arrive_cluster(mapa_shared(smem_u32(x_pair), 0));  // the leader's MMA warp waits for both tiles
                if (ew == 0) KJ_LGT(8);
            }
        }
        // x' goes to global memory as well (residual of the next LayerNorm, pooling / hidden-state output): six 16 KB stores
        // straight out of the operand tile; rows >= M are clipped by the tensor map
        if (ew == 0 && lane == 0) {
            mbar_wait(x_ready, 0);
            for (int kb = 0; kb < kLg2KB; ++kb) tma_store_2d(&tmap_x, smem_x + kb * kLnABytes, kb * kGemmBlockK, tile * kGemmBlockM);
            bulk_commit();
        }
        {
            // ===== phase 2: bias (+ activation) -> bf16 -> 32 x 32 staging chunk -> TMA store, per 192- (kTS: 128-) column tile
            uint8_t* sbuf = smem_epi2 + ew * kEpiStageBytes;
            mbar_wait(x_ready, 0);  // the staging chunk aliases the LayerNorm staging of OTHER warps
            for (int nb = 0; nb < n2_tiles; ++nb) {
                const int acc = nb & 1;
                mbar_wait(&acc_full[acc], (nb >> 1) & 1);
                const bool tr = KJ_LG_TRACE && lane == 0 && nb < 8 && (ew == 0 || ew == 7);
                const int ts0 = (ew == 0 ? 48 : 128) + 8 * nb;
                if (tr) KJ_LGT(ts0);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (kTS ? kLgTAcc0 : 0) + acc * kBN2;
                // one 32-column chunk at column `cc` of the tile; `last`: this warp's last load from the accumulator
                auto chunk = [&](int cc, bool last) {
                    uint32_t v[32];
                    tmem_ld_32x32(tacc + cc, v);
                    tmem_ld_wait();
                    if (last) {  // last load landed: release the accumulator before the math and stores of this chunk
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if constexpr (kPair) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&acc_empty[acc]), 0));
                            else mbar_arrive(&acc_empty[acc]);
                        }
                    }
                    const int col0 = nb * kBN2 + cc;
                    if (col0 < p.N2 && !(p.dbg & 2)) {
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                        if (bias2_in_smem && col0 + 32 <= p.N2) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                add_bias4(f + 4 * j, *reinterpret_cast<const float4*>(s_bias2 + col0 + 4 * j));
                            }
                        } else if (p.bias2 != nullptr) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                if (col0 + 4 * j < p.N2) {
                                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias2 + col0) + j);
                                    f[4 * j + 0] += b.x;
                                    f[4 * j + 1] += b.y;
                                    f[4 * j + 2] += b.z;
                                    f[4 * j + 3] += b.w;
                                }
                            }
                        }
                        if (EPI2 == EPI_BIAS_ACT_BF16) apply_act_tile(f, p.act);
                        if (lane == 0) bulk_wait_read<0>();  // the previous chunk's store (and, for warp 4, the x' stores) has read smem
                        __syncwarp();
                        const uint32_t rbase = smem_u32(sbuf) + lane * 64;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            st_shared_v4(rbase + ((j ^ sw) << 4), pack_bf16(f[8 * j + 0], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                         pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tmap_out2, sbuf, col0, row0);
                            bulk_commit();
                        }
                    }
                };
                if constexpr (!kTS) {
                    // wide store: both chunks of this warp's 32 x 64 part into one 128B-swizzled 4 KB tile, one fence, one TMA store
                    uint8_t* wbuf = ew < 6 ? smem_epi2 + ew * kLg2WideBytes : (ew == 6 ? tail : smem + kLg2BaseBytes + (ew - 7) * kLg2WideBytes);
                    const uint32_t wbase = smem_u32(wbuf) + lane * 128;
                    const uint32_t wsw = static_cast<uint32_t>(lane & 7);  // 128B swizzle: 16-byte chunk index ^= row & 7
                    const int colp = nb * kBN2 + part * 64;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32(tacc + part * 64 + c * 32, v);
                        tmem_ld_wait();
                        if (tr) KJ_LGT(ts0 + 1 + 3 * c);
                        if (c == 1) {  // last load landed: release the accumulator before the math and stores of this chunk
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) {
                                if constexpr (kPair) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&acc_empty[acc]), 0));
                                else mbar_arrive(&acc_empty[acc]);
                            }
                        }
                        const int col0 = colp + c * 32;
                        if (p.dbg & 2) continue;
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                        if (bias2_in_smem && col0 + 32 <= p.N2) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                add_bias4(f + 4 * j, *reinterpret_cast<const float4*>(s_bias2 + col0 + 4 * j));
                            }
                        } else if (p.bias2 != nullptr) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                if (col0 + 4 * j < p.N2) {
                                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias2 + col0) + j);
                                    f[4 * j + 0] += b.x;
                                    f[4 * j + 1] += b.y;
                                    f[4 * j + 2] += b.z;
                                    f[4 * j + 3] += b.w;
                                }
                            }
                        }
                        if (EPI2 == EPI_BIAS_ACT_BF16) apply_act_tile(f, p.act);
                        if (tr) KJ_LGT(ts0 + 2 + 3 * c);
                        if (c == 0) {
                            if (lane == 0) bulk_wait_read<0>();  // the previous tile's store (and, for warp 4, the x' stores) has read smem
                            __syncwarp();
                            if (tr) KJ_LGT(ts0 + 3);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            st_shared_v4(wbase + ((static_cast<uint32_t>(c * 4 + j) ^ wsw) << 4), pack_bf16(f[8 * j + 0], f[8 * j + 1]),
                                         pack_bf16(f[8 * j + 2], f[8 * j + 3]), pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
                        }
                    }
                    if (!(p.dbg & 2)) {
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (tr) KJ_LGT(ts0 + 6);
                        if (lane == 0 && colp < p.N2) {  // columns >= N2 are clipped by the tensor map
                            tma_store_2d(&tmap_out2, wbuf, colp, row0);
                            bulk_commit();
                        }
                        if (tr) KJ_LGT(ts0 + 7);
                    }
                } else if constexpr (kTS) {
                    // the tile's four chunks are dealt round robin over the quadrant's three warps, continuing across tiles
                    // (chunk c of tile nb belongs to part (4 nb + c) % 3): one or two chunks per warp and tile, four per three tiles
                    const int first = (part + 3 - nb % 3) % 3;
                    if (first == 0) {
                        chunk(0, false);
                        chunk(96, true);
                    } else {
                        chunk(first * 32, true);
                    }
                }
            }
            if (lane == 0) bulk_wait_read<0>();  // shared memory stays valid until the last store has read it
        }
    }

    tc_fence_before();
    if constexpr (kPair) cluster_sync_all();  // peer shared memory / barriers stay valid until both CTAs are done
    else __syncthreads();
    if (threadIdx.x == 0) KJ_LGT(255);
    if (warp == 2) {
        tc_fence_after();
        if constexpr (kPair) tmem_dealloc_2sm<512>(tmem_base);
        else tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace kj
