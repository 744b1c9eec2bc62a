// Encoder handle: model-directory loader (config.json + model.safetensors ->
// layout -> bf16 pre-packed weights in HBM) and the batched forward that strings
// the sm_100a kernels together.  Host-side mirror of
//   EncoderLoader::load_from_pretrained      KT/pipeline/encoder/loader.rs:82-141
//   CpuTransformerEncoder::new / forward     KT/cpu/encoder/transformer_encoder.rs:30-368
//   EncoderLayer::forward_postnorm(_noalloc) KT/cpu/encoder/encoder_layer.rs:113-232
//   get_hidden_states_batch_from_ids         KT/cpu/encoder/traits.rs:66-139
// (KT = kjarni-transformers/src, KM = kjarni-models/src in the reference tree).
#include <cuda.h>

#include <algorithm>
#include <mutex>

#include "attention.cuh"
#include "attention_tc.cuh"
#include "attention_ts.cuh"
#include "encoder.hpp"
#include "gemm_ln.cuh"
#include "gemm_ln_gemm.cuh"
#include "gemm_tcgen05.cuh"
#include "rowwise.cuh"
// Two kernels that were measured SLOWER than the default path (a cta_group::2 A-resident pair GEMM and a whole-FFN fusion, DESIGN.md
// section 5) are kept in the tree with their parity tests but compiled only with -DKJ_EXPERIMENTAL_KERNELS (make EXTRA=...).
#ifdef KJ_EXPERIMENTAL_KERNELS
#include "ffn_fused.cuh"
#include "gemm_pair.cuh"
#else
namespace kj {
constexpr int kPairMaxKB = 6, kFfH = 384, kFfChunk = 64;
}
#endif

namespace kj {

// ------------------------------------------------------------ TMA descriptors
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    if (!fn) throw Error(KJC_GPU_UNAVAILABLE, "cuTensorMapEncodeTiled not available from the CUDA driver");
    return fn;
}

// 2-D row-major [rows, cols] tensor, box = [box_rows, 64 elements (128 B)], 128-byte swizzle, zero OOB fill.
CUtensorMap make_tmap_2d(const void* base, CUtensorMapDataType dt, int elem_bytes, uint64_t rows, uint64_t cols, uint32_t box_rows,
                         uint32_t box_cols, int swizzle_bytes) {
    CUtensorMap m;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * static_cast<uint64_t>(elem_bytes)};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15))
        throw Error(KJC_INVALID_CONFIG, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
    CUresult r = get_encode_fn()(&m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(KJC_INFERENCE_FAILED, "cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));
    return m;
}

// 3-D row-major [d2, d1, d0] tensor (d0 contiguous), box = [1, box1, box0].
CUtensorMap make_tmap_3d(const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box1, int swizzle_bytes) {
    CUtensorMap m;
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * static_cast<uint64_t>(elem_bytes), d0 * d1 * static_cast<uint64_t>(elem_bytes)};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15))
        throw Error(KJC_INVALID_CONFIG, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
    CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(KJC_INFERENCE_FAILED, "cuTensorMapEncodeTiled(3d) failed with code " + std::to_string(static_cast<int>(r)));
    return m;
}

// ----------------------------------------------------------------- GEMM launch
int pick_block_n(int N) {
    if (N % 192 == 0) return 192;  // 192-column tiles leave shared memory for the staged bias and a 3-deep store ring
    if (N % 256 == 0) return 256;
    if (N % 128 == 0) return 128;
    if (N <= 64) return 64;
    if (N <= 128) return 128;
    if (N <= 192) return 192;
    return 256;
}

template <int BN, int EPI>
static void launch_gemm_inst(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p, int num_sms, cudaStream_t st) {
    using Cfg = GemmCfg<BN>;
    static int configured[64] = {0};
    auto kern = gemm_tcgen05_kernel<BN, EPI>;
    ensure_smem_attr(kern, Cfg::kSmemBytes, configured);
    const int m_tiles = (p.M + kGemmBlockM - 1) / kGemmBlockM;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int grid = std::min(m_tiles * n_tiles, num_sms);
    launch_pdl(kern, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, ta, tb, tc, p);
}

template <int BN>
static void launch_gemm_bn(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p, int num_sms, cudaStream_t st) {
    switch (epi) {
        case EPI_BIAS_BF16: launch_gemm_inst<BN, EPI_BIAS_BF16>(ta, tb, tc, p, num_sms, st); break;
        case EPI_BIAS_ACT_BF16: launch_gemm_inst<BN, EPI_BIAS_ACT_BF16>(ta, tb, tc, p, num_sms, st); break;
        case EPI_BIAS_RES_F32: launch_gemm_inst<BN, EPI_BIAS_RES_F32>(ta, tb, tc, p, num_sms, st); break;
        case EPI_BIAS_F32: launch_gemm_inst<BN, EPI_BIAS_F32>(ta, tb, tc, p, num_sms, st); break;
        default: throw Error(KJC_INVALID_CONFIG, "unknown GEMM epilogue");
    }
}

// CTA-pair form of the GEMM (gemm_tcgen05_kernel<BN, EPI, true>: tcgen05.mma.cta_group::2 over two row tiles, half a weight tile per CTA):
// bf16 outputs, 192- and 256-column tiles; `tb_half` has a box of block_n / 2 weight rows.
template <int BN, int EPI>
static void launch_gemm_cta_pair_inst(const CUtensorMap& ta, const CUtensorMap& tb_half, const CUtensorMap& tc, const GemmParams& p, int num_sms,
                                      cudaStream_t st) {
    using Cfg = GemmCfg<BN, true>;
    static int configured[64] = {0};
    auto kern = gemm_tcgen05_kernel<BN, EPI, true>;
    ensure_smem_attr(kern, Cfg::kSmemBytes, configured);
    const int m_tiles = (p.M + kGemmBlockM - 1) / kGemmBlockM;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int grid = 2 * std::min(((m_tiles + 1) / 2) * n_tiles, std::max(1, num_sms / 2));
    launch_pdl_cluster(2, kern, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, ta, tb_half, tc, p);
}
bool gemm_cta_pair_supported(int block_n, int epi) { return (block_n == 192 || block_n == 256) && (epi == EPI_BIAS_BF16 || epi == EPI_BIAS_ACT_BF16); }
void launch_gemm_cta_pair(int block_n, int epi, const CUtensorMap& ta, const CUtensorMap& tb_half, const CUtensorMap& tc, const GemmParams& p,
                          int num_sms, cudaStream_t st) {
    if (p.N % 16 != 0 || p.K % 8 != 0) throw Error(KJC_INVALID_CONFIG, "GEMM needs N % 16 == 0 and K % 8 == 0");
    if (!gemm_cta_pair_supported(block_n, epi)) throw Error(KJC_INVALID_CONFIG, "CTA-pair GEMM: 192- or 256-column tiles, bf16 output");
    const bool act = epi == EPI_BIAS_ACT_BF16;
    if (block_n == 192) act ? launch_gemm_cta_pair_inst<192, EPI_BIAS_ACT_BF16>(ta, tb_half, tc, p, num_sms, st) : launch_gemm_cta_pair_inst<192, EPI_BIAS_BF16>(ta, tb_half, tc, p, num_sms, st);
    else act ? launch_gemm_cta_pair_inst<256, EPI_BIAS_ACT_BF16>(ta, tb_half, tc, p, num_sms, st) : launch_gemm_cta_pair_inst<256, EPI_BIAS_BF16>(ta, tb_half, tc, p, num_sms, st);
}

void launch_gemm(int block_n, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p, int num_sms,
                 cudaStream_t st) {
    if (p.N % 16 != 0 || p.K % 8 != 0) throw Error(KJC_INVALID_CONFIG, "GEMM needs N % 16 == 0 and K % 8 == 0");
    switch (block_n) {
        case 64: launch_gemm_bn<64>(epi, ta, tb, tc, p, num_sms, st); break;
        case 128: launch_gemm_bn<128>(epi, ta, tb, tc, p, num_sms, st); break;
        case 192: launch_gemm_bn<192>(epi, ta, tb, tc, p, num_sms, st); break;
        case 256: launch_gemm_bn<256>(epi, ta, tb, tc, p, num_sms, st); break;
        default: throw Error(KJC_INVALID_CONFIG, "unsupported GEMM block N");
    }
}

#ifdef KJ_EXPERIMENTAL_KERNELS
// CTA-pair kernel (gemm_pair.cuh): K <= 384, bf16 output; `tb_half` has a box of block_n/2 weight rows.
template <int BN, int EPI>
static void launch_gemm_pair_inst(const CUtensorMap& ta, const CUtensorMap& tb_half, const CUtensorMap& tc, const GemmParams& p, int num_sms,
                                  cudaStream_t st) {
    using Cfg = PairCfg<BN>;
    static int configured[64] = {0};
    auto kern = gemm_pair_kernel<BN, EPI>;
    ensure_smem_attr(kern, Cfg::kSmemBytes, configured);
    const int m_tiles = (p.M + 2 * kGemmBlockM - 1) / (2 * kGemmBlockM);
    const int grid = 2 * std::min(m_tiles, num_sms / 2);
    launch_pdl(kern, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, ta, tb_half, tc, p);
}

#endif
void launch_gemm_pair(int block_n, int epi, const CUtensorMap& ta, const CUtensorMap& tb_half, const CUtensorMap& tc, const GemmParams& p,
                      int num_sms, cudaStream_t st) {
#ifndef KJ_EXPERIMENTAL_KERNELS
    throw Error(KJC_INVALID_CONFIG, "the CTA-pair GEMM is an experimental kernel: rebuild with -DKJ_EXPERIMENTAL_KERNELS");
#else
    if (p.N % 16 != 0 || p.K % 8 != 0 || p.K > kPairMaxKB * kGemmBlockK) throw Error(KJC_INVALID_CONFIG, "pair GEMM needs N % 16 == 0, K % 8 == 0, K <= 384");
    if (epi != EPI_BIAS_BF16 && epi != EPI_BIAS_ACT_BF16) throw Error(KJC_INVALID_CONFIG, "pair GEMM stores bf16 only");
    const bool act = epi == EPI_BIAS_ACT_BF16;
    switch (block_n) {
        case 128: act ? launch_gemm_pair_inst<128, EPI_BIAS_ACT_BF16>(ta, tb_half, tc, p, num_sms, st) : launch_gemm_pair_inst<128, EPI_BIAS_BF16>(ta, tb_half, tc, p, num_sms, st); break;
        case 192: act ? launch_gemm_pair_inst<192, EPI_BIAS_ACT_BF16>(ta, tb_half, tc, p, num_sms, st) : launch_gemm_pair_inst<192, EPI_BIAS_BF16>(ta, tb_half, tc, p, num_sms, st); break;
        case 256: act ? launch_gemm_pair_inst<256, EPI_BIAS_ACT_BF16>(ta, tb_half, tc, p, num_sms, st) : launch_gemm_pair_inst<256, EPI_BIAS_BF16>(ta, tb_half, tc, p, num_sms, st); break;
        default: throw Error(KJC_INVALID_CONFIG, "unsupported pair GEMM block N");
    }
#endif
}

void launch_gemm_ln(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& t_io, int M, int H, int K, const float* bias, const float* gamma,
                    const float* beta, float eps, int num_sms, cudaStream_t st) {
    static int configured[64] = {0}, configured2[64] = {0};
    if (K % 8 != 0) throw Error(KJC_INVALID_CONFIG, "GEMM needs K % 8 == 0");
    GemmLnParams p;
    p.M = M; p.K = K; p.bias = bias; p.gamma = gamma; p.beta = beta; p.eps = eps;
    const int m_tiles = (M + kGemmBlockM - 1) / kGemmBlockM;
    if (H == kLnN) {
        ensure_smem_attr(gemm_ln_kernel<1>, kLnSmemBytes, configured);
        launch_pdl(gemm_ln_kernel<1>, dim3(std::min(m_tiles, num_sms)), dim3(kLnThreads), kLnSmemBytes, st, ta, tw, t_io, t_io, p);
    } else if (H == 2 * kLnN) {  // CTA pair per row tile: each CTA owns 384 of the 768 columns, row statistics exchanged through DSMEM
        ensure_smem_attr(gemm_ln_kernel<2>, kLnSmemBytes, configured2);
        launch_pdl_cluster(2, gemm_ln_kernel<2>, dim3(2 * std::min(m_tiles, std::max(1, num_sms / 2))), dim3(kLnThreads), kLnSmemBytes, st, ta, tw, t_io,
                           t_io, p);
    } else {
        throw Error(KJC_INVALID_CONFIG, "fused GEMM + LayerNorm supports hidden sizes 384 and 768");
    }
}

unsigned long long* g_lg_trace = nullptr;  // KJ_LG_TRACE builds: stamp buffer of the next chained launch (dbg_gemm_ln_gemm)
// GEMM + residual + LayerNorm chained with the next projection of the same 128-row tiles (gemm_ln_gemm.cuh); one tile per CTA.
// pair: two CTAs per cluster share every weight tile (tcgen05.mma.cta_group::2); tw / tw2 must then have 96-row boxes.
void launch_gemm_ln_gemm(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& t_res, const CUtensorMap& t_x, const CUtensorMap& tw2,
                         const CUtensorMap& t_out2, int M, int K1, const float* bias1, const float* gamma, const float* beta, float eps, int N2,
                         const float* bias2, int epi2, int act, cudaStream_t st, bool pair = false, bool ts = false, int dbg = 0, int tile_base = 0,
                         int tile_count = -1) {
    static int configured_act[64] = {0}, configured_plain[64] = {0}, configured_act2[64] = {0}, configured_plain2[64] = {0};  // per instantiation
    static int configured_act_ts[64] = {0}, configured_plain_ts[64] = {0};
    if (K1 % 8 != 0 || N2 % 8 != 0) throw Error(KJC_INVALID_CONFIG, "GEMM needs K % 8 == 0 and N % 8 == 0");
    GemmLnGemmParams p{};
    p.M = M; p.K1 = K1; p.bias1 = bias1; p.gamma = gamma; p.beta = beta; p.eps = eps; p.N2 = N2; p.bias2 = bias2; p.act = act; p.dbg = dbg; p.trace = g_lg_trace;
    p.tile_base = tile_base;
    const int m_tiles = tile_count >= 0 ? tile_count : (M + kGemmBlockM - 1) / kGemmBlockM;  // tiles of this launch
    if (pair) {
        const int ctas = 2 * ((m_tiles + 1) / 2);  // whole clusters; a tile beyond M is all padding (loads zero-filled, stores clipped)
        if (epi2 == EPI_BIAS_ACT_BF16) {
            auto kern = gemm_ln_gemm_kernel<EPI_BIAS_ACT_BF16, 0, true>;
            ensure_smem_attr(kern, kLg2SmemBytes, configured_act2);
            launch_pdl_cluster(2, kern, dim3(ctas), dim3(kLnThreads), kLg2SmemBytes, st, ta, tw, t_res, t_x, tw2, t_out2, p);
        } else {
            auto kern = gemm_ln_gemm_kernel<EPI_BIAS_BF16, 0, true>;
            ensure_smem_attr(kern, kLg2SmemBytes, configured_plain2);
            launch_pdl_cluster(2, kern, dim3(ctas), dim3(kLnThreads), kLg2SmemBytes, st, ta, tw, t_res, t_x, tw2, t_out2, p);
        }
        return;
    }
    if (ts) {  // x' as the phase-2 A operand in tensor memory; tw2 must have 128-row boxes
        if (epi2 == EPI_BIAS_ACT_BF16) {
            auto kern = gemm_ln_gemm_kernel<EPI_BIAS_ACT_BF16, 0, false, true>;
            ensure_smem_attr(kern, kLgTSmemBytes, configured_act_ts);
            launch_pdl(kern, dim3(m_tiles), dim3(kLnThreads), kLgTSmemBytes, st, ta, tw, t_res, t_x, tw2, t_out2, p);
        } else {
            auto kern = gemm_ln_gemm_kernel<EPI_BIAS_BF16, 0, false, true>;
            ensure_smem_attr(kern, kLgTSmemBytes, configured_plain_ts);
            launch_pdl(kern, dim3(m_tiles), dim3(kLnThreads), kLgTSmemBytes, st, ta, tw, t_res, t_x, tw2, t_out2, p);
        }
        return;
    }
    if (epi2 == EPI_BIAS_ACT_BF16) {
        ensure_smem_attr(gemm_ln_gemm_kernel<EPI_BIAS_ACT_BF16>, kLg2SmemBytes, configured_act);
        launch_pdl(gemm_ln_gemm_kernel<EPI_BIAS_ACT_BF16>, dim3(m_tiles), dim3(kLnThreads), kLg2SmemBytes, st, ta, tw, t_res, t_x, tw2, t_out2, p);
    } else {
        ensure_smem_attr(gemm_ln_gemm_kernel<EPI_BIAS_BF16>, kLg2SmemBytes, configured_plain);
        launch_pdl(gemm_ln_gemm_kernel<EPI_BIAS_BF16>, dim3(m_tiles), dim3(kLnThreads), kLg2SmemBytes, st, ta, tw, t_res, t_x, tw2, t_out2, p);
    }
}

// Embedding gather + embed LayerNorm chained with layer 0's QKV projection (gemm_ln_gemm.cuh, P1 = 1); one tile per CTA
void launch_embed_ln_gemm(const EmbedParams& e, const float* gamma, const float* beta, const CUtensorMap& t_x, const CUtensorMap& tw2,
                          const CUtensorMap& t_out2, int N2, const float* bias2, cudaStream_t st) {
    static int configured[64] = {0};
    GemmLnGemmParams p{};
    p.M = e.M; p.K1 = 0; p.bias1 = nullptr; p.gamma = gamma; p.beta = beta; p.eps = e.eps; p.N2 = N2; p.bias2 = bias2; p.act = ACT_NONE;
    p.ids = e.ids; p.type_ids = e.type_ids; p.word = e.word; p.pos = e.pos; p.type = e.type; p.err_flag = e.err_flag;
    p.S = e.S; p.vocab = e.vocab; p.max_pos = e.max_pos; p.type_vocab = e.type_vocab; p.pos_offset = e.pos_offset;
    const int m_tiles = (e.M + kGemmBlockM - 1) / kGemmBlockM;
    auto kern = gemm_ln_gemm_kernel<EPI_BIAS_BF16, 1>;
    ensure_smem_attr(kern, kLg2SmemBytes, configured);
    launch_pdl(kern, dim3(m_tiles), dim3(kLnThreads), kLg2SmemBytes, st, t_x, t_x, t_x, t_x, tw2, t_out2, p);
}

void launch_ffn_ln(const CUtensorMap& t_x, const CUtensorMap& t_w1, const CUtensorMap& t_w1_pair, const CUtensorMap& t_w2, int M, int I,
                   const float* b1, const float* b2, const float* gamma, const float* beta, float eps, int act, int num_sms, cudaStream_t st) {
#ifndef KJ_EXPERIMENTAL_KERNELS
    throw Error(KJC_INVALID_CONFIG, "the fused feed-forward kernel is experimental: rebuild with -DKJ_EXPERIMENTAL_KERNELS");
#else
    static int configured[64] = {0};
    if (I % kFfChunk != 0 || I <= 0) throw Error(KJC_INVALID_CONFIG, "fused FFN needs an intermediate size that is a multiple of 64");
    static int configured2[64] = {0};
    static const bool pair = getenv("KJC_FFN_NO_PAIR") == nullptr;
    FfnParams p;
    p.M = M; p.I = I; p.b1 = b1; p.b2 = b2; p.gamma = gamma; p.beta = beta; p.eps = eps; p.act = act;
    const int m_tiles = (M + kGemmBlockM - 1) / kGemmBlockM;
    if (pair && m_tiles >= 2 && num_sms >= 2) {  // CTA pairs: tcgen05.mma.cta_group::2, each CTA holds half of every weight tile
        ensure_smem_attr(ffn_ln384_kernel<true>, kFfSmemBytes, configured2);
        const int pairs = std::min((m_tiles + 1) / 2, num_sms / 2);
        launch_pdl_cluster(2, ffn_ln384_kernel<true>, dim3(2 * pairs), dim3(kFfThreads), kFfSmemBytes, st, t_x, t_w1_pair, t_w2, p);
    } else {
        ensure_smem_attr(ffn_ln384_kernel<false>, kFfSmemBytes, configured);
        launch_pdl(ffn_ln384_kernel<false>, dim3(std::min(m_tiles, num_sms)), dim3(kFfThreads), kFfSmemBytes, st, t_x, t_w1, t_w2, p);
    }
#endif
}
bool experimental_kernels_built() {
#ifdef KJ_EXPERIMENTAL_KERNELS
    return true;
#else
    return false;
#endif
}

// ------------------------------------------------------------ row-kernel launch
template <typename F>
static void dispatch_nv(int H, F&& f) {
    const int nv = (H + 127) / 128;
    switch (nv) {
        case 1: f(std::integral_constant<int, 1>()); break;
        case 2: f(std::integral_constant<int, 2>()); break;
        case 3: f(std::integral_constant<int, 3>()); break;
        case 4: f(std::integral_constant<int, 4>()); break;
        case 5: case 6: f(std::integral_constant<int, 6>()); break;
        case 7: case 8: f(std::integral_constant<int, 8>()); break;
        default: throw Error(KJC_INVALID_CONFIG, "hidden size > 1024 is not supported");
    }
}

void launch_layernorm(const float* y, const float* g, const float* b, float eps, float* x32, __nv_bfloat16* x16, int M, int H,
                      cudaStream_t st) {
    const int grid = (M + 7) / 8;
    dispatch_nv(H, [&](auto nv) {
        layernorm_kernel<decltype(nv)::value><<<grid, kRowThreads, 0, st>>>(y, g, b, eps, x32, x16, M, H);
    });
    KJ_CUDA(cudaGetLastError());
}

template <int D>
static void launch_attention_d(const AttnParams& p, cudaStream_t st) {
    static int configured[64] = {0};  // per head_dim instantiation
    const size_t smem = attention_smem_bytes(p.S, D);
    if (smem > 48 * 1024) ensure_smem_attr(attention_kernel<D>, static_cast<int>(smem), configured);
    attention_kernel<D><<<p.B * p.heads, kAttnThreads, smem, st>>>(p);
}

// tcgen05 path: S <= 128, head_dim 32 / 64.  Tensor maps depend on (buffers, B, S, H): one-entry cache per thread.
template <int D>
static void launch_attention_tc(const AttnParams& p, cudaStream_t st) {
    static int configured[64] = {0};
    struct Key { const void *q, *c; int B, S, H; };
    static thread_local Key key{nullptr, nullptr, 0, 0, 0};
    static thread_local CUtensorMap t_qkv, t_ctx;
    if (key.q != p.qkv || key.c != p.ctx || key.B != p.B || key.S != p.S || key.H != p.H) {
        t_qkv = make_tmap_3d(p.qkv, 2, 3 * static_cast<uint64_t>(p.H), p.S, p.B, D, kAtcS, D * 2);
        t_ctx = make_tmap_3d(p.ctx, 2, p.H, p.S, p.B, D, 32, D * 2);
        key = Key{p.qkv, p.ctx, p.B, p.S, p.H};
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    ensure_smem_attr(attention_tc_kernel<D>, AtcCfg<D>::kSmemBytes, configured);
    if (p.max_ctas > 0) sms = std::min(sms, p.max_ctas);
    launch_pdl(attention_tc_kernel<D>, dim3(std::min(p.B * p.heads, sms)), dim3(AtcCfg<D>::kThreads), AtcCfg<D>::kSmemBytes, st, t_qkv, t_ctx, p);
}

// Round-2 kernel (attention_ts.cuh): P in TMEM, S <= 512.  Tensor maps depend on (buffers, B, S, H): one-entry cache per thread.
template <int D, int NKB>
static void launch_attention_ts(const AttnParams& p, cudaStream_t st) {
    using Cfg = AtsCfg<D, NKB>;
    static int configured[64] = {0};
    struct Key { const void *q, *c; int B, S, H, ld; };
    static thread_local Key key{nullptr, nullptr, 0, 0, 0, 0};
    static thread_local CUtensorMap t_qkv, t_ctx;
    const int ld = p.ld_qkv > 0 ? p.ld_qkv : 3 * p.H;
    if (key.q != p.qkv || key.c != p.ctx || key.B != p.B || key.S != p.S || key.H != p.H || key.ld != ld) {
        t_qkv = make_tmap_3d(p.qkv, 2, static_cast<uint64_t>(ld), p.S, p.B, D, 128, D * 2);
        t_ctx = make_tmap_3d(p.ctx, 2, p.H, p.S, p.B, D, 32, D * 2);
        key = Key{p.qkv, p.ctx, p.B, p.S, p.H, ld};
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    ensure_smem_attr(attention_ts_kernel<D, NKB>, Cfg::kSmemBytes, configured);
    if (p.max_ctas > 0) sms = std::min(sms, p.max_ctas);
    launch_pdl(attention_ts_kernel<D, NKB>, dim3(std::min(p.B * p.heads, sms)), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, t_qkv, t_ctx, p);
}
template <int D>
static void launch_attention_ts_d(const AttnParams& p, cudaStream_t st) {
    if (p.S <= 128) launch_attention_ts<D, 1>(p, st);
    else if (p.S <= 256) launch_attention_ts<D, 2>(p, st);
    else launch_attention_ts<D, 4>(p, st);
}

// KJC_ATTN = "ts" (default: attention_ts.cuh), "tc" (round-1 tcgen05 kernel, S <= 128), "legacy" (mma.sync kernel): A/B switch
static int attention_variant() {
    static const int v = [] {
        const char* e = getenv("KJC_ATTN");
        if (getenv("KJC_ATTN_LEGACY") || (e && !strcmp(e, "legacy"))) return 2;
        if (e && !strcmp(e, "tc")) return 1;
        return 0;
    }();
    return v;
}

void launch_attention(const AttnParams& p, int D, cudaStream_t st) {
    const int variant = attention_variant();
    if (variant == 0 && p.S <= 512 && (D == 32 || D == 64) && (p.H % 8 == 0)) {
        if (D == 32) launch_attention_ts_d<32>(p, st);
        else launch_attention_ts_d<64>(p, st);
        KJ_CUDA(cudaGetLastError());
        return;
    }
    if (p.ld_qkv > 0 && p.ld_qkv != 3 * p.H) throw Error(KJC_INVALID_CONFIG, "only attention_ts takes a pitched qkv");
    if (variant <= 1 && p.S <= kAtcS && (D == 32 || D == 64) && (p.H % 8 == 0)) {
        if (D == 32) launch_attention_tc<32>(p, st);
        else launch_attention_tc<64>(p, st);
        KJ_CUDA(cudaGetLastError());
        return;
    }
    switch (D) {
        case 16: launch_attention_d<16>(p, st); break;
        case 32: launch_attention_d<32>(p, st); break;
        case 64: launch_attention_d<64>(p, st); break;
        default: throw Error(KJC_INVALID_CONFIG, "head_dim must be 16, 32 or 64");
    }
    KJ_CUDA(cudaGetLastError());
}

// -------------------------------------------------------------------- loader
static int json_int(const Json& cfg, const char* key, bool required = true, int dflt = 0) {
    const Json* j = cfg.get(key);
    if (!j || j->type != Json::Num) {
        if (required) throw Error(KJC_INVALID_CONFIG, std::string("config.json: missing field '") + key + "'");
        return dflt;
    }
    return static_cast<int>(j->num);
}

static int act_from_string(const std::string& s) {
    // encoder configs map "gelu" -> erf GELU (KM/models/sentence_encoder/configs.rs:194-200; DistilBERT hard-codes it :621)
    if (s == "gelu") return ACT_GELU_ERF;
    if (s == "gelu_new" || s == "gelu_fast" || s == "gelu_pytorch_tanh") return ACT_GELU_TANH;
    if (s == "relu") return ACT_RELU;
    throw Error(KJC_INVALID_CONFIG, "unsupported activation '" + s + "'");
}

Encoder::Encoder(const std::string& dir, int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) throw Error(KJC_GPU_UNAVAILABLE, "no CUDA device available");
    if (device < 0 || device >= ndev) throw Error(KJC_GPU_UNAVAILABLE, "device index out of range");
    cudaDeviceProp prop;
    KJ_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) throw Error(KJC_GPU_UNAVAILABLE, std::string("kjarni-b200 needs an sm_100 GPU, found ") + prop.name);
    num_sms_ = prop.multiProcessorCount;
    info_.device = device;

    struct stat sb;
    if (stat(dir.c_str(), &sb) != 0 || !S_ISDIR(sb.st_mode)) throw Error(KJC_MODEL_NOT_FOUND, "model directory not found: " + dir);
    // ModelWeights::new needs config.json + model.safetensors (KT/weights/model_weights.rs:45-66)
    const std::string cfg_text = read_text_file(dir + "/config.json", KJC_MODEL_NOT_FOUND);
    const Json cfg = JsonParser(cfg_text.data(), cfg_text.size()).parse();
    if (cfg.type != Json::Obj) throw Error(KJC_INVALID_CONFIG, "config.json is not an object");
    SafeTensors st(dir + "/model.safetensors");

    std::string model_type = cfg.string("model_type", "bert");
    for (char& c : model_type) c = static_cast<char>(tolower(static_cast<unsigned char>(c)));  // eq_ignore_ascii_case, model_weights.rs:197-211
    int pos_offset = 0;
    std::string ep, lp;  // embedding prefix, layer prefix (with "{}" for the index)
    const char *nq, *nk, *nv, *no, *nln1, *nf1, *nf2, *nln2;
    int H, L, heads, max_pos, vocab;
    int act = ACT_GELU_ERF;
    if (model_type == "distilbert") {
        // DistilBertConfig, KM/models/sentence_encoder/configs.rs:473-512,611-687
        info_.arch = KJC_ARCH_DISTILBERT;
        H = json_int(cfg, "dim");
        L = json_int(cfg, "n_layers");
        heads = json_int(cfg, "n_heads");
        max_pos = json_int(cfg, "max_position_embeddings");
        vocab = json_int(cfg, "vocab_size");
        info_.layer_norm_eps = 1e-12f;  // hard-coded, configs.rs:620
        act = ACT_GELU_ERF;             // hard-coded Activation::Gelu, configs.rs:621
        ep = "distilbert.embeddings.";
        lp = "distilbert.transformer.layer.";
        nq = "attention.q_lin"; nk = "attention.k_lin"; nv = "attention.v_lin"; no = "attention.out_lin";
        nln1 = "sa_layer_norm"; nf1 = "ffn.lin1"; nf2 = "ffn.lin2"; nln2 = "output_layer_norm";
    } else if (model_type == "roberta" || model_type == "distilroberta") {
        // RobertaConfig, KM/models/sequence_classifier/configs.rs:147-283 (also what SentenceEncoder::load_config picks for
        // RoBERTa, KM/models/sentence_encoder/model.rs:46-47): `roberta.` prefix, positions start at row 2
        info_.arch = KJC_ARCH_ROBERTA;
        H = json_int(cfg, "hidden_size");
        L = json_int(cfg, "num_hidden_layers");
        heads = json_int(cfg, "num_attention_heads");
        max_pos = json_int(cfg, "max_position_embeddings");
        vocab = json_int(cfg, "vocab_size");
        info_.layer_norm_eps = static_cast<float>(cfg.number("layer_norm_eps", 1e-5));
        const std::string a = cfg.string("hidden_act", "gelu");  // configs.rs:217-222: unknown strings fall back to erf-GELU
        act = a == "gelu_new" ? ACT_GELU_TANH : (a == "relu" ? ACT_RELU : ACT_GELU_ERF);
        pos_offset = 2;  // extra_pos_embeddings, configs.rs:223
        ep = "roberta.embeddings.";
        lp = "roberta.encoder.layer.";
        nq = "attention.self.query"; nk = "attention.self.key"; nv = "attention.self.value"; no = "attention.output.dense";
        nln1 = "attention.output.LayerNorm"; nf1 = "intermediate.dense"; nf2 = "output.dense"; nln2 = "output.LayerNorm";
    } else if (model_type == "mpnet") {
        // MpnetConfig, KM/models/sentence_encoder/configs.rs:370-468: no prefix, attention.attn.{q,k,v,o}, tanh-GELU hard-coded
        // (Activation::GeluNew :408), no token types, positions start at row 2 (:415).  As in the reference, MPNet's relative
        // attention bias tensor is not part of the layout and is ignored.
        info_.arch = KJC_ARCH_MPNET;
        H = json_int(cfg, "hidden_size");
        L = json_int(cfg, "num_hidden_layers");
        heads = json_int(cfg, "num_attention_heads");
        max_pos = json_int(cfg, "max_position_embeddings");
        vocab = json_int(cfg, "vocab_size");
        info_.layer_norm_eps = static_cast<float>(cfg.number("layer_norm_eps", 1e-5));
        act = ACT_GELU_TANH;
        pos_offset = 2;
        ep = "embeddings.";
        lp = "encoder.layer.";
        nq = "attention.attn.q"; nk = "attention.attn.k"; nv = "attention.attn.v"; no = "attention.attn.o";
        nln1 = "attention.LayerNorm"; nf1 = "intermediate.dense"; nf2 = "output.dense"; nln2 = "output.LayerNorm";
    } else {
        // BertConfig, KM/models/sentence_encoder/configs.rs:15-65,174-366
        const bool prefixed = cfg.has("id2label") || cfg.has("num_labels");  // is_hf_classification, configs.rs:92-95
        info_.arch = prefixed ? KJC_ARCH_BERT_PREFIXED : KJC_ARCH_BERT;
        H = json_int(cfg, "hidden_size");
        L = json_int(cfg, "num_hidden_layers");
        heads = json_int(cfg, "num_attention_heads");
        max_pos = json_int(cfg, "max_position_embeddings");
        vocab = json_int(cfg, "vocab_size");
        info_.layer_norm_eps = static_cast<float>(cfg.number("layer_norm_eps", 1e-12));
        act = act_from_string(cfg.string("hidden_act", "gelu"));
        const std::string pre = prefixed ? "bert." : "";
        ep = pre + "embeddings.";
        lp = pre + "encoder.layer.";
        nq = "attention.self.query"; nk = "attention.self.key"; nv = "attention.self.value"; no = "attention.output.dense";
        nln1 = "attention.output.LayerNorm"; nf1 = "intermediate.dense"; nf2 = "output.dense"; nln2 = "output.LayerNorm";
    }
    if (H <= 0 || L <= 0 || heads <= 0 || H % heads != 0) throw Error(KJC_INVALID_CONFIG, "config.json: inconsistent hidden/heads/layers");
    const int d = H / heads;
    if (d != 16 && d != 32 && d != 64) throw Error(KJC_INVALID_CONFIG, "head_dim " + std::to_string(d) + " unsupported (16/32/64)");
    if (H % 16 != 0 || H > 1024) throw Error(KJC_INVALID_CONFIG, "hidden size must be a multiple of 16 and <= 1024");
    info_.hidden_size = H;
    info_.num_layers = L;
    info_.num_heads = heads;
    info_.vocab_size = vocab;
    info_.max_position_embeddings = max_pos;
    cfg_max_seq_len_ = max_pos;  // ModelMetadata::max_seq_len = config.json max_position_embeddings (the tokenizer's truncation length)
    info_.position_offset = pos_offset;
    act_ = act;

    auto shape_is = [&](const std::string& name, std::initializer_list<int64_t> want) {
        const StTensor& t = st.at(name);
        if (t.shape != std::vector<int64_t>(want)) {
            std::string got;
            for (auto v : t.shape) got += std::to_string(v) + " ";
            throw Error(KJC_LOAD_FAILED, "tensor '" + name + "' has shape [" + got + "] (unexpected)");
        }
    };

    // ---- collect fp32 tensors on the host, then one arena upload
    std::vector<float> f32;       // all fp32 parameters
    std::vector<float> w_gemm;    // all GEMM weights (converted to bf16 on device)
    auto push_f32 = [&](const std::vector<float>& v) {
        const size_t off = f32.size();
        f32.insert(f32.end(), v.begin(), v.end());
        while (f32.size() % 64) f32.push_back(0.f);  // 256-byte alignment of every tensor
        return off;
    };
    auto push_w = [&](const std::vector<float>& v) {
        const size_t off = w_gemm.size();
        w_gemm.insert(w_gemm.end(), v.begin(), v.end());
        while (w_gemm.size() % 128) w_gemm.push_back(0.f);
        return off;
    };
    auto opt_bias = [&](const std::string& name, int n, bool& present) {
        present = st.contains(name);
        if (!present) return std::vector<float>(static_cast<size_t>(n), 0.f);
        shape_is(name, {n});
        return st.as_f32(name);
    };

    const std::string word_name = ep + "word_embeddings.weight";
    const StTensor& wt = st.at(word_name);
    if (wt.shape.size() != 2 || wt.shape[1] != H) throw Error(KJC_LOAD_FAILED, "word embedding table has the wrong hidden size");
    info_.vocab_size = static_cast<int>(wt.shape[0]);  // the table is the truth for the bounds check (embeddings/mod.rs:227-246)
    const size_t off_word = push_f32(st.as_f32(word_name));
    const std::string pos_name = ep + "position_embeddings.weight";
    const StTensor& pt = st.at(pos_name);
    if (pt.shape.size() != 2 || pt.shape[1] != H) throw Error(KJC_LOAD_FAILED, "position embedding table has the wrong hidden size");
    info_.max_position_embeddings = static_cast<int>(pt.shape[0]);
    const size_t off_pos = push_f32(st.as_f32(pos_name));
    size_t off_type = 0;
    info_.type_vocab_size = 0;
    const std::string type_name = ep + "token_type_embeddings.weight";
    if (info_.arch != KJC_ARCH_MPNET && st.contains(type_name)) {
        const StTensor& tt = st.at(type_name);
        if (tt.shape.size() != 2 || tt.shape[1] != H) throw Error(KJC_LOAD_FAILED, "token-type table has the wrong hidden size");
        info_.type_vocab_size = static_cast<int>(tt.shape[0]);
        off_type = push_f32(st.as_f32(type_name));
    }
    shape_is(ep + "LayerNorm.weight", {H});
    shape_is(ep + "LayerNorm.bias", {H});
    const size_t off_eg = push_f32(st.as_f32(ep + "LayerNorm.weight"));
    const size_t off_eb = push_f32(st.as_f32(ep + "LayerNorm.bias"));

    struct LayerOff { size_t wqkv, wo, w1, w2, bqkv, bo, b1, b2, g1, be1, g2, be2; };
    std::vector<LayerOff> lo(L);
    int I = -1;
    for (int l = 0; l < L; ++l) {
        const std::string p = lp + std::to_string(l) + ".";
        auto W = [&](const char* n) { return p + n + ".weight"; };
        auto Bn = [&](const char* n) { return p + n + ".bias"; };
        shape_is(W(nq), {H, H}); shape_is(W(nk), {H, H}); shape_is(W(nv), {H, H}); shape_is(W(no), {H, H});
        const StTensor& f1 = st.at(W(nf1));
        if (f1.shape.size() != 2 || f1.shape[1] != H) throw Error(KJC_LOAD_FAILED, "fc1 weight has the wrong shape in layer " + std::to_string(l));
        if (I < 0) I = static_cast<int>(f1.shape[0]);  // intermediate size from the weight, never from metadata
        shape_is(W(nf1), {I, H});
        shape_is(W(nf2), {H, I});
        // fused [3H, H] QKV weight: rows = q | k | v (KT/cpu/encoder/qkv_projection.rs:109-136)
        std::vector<float> wqkv = st.as_f32(W(nq));
        { auto k = st.as_f32(W(nk)); wqkv.insert(wqkv.end(), k.begin(), k.end()); }
        { auto v = st.as_f32(W(nv)); wqkv.insert(wqkv.end(), v.begin(), v.end()); }
        lo[l].wqkv = push_w(wqkv);
        lo[l].wo = push_w(st.as_f32(W(no)));
        lo[l].w1 = push_w(st.as_f32(W(nf1)));
        lo[l].w2 = push_w(st.as_f32(W(nf2)));
        bool hq, hk, hv, hb;
        std::vector<float> bq = opt_bias(Bn(nq), H, hq), bk = opt_bias(Bn(nk), H, hk), bv = opt_bias(Bn(nv), H, hv);
        if (H <= 512 && !(hq && hk && hv)) {
            // the reference's fused-QKV path drops every bias unless all three exist (qkv_projection.rs:235-238)
            std::fill(bq.begin(), bq.end(), 0.f); std::fill(bk.begin(), bk.end(), 0.f); std::fill(bv.begin(), bv.end(), 0.f);
        }
        std::vector<float> bqkv = bq;
        bqkv.insert(bqkv.end(), bk.begin(), bk.end());
        bqkv.insert(bqkv.end(), bv.begin(), bv.end());
        lo[l].bqkv = push_f32(bqkv);
        lo[l].bo = push_f32(opt_bias(Bn(no), H, hb));
        lo[l].b1 = push_f32(opt_bias(Bn(nf1), I, hb));
        lo[l].b2 = push_f32(opt_bias(Bn(nf2), H, hb));
        shape_is(W(nln1), {H}); shape_is(Bn(nln1), {H}); shape_is(W(nln2), {H}); shape_is(Bn(nln2), {H});
        lo[l].g1 = push_f32(st.as_f32(W(nln1)));
        lo[l].be1 = push_f32(st.as_f32(Bn(nln1)));
        lo[l].g2 = push_f32(st.as_f32(W(nln2)));
        lo[l].be2 = push_f32(st.as_f32(Bn(nln2)));
    }
    if (I % 16 != 0) throw Error(KJC_INVALID_CONFIG, "intermediate size must be a multiple of 16");
    info_.intermediate_size = I;

    // ---- classification head, first match wins (KT/cpu/encoder/classifier.rs:113-206)
    size_t off_wpre = 0, off_bpre = 0, off_wcls = 0, off_bcls = 0;
    bool has_bpre = false, has_bcls = false;
    info_.head_kind = KJC_HEAD_ABSENT;
    info_.num_labels = 0;
    std::string pre_w, pre_b, cls_w, cls_b;
    if (st.contains("classifier.dense.weight")) {
        info_.head_kind = KJC_HEAD_DENSE_TANH;
        pre_w = "classifier.dense.weight"; pre_b = "classifier.dense.bias"; cls_w = "classifier.out_proj.weight"; cls_b = "classifier.out_proj.bias";
    } else if (st.contains("pre_classifier.weight")) {
        info_.head_kind = KJC_HEAD_PRE_RELU;
        pre_w = "pre_classifier.weight"; pre_b = "pre_classifier.bias"; cls_w = "classifier.weight"; cls_b = "classifier.bias";
    } else if (st.contains("bert.pooler.dense.weight")) {
        info_.head_kind = KJC_HEAD_POOLER_TANH;
        pre_w = "bert.pooler.dense.weight"; pre_b = "bert.pooler.dense.bias"; cls_w = "classifier.weight"; cls_b = "classifier.bias";
    } else if (st.contains("classifier.weight")) {
        info_.head_kind = KJC_HEAD_LINEAR;
        cls_w = "classifier.weight"; cls_b = "classifier.bias";
    }
    if (info_.head_kind != KJC_HEAD_ABSENT) {
        if (!pre_w.empty()) {
            shape_is(pre_w, {H, H});
            off_wpre = push_f32(st.as_f32(pre_w));
            off_bpre = push_f32(opt_bias(pre_b, H, has_bpre));
        }
        const StTensor& cw = st.at(cls_w);
        if (cw.shape.size() != 2 || cw.shape[1] != H) throw Error(KJC_LOAD_FAILED, "classifier weight has the wrong shape");
        info_.num_labels = static_cast<int>(cw.shape[0]);
        off_wcls = push_f32(st.as_f32(cls_w));
        off_bcls = push_f32(opt_bias(cls_b, info_.num_labels, has_bcls));
    }
    // labels from id2label sorted by numeric key (KF/src/classifier.rs:253-300)
    if (const Json* id2 = cfg.get("id2label")) {
        if (id2->type == Json::Obj) {
            std::vector<std::pair<long, std::string>> items;
            for (auto& kv : id2->obj)
                if (kv.second.type == Json::Str) items.emplace_back(strtol(kv.first.c_str(), nullptr, 10), kv.second.str);
            std::sort(items.begin(), items.end(), [](auto& a, auto& b) { return a.first < b.first; });
            for (auto& it : items) labels_.push_back(it.second);
        }
    }

    // ---- upload
    KJ_CUDA(cudaSetDevice(device));
    KJ_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    KJ_CUDA(cudaMalloc(&d_f32_, f32.size() * sizeof(float)));
    KJ_CUDA(cudaMemcpy(d_f32_, f32.data(), f32.size() * sizeof(float), cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMalloc(&d_w16_, w_gemm.size() * sizeof(__nv_bfloat16)));
    {
        float* tmp = nullptr;
        KJ_CUDA(cudaMalloc(&tmp, w_gemm.size() * sizeof(float)));
        KJ_CUDA(cudaMemcpy(tmp, w_gemm.data(), w_gemm.size() * sizeof(float), cudaMemcpyHostToDevice));
        const size_t n = w_gemm.size();
        f32_to_bf16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream_>>>(tmp, d_w16_, n);
        KJ_CUDA(cudaGetLastError());
        KJ_CUDA(cudaStreamSynchronize(stream_));
        KJ_CUDA(cudaFree(tmp));
    }
    KJ_CUDA(cudaMalloc(&d_err_, sizeof(int)));
    KJ_CUDA(cudaMemset(d_err_, 0, sizeof(int)));

    word_ = d_f32_ + off_word;
    pos_ = d_f32_ + off_pos;
    type_ = info_.type_vocab_size ? d_f32_ + off_type : nullptr;
    emb_g_ = d_f32_ + off_eg;
    emb_b_ = d_f32_ + off_eb;
    bn_qkv_ = pick_block_n(3 * H);
    bn_h_ = pick_block_n(H);
    bn_i_ = pick_block_n(I);
    // FFN-up: 256-column tiles where they divide I.  A tcgen05.mma of M = 128 costs ~128 clk for any N <= 256 (scripts/mma_rate.py),
    // so 1536 = 6 x 256 needs 25 % fewer MMA instructions than 8 x 192; with the 16-warp epilogue the launch is 26.4 us against
    // 28.8 us (profiles/r01_gemm_probes.md).  QKV (1152 = 4.5 x 256) measures the same either way and keeps 192.
    if (I % 256 == 0 && gemm_wide_store(256)) bn_i_ = 256;
    // The same holds for every projection whose width is a whole number of 256-column tiles: a tcgen05.mma of M = 128 retires in
    // N / 2 clk (scripts/ubench/mma_issue.cu) and the issuing warp shares its scheduler with busy epilogue warps, so the widest
    // tile amortises the per-instruction issue cost best.  Hidden 768: Q|K|V = 9 x 256, out-proj / FFN-down = 3 x 256.
    if ((3 * H) % 256 == 0 && gemm_wide_store(256)) bn_qkv_ = 256;
    if (H % 256 == 0 && gemm_wide_store(256)) bn_h_ = 256;
    if (const char* e = getenv("KJC_BN_I")) bn_i_ = atoi(e);  // tuning hooks
    if (const char* e = getenv("KJC_BN_QKV")) bn_qkv_ = atoi(e);
    layers_.resize(L);
    for (int l = 0; l < L; ++l) {
        LayerDev& ld = layers_[l];
        ld.wqkv = d_w16_ + lo[l].wqkv; ld.wo = d_w16_ + lo[l].wo; ld.w1 = d_w16_ + lo[l].w1; ld.w2 = d_w16_ + lo[l].w2;
        ld.bqkv = d_f32_ + lo[l].bqkv; ld.bo = d_f32_ + lo[l].bo; ld.b1 = d_f32_ + lo[l].b1; ld.b2 = d_f32_ + lo[l].b2;
        ld.g1 = d_f32_ + lo[l].g1; ld.be1 = d_f32_ + lo[l].be1; ld.g2 = d_f32_ + lo[l].g2; ld.be2 = d_f32_ + lo[l].be2;
        ld.t_wqkv = make_tmap_2d(ld.wqkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3 * H, H, bn_qkv_, kGemmBlockK, 128);
        ld.t_wo = make_tmap_2d(ld.wo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, H, H, bn_h_, kGemmBlockK, 128);
        ld.t_w1 = make_tmap_2d(ld.w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, I, H, bn_i_, kGemmBlockK, 128);
        ld.t_w2 = make_tmap_2d(ld.w2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, H, I, bn_h_, kGemmBlockK, 128);
        ld.t_wo_ln = make_tmap_2d(ld.wo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, H, H, kLnHalfN, kGemmBlockK, 128);  // GEMM + LN kernels: 192-row boxes
        ld.t_w2_ln = make_tmap_2d(ld.w2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, H, I, kLnHalfN, kGemmBlockK, 128);
        if (H == kLnN) {  // CTA-pair chained kernels: each CTA loads 96 of the 192 rows of a weight tile
            ld.t_wo_96 = make_tmap_2d(ld.wo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, H, H, kLnHalfN / 2, kGemmBlockK, 128);
            ld.t_w2_96 = make_tmap_2d(ld.w2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, H, I, kLnHalfN / 2, kGemmBlockK, 128);
            ld.t_w1_96 = make_tmap_2d(ld.w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, I, H, kLg2BN / 2, kGemmBlockK, 128);
            ld.t_wqkv_96 = make_tmap_2d(ld.wqkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3 * H, H, kLg2BN / 2, kGemmBlockK, 128);
        }
        ld.t_w1_192 = make_tmap_2d(ld.w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, I, H, kLg2BN, kGemmBlockK, 128);      // chained kernels: 192-row boxes
        ld.t_wqkv_192 = make_tmap_2d(ld.wqkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3 * H, H, kLg2BN, kGemmBlockK, 128);
        ld.t_w1_128 = make_tmap_2d(ld.w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, I, H, kLgTBN, kGemmBlockK, 128);      // x' in tensor memory: 128-row boxes
        ld.t_wqkv_128 = make_tmap_2d(ld.wqkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3 * H, H, kLgTBN, kGemmBlockK, 128);
        if (H == kFfH && I % kFfChunk == 0) {
            ld.t_w1_ffn = make_tmap_2d(ld.w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, I, H, 64, kGemmBlockK, 128);
            ld.t_w1_ffn32 = make_tmap_2d(ld.w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, I, H, 32, kGemmBlockK, 128);
            ld.t_w2_ffn = make_tmap_2d(ld.w2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, H, I, 64, kGemmBlockK, 128);
        }
        if (bn_qkv_ >= 128 && bn_i_ >= 128) {  // CTA-pair kernels: each CTA loads half of a weight tile
            ld.t_wqkv_half = make_tmap_2d(ld.wqkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3 * H, H, bn_qkv_ / 2, kGemmBlockK, 128);
            ld.t_w1_half = make_tmap_2d(ld.w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, I, H, bn_i_ / 2, kGemmBlockK, 128);
        }
    }
    {
        // QKV and FFN-up as CTA pairs (gemm_tcgen05_kernel<BN, EPI, true>): bit 0 = QKV, bit 1 = FFN-up; KJC_GEMM_PAIR overrides
        const char* e = getenv("KJC_GEMM_PAIR");
        const int want = e != nullptr ? atoi(e) : KJ_GEMM_PAIR_DEFAULT;
        gemm_pair_mask_ = num_sms_ % 2 == 0 ? ((gemm_cta_pair_supported(bn_qkv_, EPI_BIAS_BF16) ? want & 1 : 0) | (gemm_cta_pair_supported(bn_i_, EPI_BIAS_ACT_BF16) ? want & 2 : 0)) : 0;
    }
    pair_gemm_ = H <= kPairMaxKB * kGemmBlockK && H % kGemmBlockK == 0 && bn_qkv_ >= 128 && bn_i_ >= 128 && experimental_kernels_built() && getenv("KJC_PAIR_GEMM") != nullptr;  // slower than the 1-CTA kernel at one 256-row tile per pair (launch-bound regime): opt-in
    if (info_.head_kind != KJC_HEAD_ABSENT) {
        w_pre_ = pre_w.empty() ? nullptr : d_f32_ + off_wpre;
        b_pre_ = (pre_w.empty() || !has_bpre) ? nullptr : d_f32_ + off_bpre;
        w_cls_ = d_f32_ + off_wcls;
        b_cls_ = has_bcls ? d_f32_ + off_bcls : nullptr;
    }
    // hidden 384: out-proj / FFN-down run as one GEMM + bias + residual + LayerNorm kernel (full rows per CTA)
    // hidden 768: the same kernel on a CTA pair per row tile (row statistics exchanged through distributed shared memory)
    fused_ln_ = (H == kLnN || H == 2 * kLnN) && !getenv("KJC_NO_FUSED_LN");
    // whole-FFN fusion (ffn_fused.cuh) is correct but shared-memory-bandwidth-bound (the 128 x 384 x tile is re-read for every 64
    // intermediate columns): 64-75 us per launch against 35 + 33 us for the two-kernel path, so it is opt-in
    fused_ffn_ = fused_ln_ && H == kFfH && I % kFfChunk == 0 && experimental_kernels_built() && getenv("KJC_FUSED_FFN") != nullptr;
    // out-proj + LN1 -> FFN-up and FFN-down + LN2 -> next layer's QKV as one launch each (gemm_ln_gemm.cuh)
    chain_ = fused_ln_ && H == kLnN && !fused_ffn_ && I <= kLg2BiasMax && 3 * H <= kLg2BiasMax && !getenv("KJC_NO_CHAIN");
    chain_embed_ = chain_ && getenv("KJC_CHAIN_EMBED") != nullptr;
    // two CTAs per cluster share the weight tiles of the chained kernels (cta_group::2): needs an even number of CTAs resident
    {
        // CTA-pair form of the chained launches (tcgen05.mma.cta_group::2, half a weight tile per CTA): bit 0 = FFN-down + LN2 -> QKV,
        // bit 1 = out-proj + LN1 -> FFN-up
        const char* e = getenv("KJC_CHAIN_PAIR");
        chain_pair_mask_ = (chain_ && num_sms_ % 2 == 0) ? (e != nullptr ? atoi(e) : KJ_CHAIN_PAIR_DEFAULT) & 3 : 0;
        chain_pair_ = chain_pair_mask_ != 0;
        last_ln_pair_ = getenv("KJC_LAST_LN_PAIR") != nullptr;
    }
    {
        const char* e = getenv("KJC_CHAIN_TS");  // phase 2 reads x' from tensor memory (gemm_ln_gemm.cuh, kTS)
        chain_ts_ = chain_ && (e == nullptr ? KJ_CHAIN_TS_DEFAULT != 0 : atoi(e) != 0);
    }
    if (const char* e = getenv("KJC_FP32_RESIDUAL")) set_fp32_residual(atoi(e));
    alias_qkv_ = getenv("KJC_NO_ALIAS_QKV") == nullptr;
    chain_min_tiles_ = num_sms_ / 2 + 1;
    if (const char* e = getenv("KJC_CHAIN_MIN_TILES")) chain_min_tiles_ = atoi(e);
    const char* env = getenv("KJC_MICRO_TOKENS");
    micro_tokens_ = env ? std::max(128, atoi(env)) : num_sms_ * 128;
    lanes_ = 1;  // measured: no gain from concurrent lanes (the kernels are epilogue-issue-bound, not launch-latency-bound)
    if (const char* e = getenv("KJC_LANES")) lanes_ = std::min(8, std::max(1, atoi(e)));
    ws_.resize(lanes_);
    for (Workspace& w : ws_) {
        KJ_CUDA(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
        KJ_CUDA(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
    }
    KJ_CUDA(cudaEventCreateWithFlags(&ev_in_, cudaEventDisableTiming));
}

Encoder::~Encoder() {
    cudaSetDevice(info_.device);
    for (Workspace& w : ws_) {
        free_workspace(w);
        if (w.stream) cudaStreamDestroy(w.stream);
        if (w.done) cudaEventDestroy(w.done);
    }
    if (ev_in_) cudaEventDestroy(ev_in_);
    for (auto* v : {&ev_in_chunk_, &ev_done_chunk_, &ev_out_chunk_})
        for (cudaEvent_t e : *v) cudaEventDestroy(e);
    if (s_in_) cudaStreamDestroy(s_in_);
    if (s_out_) cudaStreamDestroy(s_out_);
    if (d_f32_) cudaFree(d_f32_);
    if (d_w16_) cudaFree(d_w16_);
    if (d_err_) cudaFree(d_err_);
    for (ProfRec& r : prof_recs_) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (cudaEvent_t e : prof_pool_) cudaEventDestroy(e);
    if (d_in_) cudaFree(d_in_);
    if (d_out_) cudaFree(d_out_);
    if (h_stage_in_) cudaFreeHost(h_stage_in_);
    if (h_stage_out_) cudaFreeHost(h_stage_out_);
    if (stream_) cudaStreamDestroy(stream_);
}

void Encoder::free_workspace(Workspace& w) {
    if (w.qkv16 == w.h16) w.qkv16 = nullptr;  // aliased (ensure_workspace)
    for (void* p : {(void*)w.y32, (void*)w.x32, (void*)w.x16, (void*)w.qkv16, (void*)w.ctx16, (void*)w.h16, (void*)w.head32})
        if (p) cudaFree(p);
    w.y32 = w.x32 = nullptr; w.x16 = w.qkv16 = w.ctx16 = w.h16 = nullptr;
    w.head32 = nullptr;
    w.tokens = 0;
}

int Encoder::micro_batch(int S) const { return std::max(1, micro_tokens_ / std::max(S, 1)); }

// Activations for one micro-batch; sized once for the largest token count seen (>= 128 rows so TMA boxes fit).
void Encoder::ensure_workspace(Workspace& w, int tokens) {
    if (tokens <= w.tokens) return;
    KJ_CUDA(cudaStreamSynchronize(w.stream));
    free_workspace(w);
    const size_t T = static_cast<size_t>(std::max(tokens, 128));
    const int H = info_.hidden_size, I = info_.intermediate_size;
    if (!fused_ln_) KJ_CUDA(cudaMalloc(&w.y32, T * H * 4));  // pre-LayerNorm sums of the unfused path (fp32-residual mode allocates lazily)
    KJ_CUDA(cudaMalloc(&w.x16, T * H * 2));
    KJ_CUDA(cudaMalloc(&w.ctx16, T * H * 2));
    KJ_CUDA(cudaMalloc(&w.h16, T * I * 2));
    // Q|K|V of a row are dead once attention has run and the FFN activations of a row are dead once FFN-down has read them, and every
    // kernel that writes one of them for a row tile has finished reading the other for that tile (the chained FFN-down launch stores QKV
    // only after all phase-1 MMAs -- the readers of h -- have retired).  So row r of qkv lives in the first 3H elements of row r of h
    // (pitch I): a layer's working set drops from 131 MB to 87 MB per 148 sequences and the dead tensor is overwritten in L2 instead of
    // being written back to HBM (ncu --cache-control none: 138 MB of DRAM writes per layer before).  Needs the pitched attention_ts.
    {
        const int d = H / std::max(1, info_.num_heads);
        const bool ts_attention = attention_variant() == 0 && (d == 32 || d == 64) && H % 8 == 0;
        if (alias_qkv_ && ts_attention && I >= 3 * H) {
            w.qkv16 = w.h16;
            w.qkv_ld = I;
        } else {
            KJ_CUDA(cudaMalloc(&w.qkv16, T * 3 * H * 2));
            w.qkv_ld = 3 * H;
        }
    }
    if (w_pre_) KJ_CUDA(cudaMalloc(&w.head32, T * H * 4));  // pre-classifier output, at most one sequence per token
    // stale rows beyond the live token count are read by TMA (results discarded): keep them finite
    KJ_CUDA(cudaMemsetAsync(w.x16, 0, T * H * 2, w.stream));
    KJ_CUDA(cudaMemsetAsync(w.ctx16, 0, T * H * 2, w.stream));
    KJ_CUDA(cudaMemsetAsync(w.h16, 0, T * I * 2, w.stream));
    KJ_CUDA(cudaStreamSynchronize(w.stream));  // the forward may run on a caller stream
    w.t_x16 = make_tmap_2d(w.x16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, T, H, kGemmBlockM, kGemmBlockK, 128);
    w.t_ctx16 = make_tmap_2d(w.ctx16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, T, H, kGemmBlockM, kGemmBlockK, 128);
    w.t_h16 = make_tmap_2d(w.h16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, T, I, kGemmBlockM, kGemmBlockK, 128);
    // store boxes: 192-column tiles store 32 x 64 parts (128B swizzle), the other widths 32 x 32 chunks (64B swizzle)
    w.t_qkv16_out = (gemm_wide_store(bn_qkv_) || (bn_qkv_ == 192 && pair_gemm_)) ? make_tmap_2d(w.qkv16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, T, w.qkv_ld, 32, 64, 128)
                                   : make_tmap_2d(w.qkv16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, T, w.qkv_ld, 32, kEpiChunkCols, 64);
    w.t_qkv16_out32 = make_tmap_2d(w.qkv16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, T, w.qkv_ld, 32, kEpiChunkCols, 64);  // CTA-pair kernel
    w.t_h16_out32 = make_tmap_2d(w.h16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, T, I, 32, kEpiChunkCols, 64);
    w.t_qkv16_out64 = make_tmap_2d(w.qkv16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, T, w.qkv_ld, 32, 64, 128);  // chained kernels: one 32 x 64 store per warp and tile
    w.t_h16_out64 = make_tmap_2d(w.h16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, T, I, 32, 64, 128);
    w.t_h16_out = (gemm_wide_store(bn_i_) || (bn_i_ == 192 && pair_gemm_)) ? make_tmap_2d(w.h16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, T, I, 32, 64, 128)
                               : make_tmap_2d(w.h16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, T, I, 32, kEpiChunkCols, 64);
    w.t_x16_io = make_tmap_2d(w.x16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, T, H, 32, kEpiChunkCols, 64);
    w.tokens = static_cast<int>(T);
}

// fp32-residual mode: the fp32 copy of the residual stream and the pre-LayerNorm sums, allocated on first use
void Encoder::ensure_fp32_stream(Workspace& w) {
    const size_t T = static_cast<size_t>(w.tokens), H = info_.hidden_size;
    if (!w.y32) KJ_CUDA(cudaMalloc(&w.y32, T * H * 4));
    if (!w.x32) KJ_CUDA(cudaMalloc(&w.x32, T * H * 4));
}

// One micro-batch: ids/mask/types are device pointers for `nb` sequences of length S.
void Encoder::forward_micro(Workspace& w, int sms, const uint32_t* d_ids, const float* d_mask, const uint32_t* d_types, int nb, int S,
                            const KjcForwardOptions& o, bool noalloc_convention, float* d_out, cudaStream_t st) {
    const int H = info_.hidden_size, I = info_.intermediate_size, M = nb * S, d = H / info_.num_heads;
    const float eps = info_.layer_norm_eps;
    // fp32 residual stream (default for hidden-state output): x32 is the stream, x16 its bf16 copy feeding the tensor cores; the
    // projections run as GEMM -> fp32 sums (+ fp32 residual) -> LayerNorm kernel (writes x32 and x16) instead of the fused kernels
    const bool precise = fp32_residual_for(o);
    if (precise) ensure_fp32_stream(w);
    const bool fused_ln = fused_ln_ && !precise;
    // chained launches run one 128-row tile per CTA: a micro-batch of more tiles than SMs takes them in row chunks of `chunk_tiles`
    // (attention, the embedding and the output kernel still cover the whole micro-batch in one launch each)
    const int m_tiles = (M + kGemmBlockM - 1) / kGemmBlockM;
    // ... and pay off only when most SMs own a tile: a chained launch keeps a tile's two projections on ONE SM (or CTA pair), so a batch
    // of 32 x 128 tokens occupies 32 SMs for the whole layer, while the stand-alone QKV / FFN-up GEMMs spread their column tiles over
    // every SM.  Measured (scripts/c1_latency_ab.py): 32 tiles 366 -> 327 us per forward, 8 tiles 342 -> 284 us, 64 tiles 397 -> 390 us
    const bool chain = chain_ && !precise && !pair_gemm_ && m_tiles >= chain_min_tiles_;
    const int chunk_tiles = std::max(2, sms & ~1);
    // the embedding front end of the chained kernel is bit-identical but measured slower than the two launches (12 gathering warps
    // per SM are latency-bound: 50 us against 17 + 27 us), so it stays opt-in (KJC_CHAIN_EMBED)
    const bool chain_embed = chain && chain_embed_ && !layers_.empty() && m_tiles <= sms;
    {
        EmbedParams e;
        e.ids = d_ids; e.type_ids = d_types; e.word = word_; e.pos = pos_; e.type = type_; e.gamma = emb_g_; e.beta = emb_b_;
        e.x32 = precise ? w.x32 : nullptr; e.x16 = w.x16; e.err_flag = d_err_;
        e.M = M; e.S = S; e.H = H; e.vocab = info_.vocab_size; e.max_pos = info_.max_position_embeddings;
        e.type_vocab = info_.type_vocab_size; e.pos_offset = info_.position_offset; e.eps = eps;
        const int grid = (M + 7) / 8;
        if (chain_embed) {
            // embeddings + embed LN -> layer 0's Q|K|V in one launch   (embeddings/mod.rs:181-326, qkv_projection.rs:93-138)
            prof_begin(KJC_K_GEMM_QKV, st);
            launch_embed_ln_gemm(e, emb_g_, emb_b_, w.t_x16, layers_[0].t_wqkv_192, w.t_qkv16_out64, 3 * H, layers_[0].bqkv, st);
            prof_end(st);
        } else {
            prof_begin(KJC_K_EMBED_LN, st);
            dispatch_nv(H, [&](auto nv) { launch_pdl(embed_layernorm_kernel<decltype(nv)::value>, dim3(grid), dim3(kRowThreads), 0, st, e); });
            KJ_CUDA(cudaGetLastError());
            prof_end(st);
        }
        ++launches_;
    }
    for (size_t li = 0; li < layers_.size(); ++li) {
        const LayerDev& L = layers_[li];
        GemmParams g{};
        // Q|K|V = x Wqkv^T + b                                   (qkv_projection.rs:93-138)
        if (!chain || (li == 0 && !chain_embed)) {  // chained: QKV comes from the embedding launch / the previous layer's FFN-down + LN2 launch
            g.M = M; g.N = 3 * H; g.K = H; g.bias = L.bqkv; g.out = w.qkv16; g.ldo = w.qkv_ld; g.act = ACT_NONE;
            prof_begin(KJC_K_GEMM_QKV, st);
            if (pair_gemm_) launch_gemm_pair(bn_qkv_, EPI_BIAS_BF16, w.t_x16, L.t_wqkv_half, (bn_qkv_ == 192 ? w.t_qkv16_out : w.t_qkv16_out32), g, sms, st);
            else if ((gemm_pair_mask_ & 1) && M > kGemmBlockM) launch_gemm_cta_pair(bn_qkv_, EPI_BIAS_BF16, w.t_x16, L.t_wqkv_half, w.t_qkv16_out, g, sms, st);
            else launch_gemm(bn_qkv_, EPI_BIAS_BF16, w.t_x16, L.t_wqkv, w.t_qkv16_out, g, sms, st);
            prof_end(st);
            ++launches_;
        }
        // softmax(QK^T/sqrt(d) + mask) V, heads merged             (encoder_self_attention.rs:213-298)
        AttnParams a;
        a.qkv = w.qkv16; a.mask = d_mask; a.ctx = w.ctx16; a.B = nb; a.S = S; a.H = H; a.heads = info_.num_heads;
        a.ld_qkv = w.qkv_ld;
        a.scale_log2e = (1.0f / sqrtf(static_cast<float>(d))) * 1.4426950408889634f;
        a.nan_if_all_masked = noalloc_convention ? 1 : 0;
        a.max_ctas = sms;
        prof_begin(KJC_K_ATTENTION, st);
        launch_attention(a, d, st);
        prof_end(st);
        if (chain) {
            const bool last = li + 1 == layers_.size();
            const bool pair_up = (chain_pair_mask_ & 2) != 0, ts_up = chain_ts_ && !pair_up;
            const bool pair_dn = (chain_pair_mask_ & 1) != 0, ts_dn = chain_ts_ && !pair_dn;
            for (int t0 = 0; t0 < m_tiles; t0 += chunk_tiles) {
                const int tn = std::min(chunk_tiles, m_tiles - t0);
                // x = LN1(x + ctx Wo^T + bo) ; t = act(x W1^T + b1)          (encoder_layer.rs:120-147, standard_new.rs:47-73)
                prof_begin(KJC_K_GEMM_FFN_UP, st);
                launch_gemm_ln_gemm(w.t_ctx16, pair_up ? L.t_wo_96 : L.t_wo_ln, w.t_x16_io, w.t_x16, pair_up ? L.t_w1_96 : (ts_up ? L.t_w1_128 : L.t_w1_192),
                                    ts_up ? w.t_h16_out32 : w.t_h16_out64, M, H, L.bo, L.g1, L.be1, eps, I, L.b1, EPI_BIAS_ACT_BF16, act_, st, pair_up, ts_up, 0, t0, tn);
                prof_end(st);
                ++launches_;
                // x = LN2(x + t W2^T + b2) ; next layer's Q|K|V              (standard_new.rs:76-79, encoder_layer.rs:150-176, qkv_projection.rs:93-138)
                if (!last) {
                    const LayerDev& Ln = layers_[li + 1];
                    prof_begin(KJC_K_GEMM_FFN_DOWN, st);
                    launch_gemm_ln_gemm(w.t_h16, pair_dn ? L.t_w2_96 : L.t_w2_ln, w.t_x16_io, w.t_x16,
                                        pair_dn ? Ln.t_wqkv_96 : (ts_dn ? Ln.t_wqkv_128 : Ln.t_wqkv_192), ts_dn ? w.t_qkv16_out32 : w.t_qkv16_out64, M, I, L.b2, L.g2, L.be2, eps,
                                        3 * H, Ln.bqkv, EPI_BIAS_BF16, ACT_NONE, st, pair_dn, ts_dn, 0, t0, tn);
                    prof_end(st);
                    ++launches_;
                } else if (pair_dn && last_ln_pair_) {
                    // last layer, opt-in (KJC_LAST_LN_PAIR): the same CTA-pair launch with an empty second projection (N2 = 0: no phase-2 tiles).
                    // Bit-identical; measured 0.5 % slower in the whole step than gemm_ln_kernel<1> (268.9 vs 270.2 k emb/s), so not the default
                    prof_begin(KJC_K_GEMM_FFN_DOWN, st);
                    launch_gemm_ln_gemm(w.t_h16, L.t_w2_96, w.t_x16_io, w.t_x16, L.t_wqkv_96, w.t_qkv16_out64, M, I, L.b2, L.g2, L.be2, eps, 0, nullptr, EPI_BIAS_BF16,
                                        ACT_NONE, st, true, false, 0, t0, tn);
                    prof_end(st);
                    ++launches_;
                }
            }
            if (last && !(pair_dn && last_ln_pair_)) {
                prof_begin(KJC_K_GEMM_FFN_DOWN, st);
                launch_gemm_ln(w.t_h16, L.t_w2_ln, w.t_x16_io, M, H, I, L.b2, L.g2, L.be2, eps, sms, st);
                prof_end(st);
                ++launches_;
            }
            ++launches_;  // attention
            continue;
        }
        // y = x + ctx Wo^T + bo ; x = LN1(y)                       (encoder_layer.rs:120-147)
        if (fused_ln) {
            prof_begin(KJC_K_GEMM_OUT, st);
            launch_gemm_ln(w.t_ctx16, L.t_wo_ln, w.t_x16_io, M, H, H, L.bo, L.g1, L.be1, eps, sms, st);
            prof_end(st);
        } else {
            g = GemmParams{};
            g.M = M; g.N = H; g.K = H; g.bias = L.bo; g.residual = w.x16; g.residual32 = precise ? w.x32 : nullptr; g.ldr = H; g.out = w.y32; g.ldo = H; g.act = ACT_NONE;
            prof_begin(KJC_K_GEMM_OUT, st);
            launch_gemm(bn_h_, EPI_BIAS_RES_F32, w.t_ctx16, L.t_wo, w.t_qkv16_out, g, sms, st);
            prof_end(st);
            prof_begin(KJC_K_LAYERNORM, st);
            launch_layernorm(w.y32, L.g1, L.be1, eps, precise ? w.x32 : nullptr, w.x16, M, H, st);
            prof_end(st);
            ++launches_;
        }
        if (fused_ffn_ && !precise) {
            // x = LN2(x + act(x W1^T + b1) W2^T + b2) in one kernel      (standard_new.rs:47-80, encoder_layer.rs:150-176)
            prof_begin(KJC_K_GEMM_FFN_UP, st);
            launch_ffn_ln(w.t_x16, L.t_w1_ffn, L.t_w1_ffn32, L.t_w2_ffn, M, I, L.b1, L.b2, L.g2, L.be2, eps, act_, sms, st);
            prof_end(st);
            launches_ += 3;
            continue;
        }
        // t = act(x W1^T + b1)                                     (standard_new.rs:47-73)
        g = GemmParams{};
        g.M = M; g.N = I; g.K = H; g.bias = L.b1; g.out = w.h16; g.ldo = I; g.act = act_;
        prof_begin(KJC_K_GEMM_FFN_UP, st);
        if (pair_gemm_) launch_gemm_pair(bn_i_, EPI_BIAS_ACT_BF16, w.t_x16, L.t_w1_half, (bn_i_ == 192 ? w.t_h16_out : w.t_h16_out32), g, sms, st);
        else if ((gemm_pair_mask_ & 2) && M > kGemmBlockM) launch_gemm_cta_pair(bn_i_, EPI_BIAS_ACT_BF16, w.t_x16, L.t_w1_half, w.t_h16_out, g, sms, st);
        else launch_gemm(bn_i_, EPI_BIAS_ACT_BF16, w.t_x16, L.t_w1, w.t_h16_out, g, sms, st);
        prof_end(st);
        // y = x + t W2^T + b2 ; x = LN2(y)                         (standard_new.rs:76-79, encoder_layer.rs:150-176)
        if (fused_ln) {
            prof_begin(KJC_K_GEMM_FFN_DOWN, st);
            launch_gemm_ln(w.t_h16, L.t_w2_ln, w.t_x16_io, M, H, I, L.b2, L.g2, L.be2, eps, sms, st);
            prof_end(st);
        } else {
            g = GemmParams{};
            g.M = M; g.N = H; g.K = I; g.bias = L.b2; g.residual = w.x16; g.residual32 = precise ? w.x32 : nullptr; g.ldr = H; g.out = w.y32; g.ldo = H; g.act = ACT_NONE;
            prof_begin(KJC_K_GEMM_FFN_DOWN, st);
            launch_gemm(bn_h_, EPI_BIAS_RES_F32, w.t_h16, L.t_w2, w.t_qkv16_out, g, sms, st);
            prof_end(st);
            prof_begin(KJC_K_LAYERNORM, st);
            launch_layernorm(w.y32, L.g2, L.be2, eps, precise ? w.x32 : nullptr, w.x16, M, H, st);
            prof_end(st);
            ++launches_;
        }
        launches_ += 4;
    }
    prof_begin(KJC_K_OUTPUT, st);
    if (o.output == KJC_OUT_HIDDEN && precise) {
        KJ_CUDA(cudaMemcpyAsync(d_out, w.x32, static_cast<size_t>(M) * H * 4, cudaMemcpyDeviceToDevice, st));  // the fp32 stream itself
    } else if (o.output == KJC_OUT_HIDDEN) {
        const size_t n4 = static_cast<size_t>(M) * H / 4;
        bf16_to_f32_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(w.x16, d_out, n4);
        KJ_CUDA(cudaGetLastError());
        ++launches_;
    } else if (o.output == KJC_OUT_POOLED && precise) {
        pool_l2_kernel<float><<<nb, 256, 0, st>>>(w.x32, d_mask, d_out, S, H, o.pooling, o.normalize);
        KJ_CUDA(cudaGetLastError());
        ++launches_;
    } else if (o.output == KJC_OUT_POOLED) {
        if (o.pooling == KJC_POOL_MEAN && H <= 512 && H % 4 == 0) {
            // the 16-warp kernel keeps NV float4 accumulators per lane: up to hidden 512 without spills; wider models (768, 1024: a
            // negligible share of a 12-layer step) take the generic kernel below
            const size_t smem = static_cast<size_t>(kPoolWarps) * H * sizeof(float);  // <= 32 KB
            dispatch_nv(H, [&](auto nv) {
                constexpr int NV = decltype(nv)::value;
                if constexpr (NV <= 4) {
                    launch_pdl(mean_pool_l2_kernel<NV>, dim3(nb), dim3(kPoolThreads), smem, st, static_cast<const __nv_bfloat16*>(w.x16), d_mask, d_out, S, H,
                               static_cast<int>(o.normalize));
                }
            });
        } else {
            pool_l2_kernel<__nv_bfloat16><<<nb, 256, 0, st>>>(w.x16, d_mask, d_out, S, H, o.pooling, o.normalize);
        }
        KJ_CUDA(cudaGetLastError());
        ++launches_;
    } else if (precise) {
        HeadParams<float> hp;
        hp.x = w.x32; hp.w_pre = w_pre_; hp.b_pre = b_pre_; hp.w_cls = w_cls_; hp.b_cls = b_cls_; hp.logits = d_out;
        hp.B = nb; hp.S = S; hp.H = H; hp.C = info_.num_labels;
        hp.act = info_.head_kind == KJC_HEAD_PRE_RELU ? HEAD_RELU : (info_.head_kind == KJC_HEAD_LINEAR ? HEAD_NONE : HEAD_TANH);
        hp.z1 = w.head32;
        launch_cls_head(hp, st);
        KJ_CUDA(cudaGetLastError());
        launches_ += w_pre_ ? 2 : 1;
    } else {
        HeadParams<__nv_bfloat16> hp;
        hp.x = w.x16; hp.w_pre = w_pre_; hp.b_pre = b_pre_; hp.w_cls = w_cls_; hp.b_cls = b_cls_; hp.logits = d_out;
        hp.B = nb; hp.S = S; hp.H = H; hp.C = info_.num_labels;
        hp.act = info_.head_kind == KJC_HEAD_PRE_RELU ? HEAD_RELU : (info_.head_kind == KJC_HEAD_LINEAR ? HEAD_NONE : HEAD_TANH);
        hp.z1 = w.head32;  // sized by ensure_workspace for one sequence per token
        launch_cls_head(hp, st);
        KJ_CUDA(cudaGetLastError());
        launches_ += w_pre_ ? 2 : 1;
    }
    prof_end(st);
}

// ---------------------------------------------------------------- profiling
cudaEvent_t Encoder::prof_event() {
    if (!prof_pool_.empty()) {
        cudaEvent_t e = prof_pool_.back();
        prof_pool_.pop_back();
        return e;
    }
    cudaEvent_t e;
    KJ_CUDA(cudaEventCreate(&e));
    return e;
}
void Encoder::prof_begin(int cls, cudaStream_t st) {
    if (!profiling_) return;
    ProfRec r{cls, prof_event(), prof_event()};
    KJ_CUDA(cudaEventRecord(r.a, st));
    prof_recs_.push_back(r);
}
void Encoder::prof_end(cudaStream_t st) {
    if (!profiling_) return;
    KJ_CUDA(cudaEventRecord(prof_recs_.back().b, st));
}
void Encoder::prof_collect() {
    for (ProfRec& r : prof_recs_) {
        KJ_CUDA(cudaEventSynchronize(r.b));
        float ms = 0.f;
        KJ_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        prof_ms_[r.cls] += ms;
        prof_n_[r.cls] += 1;
        prof_pool_.push_back(r.a);
        prof_pool_.push_back(r.b);
    }
    prof_recs_.clear();
}
void Encoder::set_profiling(bool on) {
    std::lock_guard<std::mutex> lock(mu_);
    KJ_CUDA(cudaSetDevice(info_.device));
    prof_collect();
    profiling_ = on;
    if (on) {
        for (int i = 0; i < KJC_NUM_KERNEL_CLASSES; ++i) { prof_ms_[i] = 0; prof_n_[i] = 0; }
    }
}
void Encoder::get_profile(double* ms, int64_t* launches) {
    std::lock_guard<std::mutex> lock(mu_);
    KJ_CUDA(cudaSetDevice(info_.device));
    prof_collect();
    for (int i = 0; i < KJC_NUM_KERNEL_CLASSES; ++i) { ms[i] = prof_ms_[i]; launches[i] = prof_n_[i]; }
}

size_t Encoder::out_row_elems(const KjcForwardOptions& o, int S) const {
    if (o.output == KJC_OUT_HIDDEN) return static_cast<size_t>(S) * info_.hidden_size;
    if (o.output == KJC_OUT_POOLED) return static_cast<size_t>(info_.hidden_size);
    return static_cast<size_t>(info_.num_labels);
}

void Encoder::validate(int B, int S, const KjcForwardOptions& o) const {
    if (B <= 0 || S <= 0) throw Error(KJC_INVALID_CONFIG, "batch and seq_len must be positive");
    if (S > 512) throw Error(KJC_INVALID_CONFIG, "seq_len > 512 is not supported by the fused attention kernel");
    if (o.output < KJC_OUT_HIDDEN || o.output > KJC_OUT_LOGITS) throw Error(KJC_INVALID_CONFIG, "unknown output mode");
    if (o.output == KJC_OUT_LOGITS && info_.head_kind == KJC_HEAD_ABSENT)
        throw Error(KJC_INVALID_CONFIG, "model has no classification head (no classifier tensors in the checkpoint)");
    if (o.output == KJC_OUT_POOLED && (o.pooling < KJC_POOL_MEAN || o.pooling > KJC_POOL_LAST))
        throw Error(KJC_INVALID_CONFIG, "unknown pooling strategy");
}

bool Encoder::resolve_noalloc(int B, int S, const KjcForwardOptions& o) const {
    if (o.mask_convention == KJC_MASK_ALLOC) return false;
    if (o.mask_convention == KJC_MASK_NOALLOC) return true;
    if (o.output == KJC_OUT_LOGITS) return false;  // Classifier / Reranker always take the alloc path
    const long tokens = static_cast<long>(B) * S;  // ComputeStrategy::select, KT/cpu/strategy.rs:29-47
    return tokens <= 1 || tokens >= 1000;
}

// Micro-batches round-robin over the lanes.  `st` is the caller-visible stream: lane streams start after everything
// already enqueued on it and `st` resumes after the last lane has finished.
void Encoder::forward_batches(const uint32_t* d_ids, const float* d_mask, const uint32_t* d_types, int B, int S, const KjcForwardOptions& o,
                              float* d_out, cudaStream_t st) {
    const int mb = micro_batch(S);
    const bool noalloc = resolve_noalloc(B, S, o);
    const size_t row = out_row_elems(o, S);
    const int n_mb = (B + mb - 1) / mb;
    // one lane (all SMs, caller's stream) for a single micro-batch or while per-kernel profiling is on
    const int lanes = (profiling_ || n_mb < 2) ? 1 : std::min(lanes_, n_mb);
    if (lanes == 1) {
        ensure_workspace(ws_[0], std::min(B, mb) * S);
        for (int b0 = 0; b0 < B; b0 += mb) {
            const int nb = std::min(mb, B - b0);
            const size_t t0 = static_cast<size_t>(b0) * S;
            forward_micro(ws_[0], num_sms_, d_ids + t0, d_mask ? d_mask + t0 : nullptr, d_types ? d_types + t0 : nullptr, nb, S, o, noalloc,
                          d_out + b0 * row, st);
        }
        return;
    }
    const int sms = std::max(1, num_sms_ / lanes);
    for (int l = 0; l < lanes; ++l) ensure_workspace(ws_[l], mb * S);
    KJ_CUDA(cudaEventRecord(ev_in_, st));
    for (int l = 0; l < lanes; ++l) KJ_CUDA(cudaStreamWaitEvent(ws_[l].stream, ev_in_, 0));
    int i = 0;
    for (int b0 = 0; b0 < B; b0 += mb, ++i) {
        Workspace& w = ws_[i % lanes];
        const int nb = std::min(mb, B - b0);
        const size_t t0 = static_cast<size_t>(b0) * S;
        forward_micro(w, sms, d_ids + t0, d_mask ? d_mask + t0 : nullptr, d_types ? d_types + t0 : nullptr, nb, S, o, noalloc,
                      d_out + b0 * row, w.stream);
    }
    for (int l = 0; l < lanes; ++l) {
        KJ_CUDA(cudaEventRecord(ws_[l].done, ws_[l].stream));
        KJ_CUDA(cudaStreamWaitEvent(st, ws_[l].done, 0));
    }
}

void Encoder::forward_device(const uint32_t* d_ids, const float* d_mask, const uint32_t* d_types, int B, int S,
                             const KjcForwardOptions& o, float* d_out, cudaStream_t st) {
    validate(B, S, o);
    std::lock_guard<std::mutex> lock(mu_);
    KJ_CUDA(cudaSetDevice(info_.device));
    if (!st) st = stream_;
    launches_ = 0;
    forward_batches(d_ids, d_mask, d_types, B, S, o, d_out, st);
}

void Encoder::forward_host(const uint32_t* ids, const float* mask, const uint32_t* types, int B, int S, const KjcForwardOptions& o,
                           float* out, const RowSink* sink) {
    validate(B, S, o);
    std::lock_guard<std::mutex> lock(mu_);
    KJ_CUDA(cudaSetDevice(info_.device));
    launches_ = 0;
    const size_t T = static_cast<size_t>(B) * S;
    const size_t row = out_row_elems(o, S);
    const size_t in_words = T * 3, out_elems = static_cast<size_t>(B) * row;
    if (in_words > in_cap_) {
        if (d_in_) cudaFree(d_in_);
        if (h_stage_in_) cudaFreeHost(h_stage_in_);
        KJ_CUDA(cudaMalloc(&d_in_, in_words * 4));
        KJ_CUDA(cudaMallocHost(&h_stage_in_, in_words * 4));
        in_cap_ = in_words;
    }
    if (out_elems > out_cap_) {
        if (d_out_) cudaFree(d_out_);
        if (h_stage_out_) cudaFreeHost(h_stage_out_);
        KJ_CUDA(cudaMalloc(&d_out_, out_elems * 4));
        KJ_CUDA(cudaMallocHost(&h_stage_out_, out_elems * 4));
        out_cap_ = out_elems;
    }
    // Stage through pinned memory so the copies are true async DMA transfers, in chunks of two micro-batches, and keep the copies
    // OFF the compute stream: inputs go up on a copy-in stream (chunk c + 1 lands while chunk c computes), results come down on a
    // copy-out stream, events order the three.  With everything on one stream every chunk paid its H2D + D2H in SM idle time
    // (14 chunks x ~35 us of a 17 ms call).  The host-side copy of chunk c + 1 into the staging buffer and of chunk c - 1 out of it
    // overlap the GPU work on chunk c, so only the first stage-in and the last stage-out are exposed.
    uint32_t* hs = h_stage_in_;
    const uint32_t* d_ids = d_in_;
    const float* d_mask = mask ? reinterpret_cast<const float*>(d_in_ + T) : nullptr;
    const uint32_t* d_types = types ? d_in_ + 2 * T : nullptr;
    KjcForwardOptions oc = o;  // the padding convention is decided once for the whole call, not per chunk
    oc.mask_convention = resolve_noalloc(B, S, o) ? KJC_MASK_NOALLOC : KJC_MASK_ALLOC;
    const int chunk = std::max(1, 2 * micro_batch(S));
    const int n_chunks = (B + chunk - 1) / chunk;
    if (n_chunks == 1) {
        // One chunk (the latency case): nothing to overlap, so everything goes on the compute stream -- one H2D of the contiguous
        // ids | mask | types span, the kernels, the D2H of the rows and of the error flag, one synchronise -- instead of three streams
        // and four event hops (~35 us of a 0.38 ms call at 32 x 128 tokens).
        memcpy(hs, ids, T * 4);
        if (mask) memcpy(hs + T, mask, T * 4);
        if (types) memcpy(hs + 2 * T, types, T * 4);
        const size_t span = types ? 3 * T : (mask ? 2 * T : T);
        KJ_CUDA(cudaMemcpyAsync(d_in_, hs, span * 4, cudaMemcpyHostToDevice, stream_));
        forward_batches(d_ids, d_mask, d_types, B, S, oc, d_out_, stream_);
        KJ_CUDA(cudaMemcpyAsync(h_stage_out_, d_out_, out_elems * 4, cudaMemcpyDeviceToHost, stream_));
        KJ_CUDA(cudaMemcpyAsync(&err_host_, d_err_, sizeof(int), cudaMemcpyDeviceToHost, stream_));
        KJ_CUDA(cudaStreamSynchronize(stream_));
        if (err_host_) {
            KJ_CUDA(cudaMemsetAsync(d_err_, 0, sizeof(int), stream_));
            KJ_CUDA(cudaStreamSynchronize(stream_));
            throw Error(KJC_INFERENCE_FAILED, "Token type ID out of range");
        }
        if (sink) (*sink)(h_stage_out_, 0, static_cast<size_t>(B));
        else memcpy(out, h_stage_out_, out_elems * 4);
        return;
    }
    if (!s_in_) {
        KJ_CUDA(cudaStreamCreateWithFlags(&s_in_, cudaStreamNonBlocking));
        KJ_CUDA(cudaStreamCreateWithFlags(&s_out_, cudaStreamNonBlocking));
    }
    while (static_cast<int>(ev_in_chunk_.size()) < n_chunks) {
        cudaEvent_t e[3];
        for (cudaEvent_t& x : e) KJ_CUDA(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
        ev_in_chunk_.push_back(e[0]);
        ev_done_chunk_.push_back(e[1]);
        ev_out_chunk_.push_back(e[2]);
    }
    // the previous call on this handle has fully drained (it ended with a synchronise), so the staging buffers are free
    auto stage_in = [&](int c) {
        const int b0 = c * chunk, nb = std::min(chunk, B - b0);
        const size_t t0 = static_cast<size_t>(b0) * S, tn = static_cast<size_t>(nb) * S;
        memcpy(hs + t0, ids + t0, tn * 4);
        KJ_CUDA(cudaMemcpyAsync(d_in_ + t0, hs + t0, tn * 4, cudaMemcpyHostToDevice, s_in_));
        if (mask) {
            memcpy(hs + T + t0, mask + t0, tn * 4);
            KJ_CUDA(cudaMemcpyAsync(d_in_ + T + t0, hs + T + t0, tn * 4, cudaMemcpyHostToDevice, s_in_));
        }
        if (types) {
            memcpy(hs + 2 * T + t0, types + t0, tn * 4);
            KJ_CUDA(cudaMemcpyAsync(d_in_ + 2 * T + t0, hs + 2 * T + t0, tn * 4, cudaMemcpyHostToDevice, s_in_));
        }
        KJ_CUDA(cudaEventRecord(ev_in_chunk_[c], s_in_));
    };
    auto hand_over = [&](int c) {  // chunk c's rows sit in the pinned output buffer
        const int b0 = c * chunk, nb = std::min(chunk, B - b0);
        KJ_CUDA(cudaEventSynchronize(ev_out_chunk_[c]));
        if (sink) (*sink)(h_stage_out_ + b0 * row, static_cast<size_t>(b0), static_cast<size_t>(nb));
        else memcpy(out + b0 * row, h_stage_out_ + b0 * row, nb * row * 4);
    };
    stage_in(0);
    for (int c = 0; c < n_chunks; ++c) {
        const int b0 = c * chunk, nb = std::min(chunk, B - b0);
        const size_t t0 = static_cast<size_t>(b0) * S;
        if (c + 1 < n_chunks) stage_in(c + 1);  // one chunk ahead of the compute stream
        KJ_CUDA(cudaStreamWaitEvent(stream_, ev_in_chunk_[c], 0));
        forward_batches(d_ids + t0, d_mask ? d_mask + t0 : nullptr, d_types ? d_types + t0 : nullptr, nb, S, oc, d_out_ + b0 * row, stream_);
        KJ_CUDA(cudaEventRecord(ev_done_chunk_[c], stream_));
        KJ_CUDA(cudaStreamWaitEvent(s_out_, ev_done_chunk_[c], 0));
        KJ_CUDA(cudaMemcpyAsync(h_stage_out_ + b0 * row, d_out_ + b0 * row, nb * row * 4, cudaMemcpyDeviceToHost, s_out_));
        KJ_CUDA(cudaEventRecord(ev_out_chunk_[c], s_out_));
        if (c > 0) hand_over(c - 1);
    }
    int err = 0;
    KJ_CUDA(cudaMemcpyAsync(&err_host_, d_err_, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    KJ_CUDA(cudaStreamSynchronize(stream_));
    err = err_host_;
    if (err) {
        KJ_CUDA(cudaMemsetAsync(d_err_, 0, sizeof(int), stream_));
        KJ_CUDA(cudaStreamSynchronize(s_out_));
        throw Error(KJC_INFERENCE_FAILED, "Token type ID out of range");
    }
    hand_over(n_chunks - 1);  // the last chunk (earlier ones were handed over in the loop)
}

// Head stage alone on caller-supplied fp32 hidden states (debug hook: the argmax stage must be
// bit-exact against the oracle when both see the same fp32 input).
void Encoder::head_only_host(const float* hidden, int B, int S, float* logits) {
    if (info_.head_kind == KJC_HEAD_ABSENT) throw Error(KJC_INVALID_CONFIG, "model has no classification head");
    std::lock_guard<std::mutex> lock(mu_);
    KJ_CUDA(cudaSetDevice(info_.device));
    const int H = info_.hidden_size, Cn = info_.num_labels;
    float *dh, *dl;
    KJ_CUDA(cudaMalloc(&dh, static_cast<size_t>(B) * S * H * 4));
    KJ_CUDA(cudaMalloc(&dl, static_cast<size_t>(B) * Cn * 4));
    KJ_CUDA(cudaMemcpy(dh, hidden, static_cast<size_t>(B) * S * H * 4, cudaMemcpyHostToDevice));
    HeadParams<float> hp;
    hp.x = dh; hp.w_pre = w_pre_; hp.b_pre = b_pre_; hp.w_cls = w_cls_; hp.b_cls = b_cls_; hp.logits = dl;
    hp.B = B; hp.S = S; hp.H = H; hp.C = Cn;
    hp.act = info_.head_kind == KJC_HEAD_PRE_RELU ? HEAD_RELU : (info_.head_kind == KJC_HEAD_LINEAR ? HEAD_NONE : HEAD_TANH);
    float* dz = nullptr;
    KJ_CUDA(cudaMalloc(&dz, static_cast<size_t>(B) * H * 4));
    hp.z1 = dz;
    launch_cls_head(hp, stream_);
    KJ_CUDA(cudaGetLastError());
    KJ_CUDA(cudaStreamSynchronize(stream_));
    KJ_CUDA(cudaMemcpy(logits, dl, static_cast<size_t>(B) * Cn * 4, cudaMemcpyDeviceToHost));
    cudaFree(dh);
    cudaFree(dl);
    cudaFree(dz);
}

// -------------------------------------------------------------- debug hooks
// Single-kernel entry points (include/kjarni_cuda_debug.h): host buffers in, host buffers out.
void dbg_gemm(const uint16_t* a_bf16, const uint16_t* w_bf16, const float* bias, const float* residual, int M, int N, int K, int epi,
              int act, int block_n, void* out) {
                cudaDeviceProp prop;
        int dev = 0;
        KJ_CUDA(cudaGetDevice(&dev));
        KJ_CUDA(cudaGetDeviceProperties(&prop, dev));
        const bool pair2 = block_n >= 2000;  // + 2000: the CTA-pair form of gemm_tcgen05_kernel
        if (pair2) block_n -= 2000;
        const bool pair = block_n >= 1000;   // + 1000: the experimental A-resident pair kernel (gemm_pair.cuh)
        if (pair) block_n -= 1000;
        const int bn = block_n > 0 ? block_n : pick_block_n(N);
        const size_t Mp = std::max(M, 128), Np = std::max(N, bn);
        __nv_bfloat16 *dA, *dW;
        float* dB = nullptr;
        __nv_bfloat16* dR = nullptr;
        void* dO;
        const bool f32out = epi == EPI_BIAS_RES_F32 || epi == EPI_BIAS_F32;
        KJ_CUDA(cudaMalloc(&dA, Mp * K * 2));
        KJ_CUDA(cudaMalloc(&dW, Np * K * 2));
        KJ_CUDA(cudaMemset(dA, 0, Mp * K * 2));
        KJ_CUDA(cudaMemset(dW, 0, Np * K * 2));
        KJ_CUDA(cudaMalloc(&dO, static_cast<size_t>(M) * N * (f32out ? 4 : 2)));
        KJ_CUDA(cudaMemcpy(dA, a_bf16, static_cast<size_t>(M) * K * 2, cudaMemcpyHostToDevice));
        KJ_CUDA(cudaMemcpy(dW, w_bf16, static_cast<size_t>(N) * K * 2, cudaMemcpyHostToDevice));
        if (bias) { KJ_CUDA(cudaMalloc(&dB, static_cast<size_t>(N) * 4)); KJ_CUDA(cudaMemcpy(dB, bias, static_cast<size_t>(N) * 4, cudaMemcpyHostToDevice)); }
        if (residual) {  // the residual stream is bf16 on device
            std::vector<__nv_bfloat16> rb(static_cast<size_t>(M) * N);
            for (size_t i = 0; i < rb.size(); ++i) rb[i] = __float2bfloat16_rn(residual[i]);
            KJ_CUDA(cudaMalloc(&dR, rb.size() * 2));
            KJ_CUDA(cudaMemcpy(dR, rb.data(), rb.size() * 2, cudaMemcpyHostToDevice));
        }
        CUtensorMap ta = make_tmap_2d(dA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Mp, K, kGemmBlockM, kGemmBlockK, 128);
        CUtensorMap tb = make_tmap_2d(dW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Np, K, bn, kGemmBlockK, 128);
        GemmParams p{};
        p.M = M; p.N = N; p.K = K; p.bias = dB; p.residual = dR; p.ldr = N; p.out = dO; p.ldo = N; p.act = act;
        CUtensorMap tc = ta;
        if (!f32out) tc = ((!pair && gemm_wide_store(bn)) || (bn == 192 && pair)) ? make_tmap_2d(dO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, 32, 64, 128)
                                               : make_tmap_2d(dO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, 32, kEpiChunkCols, 64);
        if (pair2) {
            CUtensorMap tbh = make_tmap_2d(dW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Np, K, bn / 2, kGemmBlockK, 128);
            launch_gemm_cta_pair(bn, epi, ta, tbh, tc, p, prop.multiProcessorCount, nullptr);
        } else if (pair) {
            CUtensorMap tbh = make_tmap_2d(dW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Np, K, bn / 2, kGemmBlockK, 128);
            launch_gemm_pair(bn, epi, ta, tbh, tc, p, prop.multiProcessorCount, nullptr);
        } else
        launch_gemm(bn, epi, ta, tb, tc, p, prop.multiProcessorCount, nullptr);
        KJ_CUDA(cudaDeviceSynchronize());
        KJ_CUDA(cudaMemcpy(out, dO, static_cast<size_t>(M) * N * (f32out ? 4 : 2), cudaMemcpyDeviceToHost));
        cudaFree(dA); cudaFree(dW); cudaFree(dO);
        if (dB) cudaFree(dB);
        if (dR) cudaFree(dR);
}

// out[M,H] (bf16), H = 384 or 768, = LN(A W^T + bias + residual) with the fused kernel; residual/out bf16 bit patterns.
void dbg_gemm_ln(const uint16_t* a_bf16, const uint16_t* w_bf16, const float* bias, const float* gamma, const float* beta, float eps,
                 const uint16_t* res_bf16, int M, int H, int K, uint16_t* out_bf16, int iters, float* us) {
    cudaDeviceProp prop;
    int dev = 0;
    KJ_CUDA(cudaGetDevice(&dev));
    KJ_CUDA(cudaGetDeviceProperties(&prop, dev));
    const size_t Mp = std::max(M, 128);
    __nv_bfloat16 *dA, *dW, *dX;
    float *dB, *dG, *dBt;
    KJ_CUDA(cudaMalloc(&dA, Mp * K * 2));
    KJ_CUDA(cudaMalloc(&dW, static_cast<size_t>(H) * K * 2));
    KJ_CUDA(cudaMalloc(&dX, Mp * H * 2));
    KJ_CUDA(cudaMemset(dA, 0, Mp * K * 2));
    KJ_CUDA(cudaMemset(dX, 0, Mp * H * 2));
    KJ_CUDA(cudaMalloc(&dB, H * 4)); KJ_CUDA(cudaMalloc(&dG, H * 4)); KJ_CUDA(cudaMalloc(&dBt, H * 4));
    KJ_CUDA(cudaMemcpy(dA, a_bf16, static_cast<size_t>(M) * K * 2, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dW, w_bf16, static_cast<size_t>(H) * K * 2, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dX, res_bf16, static_cast<size_t>(M) * H * 2, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dB, bias, H * 4, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dG, gamma, H * 4, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dBt, beta, H * 4, cudaMemcpyHostToDevice));
    CUtensorMap ta = make_tmap_2d(dA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Mp, K, kGemmBlockM, kGemmBlockK, 128);
    CUtensorMap tw = make_tmap_2d(dW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, H, K, kLnHalfN, kGemmBlockK, 128);
    CUtensorMap tio = make_tmap_2d(dX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, H, 32, kEpiChunkCols, 64);
    launch_gemm_ln(ta, tw, tio, M, H, K, dB, dG, dBt, eps, prop.multiProcessorCount, nullptr);
    KJ_CUDA(cudaDeviceSynchronize());
    KJ_CUDA(cudaMemcpy(out_bf16, dX, static_cast<size_t>(M) * H * 2, cudaMemcpyDeviceToHost));
    if (iters > 0 && us) {  // timing (in place: the values drift, the work does not)
        cudaEvent_t e0, e1;
        KJ_CUDA(cudaEventCreate(&e0));
        KJ_CUDA(cudaEventCreate(&e1));
        KJ_CUDA(cudaEventRecord(e0, nullptr));
        for (int i = 0; i < iters; ++i) launch_gemm_ln(ta, tw, tio, M, H, K, dB, dG, dBt, eps, prop.multiProcessorCount, nullptr);
        KJ_CUDA(cudaEventRecord(e1, nullptr));
        KJ_CUDA(cudaDeviceSynchronize());
        float ms = 0.f;
        KJ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        *us = ms * 1e3f / iters;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    cudaFree(dA); cudaFree(dW); cudaFree(dX); cudaFree(dB); cudaFree(dG); cudaFree(dBt);
}

// x_out[M,384] (bf16) = LN(x + act(x W1^T + b1) W2^T + b2) with the fused FFN kernel (in place on a device copy of x).
// Chained kernel alone (gemm_ln_gemm.cuh): x' = LN(A W1^T + b1 + res), out2 = act(x' W2^T + b2); optional timing.
void dbg_gemm_ln_gemm(const uint16_t* a_bf16, const uint16_t* w1_bf16, const float* bias1, const float* gamma, const float* beta, float eps,
                      const uint16_t* res_bf16, int M, int K1, const uint16_t* w2_bf16, const float* bias2, int N2, int epi2, int act,
                      uint16_t* out_x_bf16, uint16_t* out2_bf16, int iters, float* us) {
    cudaDeviceProp prop;
    int dev = 0;
    KJ_CUDA(cudaGetDevice(&dev));
    KJ_CUDA(cudaGetDeviceProperties(&prop, dev));
    if ((M + kGemmBlockM - 1) / kGemmBlockM > prop.multiProcessorCount) throw Error(KJC_INVALID_CONFIG, "chained kernel: one 128-row tile per SM at most");
    const bool pair = (epi2 & 16) != 0;  // epi2 + 16: the CTA-pair variant (cta_group::2, each CTA loads half of every weight tile)
    const bool ts = (epi2 & 32) != 0;    // epi2 + 32: x' as the phase-2 A operand in tensor memory (128-column phase-2 tiles)
    if (pair && ts) throw Error(KJC_INVALID_CONFIG, "chained kernel: the CTA-pair and tensor-memory variants exclude each other");
    const int dbg = (epi2 >> 6) & 7;     // epi2 + 64 / 128 / 256: phase-2 knock-outs (GemmLnGemmParams::dbg 1 / 2 / 4), timing only
    epi2 &= 15;
    if (N2 > kLg2BiasMax) throw Error(KJC_INVALID_CONFIG, "chained kernel: N2 too large");
    const size_t Mp = std::max(M, 128), N2p = std::max(N2, kLg2BN);
    __nv_bfloat16 *dA, *dW, *dX, *dW2, *dO;
    float *dB, *dG, *dBt, *dB2;
    KJ_CUDA(cudaMalloc(&dA, Mp * K1 * 2));
    KJ_CUDA(cudaMalloc(&dW, static_cast<size_t>(kLnN) * K1 * 2));
    KJ_CUDA(cudaMalloc(&dX, Mp * kLnN * 2));
    KJ_CUDA(cudaMalloc(&dW2, N2p * kLnN * 2));
    KJ_CUDA(cudaMalloc(&dO, Mp * N2 * 2));
    KJ_CUDA(cudaMemset(dA, 0, Mp * K1 * 2));
    KJ_CUDA(cudaMemset(dX, 0, Mp * kLnN * 2));
    KJ_CUDA(cudaMemset(dW2, 0, N2p * kLnN * 2));
    KJ_CUDA(cudaMemset(dO, 0, Mp * N2 * 2));
    KJ_CUDA(cudaMalloc(&dB, kLnN * 4)); KJ_CUDA(cudaMalloc(&dG, kLnN * 4)); KJ_CUDA(cudaMalloc(&dBt, kLnN * 4));
    KJ_CUDA(cudaMalloc(&dB2, static_cast<size_t>(N2) * 4));
    KJ_CUDA(cudaMemcpy(dA, a_bf16, static_cast<size_t>(M) * K1 * 2, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dW, w1_bf16, static_cast<size_t>(kLnN) * K1 * 2, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dX, res_bf16, static_cast<size_t>(M) * kLnN * 2, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dW2, w2_bf16, static_cast<size_t>(N2) * kLnN * 2, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dB, bias1, kLnN * 4, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dG, gamma, kLnN * 4, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dBt, beta, kLnN * 4, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dB2, bias2, static_cast<size_t>(N2) * 4, cudaMemcpyHostToDevice));
    CUtensorMap ta = make_tmap_2d(dA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Mp, K1, kGemmBlockM, kGemmBlockK, 128);
    CUtensorMap tw = make_tmap_2d(dW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kLnN, K1, pair ? kLnHalfN / 2 : kLnHalfN, kGemmBlockK, 128);
    CUtensorMap tres = make_tmap_2d(dX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, kLnN, 32, kEpiChunkCols, 64);
    CUtensorMap tx = make_tmap_2d(dX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, kLnN, kGemmBlockM, kGemmBlockK, 128);
    CUtensorMap tw2 = make_tmap_2d(dW2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N2p, kLnN, pair ? kLg2BN / 2 : (ts ? kLgTBN : kLg2BN), kGemmBlockK, 128);
    CUtensorMap to2 = ts ? make_tmap_2d(dO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N2, 32, kEpiChunkCols, 64)
                         : make_tmap_2d(dO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N2, 32, 64, 128);
    auto go = [&] { launch_gemm_ln_gemm(ta, tw, tres, tx, tw2, to2, M, K1, dB, dG, dBt, eps, N2, dB2, epi2, act, nullptr, pair, ts, dbg); };
    go();
    KJ_CUDA(cudaDeviceSynchronize());
#if KJ_LG_TRACE
    if (getenv("KJC_LG_TRACE")) {  // one traced launch after two warm ones: clock64 stamps of CTAs 0 and 100, in clocks since the CTA's entry
        const int ctas = (M + kGemmBlockM - 1) / kGemmBlockM;
        unsigned long long* dT;
        KJ_CUDA(cudaMalloc(&dT, static_cast<size_t>(ctas + 1) * 256 * 8));
        KJ_CUDA(cudaMemset(dT, 0, static_cast<size_t>(ctas + 1) * 256 * 8));
        go(); go();
        g_lg_trace = dT;
        go();
        g_lg_trace = nullptr;
        KJ_CUDA(cudaDeviceSynchronize());
        std::vector<unsigned long long> h(static_cast<size_t>(ctas) * 256);
        KJ_CUDA(cudaMemcpy(h.data(), dT, h.size() * 8, cudaMemcpyDeviceToHost));
        for (int cta : {0, std::min(100, ctas - 1)}) {
            const unsigned long long* t = h.data() + static_cast<size_t>(cta) * 256;
            fprintf(stderr, "lgtrace K1=%d N2=%d epi2=%d dbg=%d cta %d:", K1, N2, epi2, dbg, cta);
            for (int i = 1; i < 256; ++i)
                if (t[i] != 0) fprintf(stderr, " %d=%lld", i, static_cast<long long>(t[i] - t[0]));
            fprintf(stderr, "\n");
        }
        cudaFree(dT);
    }
#endif
    KJ_CUDA(cudaMemcpy(out_x_bf16, dX, static_cast<size_t>(M) * kLnN * 2, cudaMemcpyDeviceToHost));
    KJ_CUDA(cudaMemcpy(out2_bf16, dO, static_cast<size_t>(M) * N2 * 2, cudaMemcpyDeviceToHost));
    if (iters > 0 && us) {  // timing (in place: the values drift, the work does not)
        cudaEvent_t e0, e1;
        KJ_CUDA(cudaEventCreate(&e0));
        KJ_CUDA(cudaEventCreate(&e1));
        KJ_CUDA(cudaEventRecord(e0, nullptr));
        for (int i = 0; i < iters; ++i) go();
        KJ_CUDA(cudaEventRecord(e1, nullptr));
        KJ_CUDA(cudaDeviceSynchronize());
        float ms = 0.f;
        KJ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        *us = ms * 1e3f / iters;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    cudaFree(dA); cudaFree(dW); cudaFree(dX); cudaFree(dW2); cudaFree(dO); cudaFree(dB); cudaFree(dG); cudaFree(dBt); cudaFree(dB2);
}

void dbg_ffn_ln(const uint16_t* x_bf16, const uint16_t* w1_bf16, const float* b1, const uint16_t* w2_bf16, const float* b2, const float* gamma,
                const float* beta, float eps, int M, int I, int act, uint16_t* out_bf16, int iters, float* us) {
    cudaDeviceProp prop;
    int dev = 0;
    KJ_CUDA(cudaGetDevice(&dev));
    KJ_CUDA(cudaGetDeviceProperties(&prop, dev));
    const size_t Mp = std::max(M, 128);
    __nv_bfloat16 *dX, *dW1, *dW2;
    float *dB1, *dB2, *dG, *dBt;
    KJ_CUDA(cudaMalloc(&dX, Mp * kFfH * 2));
    KJ_CUDA(cudaMemset(dX, 0, Mp * kFfH * 2));
    KJ_CUDA(cudaMalloc(&dW1, static_cast<size_t>(I) * kFfH * 2));
    KJ_CUDA(cudaMalloc(&dW2, static_cast<size_t>(I) * kFfH * 2));
    KJ_CUDA(cudaMalloc(&dB1, static_cast<size_t>(I) * 4));
    KJ_CUDA(cudaMalloc(&dB2, kFfH * 4)); KJ_CUDA(cudaMalloc(&dG, kFfH * 4)); KJ_CUDA(cudaMalloc(&dBt, kFfH * 4));
    KJ_CUDA(cudaMemcpy(dX, x_bf16, static_cast<size_t>(M) * kFfH * 2, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dW1, w1_bf16, static_cast<size_t>(I) * kFfH * 2, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dW2, w2_bf16, static_cast<size_t>(I) * kFfH * 2, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dB1, b1, static_cast<size_t>(I) * 4, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dB2, b2, kFfH * 4, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dG, gamma, kFfH * 4, cudaMemcpyHostToDevice));
    KJ_CUDA(cudaMemcpy(dBt, beta, kFfH * 4, cudaMemcpyHostToDevice));
    CUtensorMap tx = make_tmap_2d(dX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Mp, kFfH, kGemmBlockM, kGemmBlockK, 128);
    CUtensorMap t1 = make_tmap_2d(dW1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, I, kFfH, 64, kGemmBlockK, 128);
    CUtensorMap t1p = make_tmap_2d(dW1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, I, kFfH, 32, kGemmBlockK, 128);
    CUtensorMap t2 = make_tmap_2d(dW2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kFfH, I, 64, kGemmBlockK, 128);
    launch_ffn_ln(tx, t1, t1p, t2, M, I, dB1, dB2, dG, dBt, eps, act, prop.multiProcessorCount, nullptr);
    KJ_CUDA(cudaDeviceSynchronize());
    KJ_CUDA(cudaMemcpy(out_bf16, dX, static_cast<size_t>(M) * kFfH * 2, cudaMemcpyDeviceToHost));
    if (iters > 0 && us) {  // timing (in place: the values drift, the work does not)
        cudaEvent_t e0, e1;
        KJ_CUDA(cudaEventCreate(&e0));
        KJ_CUDA(cudaEventCreate(&e1));
        KJ_CUDA(cudaEventRecord(e0, nullptr));
        for (int i = 0; i < iters; ++i) launch_ffn_ln(tx, t1, t1p, t2, M, I, dB1, dB2, dG, dBt, eps, act, prop.multiProcessorCount, nullptr);
        KJ_CUDA(cudaEventRecord(e1, nullptr));
        KJ_CUDA(cudaDeviceSynchronize());
        float ms = 0.f;
        KJ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        *us = ms * 1e3f / iters;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    cudaFree(dX); cudaFree(dW1); cudaFree(dW2); cudaFree(dB1); cudaFree(dB2); cudaFree(dG); cudaFree(dBt);
}

// GEMM microbenchmark: average microseconds per launch over `iters` launches (device buffers, random-ish data).
float dbg_gemm_time(int M, int N, int K, int epi, int act, int block_n, int flags, int iters) {
    cudaDeviceProp prop;
    int dev = 0;
    KJ_CUDA(cudaGetDevice(&dev));
    KJ_CUDA(cudaGetDeviceProperties(&prop, dev));
    const bool pair2 = block_n >= 2000;
    if (pair2) block_n -= 2000;
    const bool pair = block_n >= 1000;
    if (pair) block_n -= 1000;
    const int bn = block_n > 0 ? block_n : pick_block_n(N);
    const size_t Mp = std::max(M, 128), Np = std::max(N, bn);
    __nv_bfloat16 *dA, *dW;
    float* dB;
    __nv_bfloat16* dR;
    void* dO;
    KJ_CUDA(cudaMalloc(&dA, Mp * K * 2));
    KJ_CUDA(cudaMalloc(&dW, Np * K * 2));
    KJ_CUDA(cudaMemset(dA, 0x3c, Mp * K * 2));
    KJ_CUDA(cudaMemset(dW, 0x3c, Np * K * 2));
    KJ_CUDA(cudaMalloc(&dO, Mp * N * 4));
    KJ_CUDA(cudaMalloc(&dB, static_cast<size_t>(N) * 4));
    KJ_CUDA(cudaMalloc(&dR, Mp * N * 2));
    KJ_CUDA(cudaMemset(dB, 0, static_cast<size_t>(N) * 4));
    KJ_CUDA(cudaMemset(dR, 0, Mp * N * 2));
    const bool f32out = epi == EPI_BIAS_RES_F32 || epi == EPI_BIAS_F32;
    CUtensorMap ta = make_tmap_2d(dA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Mp, K, kGemmBlockM, kGemmBlockK, 128);
    CUtensorMap tb = make_tmap_2d(dW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Np, K, bn, kGemmBlockK, 128);
    CUtensorMap tc = ta;
    if (!f32out) tc = ((!pair && gemm_wide_store(bn)) || (bn == 192 && pair)) ? make_tmap_2d(dO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Mp, N, 32, 64, 128)
                                           : make_tmap_2d(dO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Mp, N, 32, kEpiChunkCols, 64);
    GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.bias = dB; p.residual = dR; p.ldr = N; p.out = dO; p.ldo = N; p.act = act; p.dbg = flags & ~8;
    cudaEvent_t e0, e1;
    KJ_CUDA(cudaEventCreate(&e0));
    KJ_CUDA(cudaEventCreate(&e1));
    CUtensorMap tbh = make_tmap_2d(dW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Np, K, bn / 2, kGemmBlockK, 128);
    auto go = [&] {
        if (pair2) launch_gemm_cta_pair(bn, epi, ta, tbh, tc, p, prop.multiProcessorCount, nullptr);
        else if (pair) launch_gemm_pair(bn, epi, ta, tbh, tc, p, prop.multiProcessorCount, nullptr);
        else launch_gemm(bn, epi, ta, tb, tc, p, prop.multiProcessorCount, nullptr);
    };
    for (int i = 0; i < 5; ++i) go();
    if (flags & 8) {  // one traced launch: per-CTA milestone stamps (ns, relative to the earliest kernel entry)
        unsigned long long* dT;
        const int ctas = prop.multiProcessorCount;
        KJ_CUDA(cudaMalloc(&dT, ctas * 32 * 8));
        KJ_CUDA(cudaMemset(dT, 0, ctas * 32 * 8));
        p.trace = dT;
        p.dbg = flags & ~8;
        go(); go(); go();
        KJ_CUDA(cudaDeviceSynchronize());
        std::vector<unsigned long long> h(ctas * 32);
        KJ_CUDA(cudaMemcpy(h.data(), dT, h.size() * 8, cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull;
        for (int c = 0; c < ctas; ++c) if (h[c * 32]) t0 = std::min(t0, h[c * 32]);
        static const char* names[18] = {"entry", "prologue", "pdl_wait", "operands0", "acc0", "epi0", "acc1", "epi1", "acc2", "epi2", "acc3", "epi3",
                                        "acc4", "epi4", "acc5", "epi5", "drained", "exit"};
        for (int c : {0, 1, ctas / 2, ctas - 1}) {
            fprintf(stderr, "trace M=%d N=%d K=%d bn=%d cta %3d:", M, N, K, bn, c);
            for (int i = 0; i < 18; ++i) if (h[c * 32 + i]) fprintf(stderr, " %s=%.2f", names[i], (h[c * 32 + i] - t0) * 1e-3);
            if (h[c * 32 + 22]) fprintf(stderr, " | epi(warp5,tile2): start=%.2f wait_read=+%.2f ld0=+%.2f ld1=+%.2f math+sts=+%.2f store=+%.2f", (h[c * 32 + 22] - t0) * 1e-3,
                                        (h[c * 32 + 23] - h[c * 32 + 22]) * 1e-3, (h[c * 32 + 24] - h[c * 32 + 23]) * 1e-3, (h[c * 32 + 25] - h[c * 32 + 24]) * 1e-3,
                                        (h[c * 32 + 26] - h[c * 32 + 25]) * 1e-3, (h[c * 32 + 27] - h[c * 32 + 26]) * 1e-3);
            if (h[c * 32 + 16] > h[c * 32 + 2]) fprintf(stderr, " sm_mhz=%.0f", (h[c * 32 + 21] - h[c * 32 + 20]) * 1e3 / double(h[c * 32 + 16] - h[c * 32 + 2]));
            fprintf(stderr, "\n");
        }
        p.trace = nullptr;
        p.dbg = flags & ~8;
        cudaFree(dT);
    }
    KJ_CUDA(cudaEventRecord(e0, nullptr));
    for (int i = 0; i < iters; ++i) go();
    KJ_CUDA(cudaEventRecord(e1, nullptr));
    KJ_CUDA(cudaDeviceSynchronize());
    float ms = 0.f;
    KJ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dA); cudaFree(dW); cudaFree(dO); cudaFree(dB); cudaFree(dR);
    return ms * 1e3f / iters;
}

void dbg_attention(const uint16_t* qkv_bf16, const float* mask, int B, int S, int H, int heads, int nan_if_all_masked, uint16_t* ctx_bf16) {
                const size_t T = static_cast<size_t>(B) * S;
        __nv_bfloat16 *dq, *dc;
        float* dm = nullptr;
        KJ_CUDA(cudaMalloc(&dq, T * 3 * H * 2));
        KJ_CUDA(cudaMalloc(&dc, T * H * 2));
        KJ_CUDA(cudaMemcpy(dq, qkv_bf16, T * 3 * H * 2, cudaMemcpyHostToDevice));
        if (mask) { KJ_CUDA(cudaMalloc(&dm, T * 4)); KJ_CUDA(cudaMemcpy(dm, mask, T * 4, cudaMemcpyHostToDevice)); }
        AttnParams a;
        const int d = H / heads;
        a.qkv = dq; a.mask = dm; a.ctx = dc; a.B = B; a.S = S; a.H = H; a.heads = heads;
        a.scale_log2e = (1.0f / sqrtf(static_cast<float>(d))) * 1.4426950408889634f;
        a.nan_if_all_masked = nan_if_all_masked;
        if (const char* e = getenv("KJC_ATTN_DBG")) a.dbg = atoi(e);
        launch_attention(a, d, nullptr);
        KJ_CUDA(cudaDeviceSynchronize());
        KJ_CUDA(cudaMemcpy(ctx_bf16, dc, T * H * 2, cudaMemcpyDeviceToHost));
        if (getenv("KJC_ATTN_TRACE")) {  // per-unit milestones of the softmax warpgroups + average launch time
            cudaDeviceProp prop;
            int dev = 0;
            KJ_CUDA(cudaGetDevice(&dev));
            KJ_CUDA(cudaGetDeviceProperties(&prop, dev));
            unsigned long long* dT;
            const int ctas = prop.multiProcessorCount;
            KJ_CUDA(cudaMalloc(&dT, ctas * 64 * 8));
            KJ_CUDA(cudaMemset(dT, 0, ctas * 64 * 8));
            for (int i = 0; i < 3; ++i) launch_attention(a, d, nullptr);
            unsigned long long* dT2;
            KJ_CUDA(cudaMalloc(&dT2, ctas * 64 * 8));
            KJ_CUDA(cudaMemset(dT2, 0, ctas * 64 * 8));
            a.trace = dT;
            launch_attention(a, d, nullptr);
            a.trace = dT2;
            launch_attention(a, d, nullptr);
            a.trace = nullptr;
            launch_attention(a, d, nullptr);
            KJ_CUDA(cudaDeviceSynchronize());
            {
                std::vector<unsigned long long> h1(ctas * 64), h2(ctas * 64);
                KJ_CUDA(cudaMemcpy(h1.data(), dT, h1.size() * 8, cudaMemcpyDeviceToHost));
                KJ_CUDA(cudaMemcpy(h2.data(), dT2, h2.size() * 8, cudaMemcpyDeviceToHost));
                unsigned long long t0 = ~0ull, e1max = 0, n2min = ~0ull, x2max = 0, e1min = ~0ull;
                for (int c = 0; c < ctas; ++c) {
                    if (!h1[c * 64 + 60]) continue;
                    t0 = std::min(t0, h1[c * 64 + 60]);
                    e1max = std::max(e1max, h1[c * 64 + 62]);
                    e1min = std::min(e1min, h1[c * 64 + 62]);
                    n2min = std::min(n2min, h2[c * 64 + 60]);
                    x2max = std::max(x2max, h2[c * 64 + 62]);
                }
                if (t0 != ~0ull)
                    fprintf(stderr, "  two launches back to back (us since first entry): launch1 exits %.2f .. %.2f | launch2 first entry %.2f, last exit %.2f\n",
                            (e1min - t0) * 1e-3, (e1max - t0) * 1e-3, (n2min - t0) * 1e-3, (x2max - t0) * 1e-3);
            }
            cudaFree(dT2);
            cudaEvent_t e0, e1;
            KJ_CUDA(cudaEventCreate(&e0));
            KJ_CUDA(cudaEventCreate(&e1));
            KJ_CUDA(cudaEventRecord(e0, nullptr));
            for (int i = 0; i < 20; ++i) launch_attention(a, d, nullptr);
            KJ_CUDA(cudaEventRecord(e1, nullptr));
            KJ_CUDA(cudaDeviceSynchronize());
            float ms = 0.f;
            KJ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            std::vector<unsigned long long> h(ctas * 64);
            KJ_CUDA(cudaMemcpy(h.data(), dT, h.size() * 8, cudaMemcpyDeviceToHost));
            fprintf(stderr, "attention B=%d S=%d H=%d heads=%d: %.1f us/launch\n", B, S, H, heads, ms * 1e3f / 20);
            const bool ts = attention_variant() == 0;  // attention_ts.cuh: 4 warpgroups x units 2, 3 at [wg * 16 + u * 6]
            for (int c : {0, ctas / 2}) {
                if (ts && h[c * 64 + 60]) {
                    fprintf(stderr, "  cta %3d: entry=0 prologue_done=+%.2f exit=+%.2f (us)\n", c, (h[c * 64 + 61] - h[c * 64 + 60]) * 1e-3,
                            (h[c * 64 + 62] - h[c * 64 + 60]) * 1e-3);
                    fprintf(stderr, "    QK issue:");
                    for (int g = 0; g < 8; ++g) if (h[c * 64 + 24 + g]) fprintf(stderr, " %.2f", (h[c * 64 + 24 + g] - h[c * 64 + 60]) * 1e-3);
                    fprintf(stderr, "\n    PV issue:");
                    for (int g = 0; g < 8; ++g) if (h[c * 64 + 32 + g]) fprintf(stderr, " %.2f", (h[c * 64 + 32 + g] - h[c * 64 + 60]) * 1e-3);
                    fprintf(stderr, "\n");
                }
                for (int wgi = 0; wgi < (ts ? 4 : 3); ++wgi)
                    for (int u = 0; u < (ts ? 1 : 3); ++u) {
                        const unsigned long long* t = &h[c * 64 + wgi * (ts ? 6 : 20) + u * 6];
                        if (!t[0]) continue;
                        fprintf(stderr, "  cta %3d wg %d unit %d: top=%.2f wait_s=+%.2f pass1=+%.2f pass2=+%.2f wait_o=+%.2f epi=+%.2f (us)\n", c, wgi, u + 1,
                                (t[0] - h[c * 64 + (ts ? 60 : 0)]) * 1e-3, (t[1] - t[0]) * 1e-3, (t[2] - t[1]) * 1e-3, (t[3] - t[2]) * 1e-3, (t[4] - t[3]) * 1e-3, (t[5] - t[4]) * 1e-3);
                    }
            }
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            cudaFree(dT);
        }
        cudaFree(dq); cudaFree(dc);
        if (dm) cudaFree(dm);
}

}  // namespace kj
