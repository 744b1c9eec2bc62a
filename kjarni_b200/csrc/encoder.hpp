// Encoder handle behind KjcEncoder (see include/kjarni_cuda.h and encoder.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "host_util.hpp"

namespace kj {

struct GemmParams;
struct AttnParams;

CUtensorMap make_tmap_2d(const void* base, CUtensorMapDataType dt, int elem_bytes, uint64_t rows, uint64_t cols, uint32_t box_rows,
                         uint32_t box_cols, int swizzle_bytes);
int pick_block_n(int N);
void launch_gemm(int block_n, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p, int num_sms,
                 cudaStream_t st);
bool experimental_kernels_built();
void launch_gemm_pair(int block_n, int epi, const CUtensorMap& ta, const CUtensorMap& tb_half, const CUtensorMap& tc, const GemmParams& p,
                      int num_sms, cudaStream_t st);
void launch_ffn_ln(const CUtensorMap& t_x, const CUtensorMap& t_w1, const CUtensorMap& t_w1_pair, const CUtensorMap& t_w2, int M, int I,
                   const float* b1, const float* b2, const float* gamma, const float* beta, float eps, int act, int num_sms, cudaStream_t st);
void dbg_ffn_ln(const uint16_t* x_bf16, const uint16_t* w1_bf16, const float* b1, const uint16_t* w2_bf16, const float* b2, const float* gamma,
                const float* beta, float eps, int M, int I, int act, uint16_t* out_bf16, int iters, float* us);
void launch_gemm_ln(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& t_io, int M, int H, int K, const float* bias, const float* gamma,
                    const float* beta, float eps, int num_sms, cudaStream_t st);
void launch_attention(const AttnParams& p, int head_dim, cudaStream_t st);
void launch_layernorm(const float* y, const float* g, const float* b, float eps, float* x32, __nv_bfloat16* x16, int M, int H,
                      cudaStream_t st);

void dbg_gemm(const uint16_t* a_bf16, const uint16_t* w_bf16, const float* bias, const float* residual, int M, int N, int K, int epi,
              int act, int block_n, void* out);
void dbg_gemm_ln_gemm(const uint16_t* a_bf16, const uint16_t* w1_bf16, const float* bias1, const float* gamma, const float* beta, float eps,
                      const uint16_t* res_bf16, int M, int K1, const uint16_t* w2_bf16, const float* bias2, int N2, int epi2, int act,
                      uint16_t* out_x_bf16, uint16_t* out2_bf16, int iters, float* us);
void dbg_gemm_ln(const uint16_t* a_bf16, const uint16_t* w_bf16, const float* bias, const float* gamma, const float* beta, float eps,
                 const uint16_t* res_bf16, int M, int H, int K, uint16_t* out_bf16, int iters, float* us);
float dbg_gemm_time(int M, int N, int K, int epi, int act, int block_n, int flags, int iters);
void dbg_attention(const uint16_t* qkv_bf16, const float* mask, int B, int S, int H, int heads, int nan_if_all_masked, uint16_t* ctx_bf16);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember what each device was configured with.
template <typename K>
inline void ensure_smem_attr(K kern, int bytes, int (&configured)[64]) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured[dev & 63] >= bytes) return;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess)
        throw Error(KJC_INFERENCE_FAILED, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
    configured[dev & 63] = bytes;
}

// Launch with programmatic stream serialization (PDL): the kernel must call pdl_wait() before it touches global memory its
// predecessor wrote.  KJC_NO_PDL=1 falls back to plain launches.
template <typename... KArgs, typename... Args>
inline void launch_pdl_cluster(int cluster_x, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    static const bool enabled = getenv("KJC_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (enabled) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster_x > 1) {  // thread-block cluster of `cluster_x` CTAs along x
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster_x;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    KJ_CUDA(cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...));
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    launch_pdl_cluster(1, kern, grid, block, smem, st, std::forward<Args>(args)...);
}

struct LayerDev {
    const __nv_bfloat16 *wqkv, *wo, *w1, *w2;
    const float *bqkv, *bo, *b1, *b2, *g1, *be1, *g2, *be2;
    CUtensorMap t_wqkv, t_wo, t_w1, t_w2;
    CUtensorMap t_w1_ffn, t_w1_ffn32, t_w2_ffn;  // fused FFN kernel: W1 boxes of 64 (one CTA) / 32 (CTA pair) rows, W2 boxes of 64 rows
    CUtensorMap t_wo_ln, t_w2_ln;        // 192-row boxes: weights of the GEMM + residual + LayerNorm kernels (hidden 384 / 768)
    CUtensorMap t_wo_96, t_w2_96, t_w1_96, t_wqkv_96;  // 96-row boxes: the CTA-pair chained kernels load half a weight tile per CTA
    CUtensorMap t_w1_192, t_wqkv_192;    // 192-row boxes: phase-2 weights of the chained GEMM+LN -> GEMM kernel
    CUtensorMap t_w1_128, t_wqkv_128;    // 128-row boxes: the same with x' in tensor memory (kTS)
    CUtensorMap t_wqkv_half, t_w1_half;  // box of block_n/2 rows: the CTA-pair GEMM loads half a weight tile per CTA
};

class Encoder {
  public:
    Encoder(const std::string& model_dir, int device);
    ~Encoder();
    Encoder(const Encoder&) = delete;
    Encoder& operator=(const Encoder&) = delete;

    const KjcEncoderInfo& info() const { return info_; }
    const std::vector<std::string>& labels() const { return labels_; }
    int micro_batch(int seq_len) const;
    int64_t last_launches() const { return launches_; }
    bool chained() const { return chain_; }
    // config.json's max_position_embeddings (meta.max_seq_len, pipeline/encoder/loader.rs:108); info().max_position_embeddings is
    // the number of ROWS of the position table, which is what bounds the gather (embeddings/mod.rs:199-214)
    int config_max_seq_len() const { return cfg_max_seq_len_; }
    // Residual stream precision: 0 = bf16 between kernels for every output (fastest), 1 = fp32 for KJC_OUT_HIDDEN only (default: the
    // hidden states then carry only the bf16 rounding of the GEMM operands, max-abs error < 2e-2 against the fp32 reference), 2 = fp32
    // for every output.  The fp32 mode runs the un-chained kernels: GEMM -> fp32 sums -> LayerNorm kernel.
    void set_fp32_residual(int mode) { fp32_residual_ = mode < 0 ? 0 : (mode > 2 ? 2 : mode); }
    int fp32_residual() const { return fp32_residual_; }
    // Per-kernel-class CUDA-event timing of the forwards issued while profiling is on (bench roofline numbers).
    void set_profiling(bool on);
    void get_profile(double* ms, int64_t* launches);  // arrays of KJC_NUM_KERNEL_CLASSES; synchronises

    // `sink` (optional): instead of copying the result rows into `out`, each finished block of rows is handed over while it still
    // sits in the pinned staging buffer -- sink(rows, first_row, n_rows) -- e.g. to be written to a file (indexer: vectors.bin).
    using RowSink = std::function<void(const float* rows, size_t first_row, size_t n_rows)>;
    void forward_host(const uint32_t* ids, const float* mask, const uint32_t* types, int B, int S, const KjcForwardOptions& o, float* out,
                      const RowSink* sink = nullptr);
    void head_only_host(const float* hidden, int B, int S, float* logits);
    void forward_device(const uint32_t* d_ids, const float* d_mask, const uint32_t* d_types, int B, int S, const KjcForwardOptions& o,
                        float* d_out, cudaStream_t st);

    // `o` with KJC_MASK_AUTO replaced by the convention the reference's ComputeStrategy picks for a batch of B x S tokens (validates
    // first).  A caller that splits one batch over several replicas resolves once for the whole batch and passes the result down.
    KjcForwardOptions resolved_options(int B, int S, const KjcForwardOptions& o) const {
        validate(B, S, o);
        KjcForwardOptions oc = o;
        oc.mask_convention = resolve_noalloc(B, S, o) ? KJC_MASK_NOALLOC : KJC_MASK_ALLOC;
        return oc;
    }
    size_t out_row_elems(const KjcForwardOptions& o, int S) const;

  private:
    void validate(int B, int S, const KjcForwardOptions& o) const;
    bool resolve_noalloc(int B, int S, const KjcForwardOptions& o) const;
    // Activations of one micro-batch in flight.  Several "lanes" run concurrently on disjoint groups of SMs (each kernel is
    // launched with num_sms / lanes persistent CTAs on the lane's own stream), so the fixed per-launch latency of one lane
    // (launch, prologue, pipeline fill, epilogue drain) is covered by the other lanes' tensor work.
    struct Workspace {
        int tokens = 0;
        float* y32 = nullptr;  // unfused path only: pre-LayerNorm sums
        float* x32 = nullptr;  // fp32-residual mode: the residual stream in fp32 (x16 is its bf16 copy for the tensor-core operands)
        float* head32 = nullptr;  // classification head: pre-classifier output [sequences, H]
        __nv_bfloat16 *x16 = nullptr, *qkv16 = nullptr, *ctx16 = nullptr, *h16 = nullptr;
        int qkv_ld = 0;        // row pitch of qkv16 in elements: 3H, or I when qkv16 aliases h16 (see ensure_workspace)
        CUtensorMap t_x16, t_ctx16, t_h16;   // A-operand loads
        CUtensorMap t_qkv16_out, t_h16_out;  // epilogue TMA stores
        CUtensorMap t_qkv16_out32, t_h16_out32;  // 32-column store boxes (CTA-pair kernel)
        CUtensorMap t_qkv16_out64, t_h16_out64;  // 64-column store boxes (chained kernels)
        CUtensorMap t_x16_io;                // residual load + LayerNorm output store of the fused kernel
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
    };
    void ensure_workspace(Workspace& w, int tokens);
    void ensure_fp32_stream(Workspace& w);
    void free_workspace(Workspace& w);
    void forward_micro(Workspace& w, int sms, const uint32_t* d_ids, const float* d_mask, const uint32_t* d_types, int nb, int S,
                       const KjcForwardOptions& o, bool noalloc_convention, float* d_out, cudaStream_t st);
    bool fp32_residual_for(const KjcForwardOptions& o) const { return fp32_residual_ == 2 || (fp32_residual_ == 1 && o.output == KJC_OUT_HIDDEN); }
    void forward_batches(const uint32_t* d_ids, const float* d_mask, const uint32_t* d_types, int B, int S, const KjcForwardOptions& o,
                         float* d_out, cudaStream_t st);

    KjcEncoderInfo info_{};
    std::vector<std::string> labels_;
    int num_sms_ = 0, act_ = 0, micro_tokens_ = 0;
    bool alias_qkv_ = true;    // qkv rows laid over the FFN activation rows (KJC_NO_ALIAS_QKV)
    int chain_min_tiles_ = 0;  // micro-batches of fewer 128-row tiles take one kernel per op (KJC_CHAIN_MIN_TILES)
    int bn_qkv_ = 0, bn_h_ = 0, bn_i_ = 0;
    std::mutex mu_;
    cudaStream_t stream_ = nullptr;
    // parameters
    float* d_f32_ = nullptr;
    __nv_bfloat16* d_w16_ = nullptr;
    const float *word_ = nullptr, *pos_ = nullptr, *type_ = nullptr, *emb_g_ = nullptr, *emb_b_ = nullptr;
    const float *w_pre_ = nullptr, *b_pre_ = nullptr, *w_cls_ = nullptr, *b_cls_ = nullptr;
    std::vector<LayerDev> layers_;
    int fp32_residual_ = 1;
    int cfg_max_seq_len_ = 0;
    bool chain_pair_ = false;
    int chain_pair_mask_ = 0;
    bool last_ln_pair_ = false;
    int gemm_pair_mask_ = 0;  // bit 0: QKV, bit 1: FFN-up run as CTA pairs (gemm_tcgen05_kernel<BN, EPI, true>)
    bool chain_ts_ = false;
    bool fused_ln_ = false, pair_gemm_ = false, fused_ffn_ = false, chain_ = false, chain_embed_ = false;
    int lanes_ = 2;
    std::vector<Workspace> ws_;
    cudaEvent_t ev_in_ = nullptr;
    // host-buffer entry point staging
    uint32_t* d_in_ = nullptr;
    float* d_out_ = nullptr;
    uint32_t* h_stage_in_ = nullptr;
    float* h_stage_out_ = nullptr;
    size_t in_cap_ = 0, out_cap_ = 0;
    int* d_err_ = nullptr;
    int err_host_ = 0;
    // forward_host: copy-in / copy-out streams and per-chunk events (inputs landed, compute done, results in pinned memory)
    cudaStream_t s_in_ = nullptr, s_out_ = nullptr;
    std::vector<cudaEvent_t> ev_in_chunk_, ev_done_chunk_, ev_out_chunk_;
    int64_t launches_ = 0;
    // profiling
    struct ProfRec { int cls; cudaEvent_t a, b; };
    bool profiling_ = false;
    std::vector<ProfRec> prof_recs_;
    std::vector<cudaEvent_t> prof_pool_;
    double prof_ms_[KJC_NUM_KERNEL_CLASSES] = {0};
    int64_t prof_n_[KJC_NUM_KERNEL_CLASSES] = {0};
    cudaEvent_t prof_event();
    void prof_begin(int cls, cudaStream_t st);
    void prof_end(cudaStream_t st);
    void prof_collect();
};

}  // namespace kj
