/* kjarni_cuda_debug.h -- single-kernel test hooks exported by libkjarni_cuda.so.
 * Not part of the drop-in boundary; used by tests/ to check kernels in isolation
 * against the oracle.  Host buffers; bf16 tensors are passed as raw uint16. */
#ifndef KJARNI_CUDA_DEBUG_H
#define KJARNI_CUDA_DEBUG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* out = epilogue(A[M,K] x W[N,K]^T): epi 0 = bias->bf16, 1 = act(bias)->bf16, 2 = bias+residual->f32, 3 = bias->f32;
 * act 0 = erf-GELU, 1 = tanh-GELU, 2 = ReLU, 3 = none; block_n 0 = auto (64/128/192/256). */
int kjc_dbg_gemm(const uint16_t* a_bf16, const uint16_t* w_bf16, const float* bias, const float* residual, int M, int N, int K, int epi,
                 int act, int block_n, void* out);
/* out[M,384] (bf16) = LayerNorm(A[M,K] x W[384,K]^T + bias + residual[M,384] (bf16)) with the fused kernel;
 * iters > 0 additionally times `iters` launches (average us in *out_us). */
int kjc_dbg_gemm_ln(const uint16_t* a_bf16, const uint16_t* w_bf16, const float* bias, const float* gamma, const float* beta, float eps,
                    const uint16_t* res_bf16, int M, int K, uint16_t* out_bf16, int iters, float* out_us);
/* The same for hidden size H = 384 (one CTA per 128-row tile) or 768 (a CTA pair per tile, statistics exchanged through DSMEM). */
int kjc_dbg_gemm_ln_h(const uint16_t* a_bf16, const uint16_t* w_bf16, const float* bias, const float* gamma, const float* beta, float eps,
                      const uint16_t* res_bf16, int M, int H, int K, uint16_t* out_bf16, int iters, float* out_us);
/* The chained kernel alone: out_x[M,384] = LayerNorm(A[M,K1] x W1[384,K1]^T + bias1 + res), out2[M,N2] = epi2(out_x x W2[N2,384]^T + bias2)
 * (epi2: 0 = bias->bf16, 1 = act(bias)->bf16); M <= 128 * number of SMs, N2 <= 1536; iters > 0 additionally times `iters` launches. */
int kjc_dbg_gemm_ln_gemm(const uint16_t* a_bf16, const uint16_t* w1_bf16, const float* bias1, const float* gamma, const float* beta, float eps,
                         const uint16_t* res_bf16, int M, int K1, const uint16_t* w2_bf16, const float* bias2, int N2, int epi2, int act,
                         uint16_t* out_x_bf16, uint16_t* out2_bf16, int iters, float* out_us);
/* 1 if the library was built with -DKJ_EXPERIMENTAL_KERNELS (the CTA-pair GEMM and the whole-FFN fusion, both slower than the
 * default path, are then available to kjc_dbg_gemm(block_n + 1000) / kjc_dbg_ffn_ln / KJC_PAIR_GEMM / KJC_FUSED_FFN); else those
 * entries return KJC_INVALID_CONFIG. */
int kjc_dbg_experimental_kernels(void);
/* out[M,384] (bf16) = LayerNorm(x + act(x W1^T + b1) W2^T + b2) with the fused feed-forward kernel; W1 [I,384], W2 [384,I];
 * iters > 0 additionally times `iters` launches (average us in *out_us). */
int kjc_dbg_ffn_ln(const uint16_t* x_bf16, const uint16_t* w1_bf16, const float* b1, const uint16_t* w2_bf16, const float* b2, const float* gamma,
                   const float* beta, float eps, int M, int I, int act, uint16_t* out_bf16, int iters, float* out_us);
/* GEMM microbenchmark: average us per launch. flags: 1 = skip epilogue work, 2 = skip MMA issue, 4 = skip TMA loads. */
int kjc_dbg_gemm_time(int M, int N, int K, int epi, int act, int block_n, int flags, int iters, float* out_us);
/* ctx[B*S,H] = attention(qkv[B*S,3H], mask[B,S]) with the fused kernel. */
int kjc_dbg_attention(const uint16_t* qkv_bf16, const float* mask, int B, int S, int H, int heads, int nan_if_all_masked, uint16_t* ctx_bf16);
/* Index filter knobs: `eps` = proof margin on |approximate - exact cosine| (default 0.0045; a huge value forces every
 * query through the exact re-run), `min_queries` = smallest batch that takes the tensor-core filter path (default 1; a huge value forces the exact scan). */
struct KjcIndex;
int kjc_dbg_index_set_filter(struct KjcIndex* idx, float eps, int min_queries);
/* logits[B,C] = classification head of `enc` applied to caller-supplied fp32 hidden states [B,S,H]. */
struct KjcEncoder;
int kjc_dbg_encoder_head(struct KjcEncoder* enc, const float* hidden, int batch, int seq_len, float* logits);
#ifdef __cplusplus
}
#endif
#endif
