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
/* ctx[B*S,H] = attention(qkv[B*S,3H], mask[B,S]) with the fused kernel. */
int kjc_dbg_attention(const uint16_t* qkv_bf16, const float* mask, int B, int S, int H, int heads, int nan_if_all_masked, uint16_t* ctx_bf16);
/* logits[B,C] = classification head of `enc` applied to caller-supplied fp32 hidden states [B,S,H]. */
struct KjcEncoder;
int kjc_dbg_encoder_head(struct KjcEncoder* enc, const float* hidden, int batch, int seq_len, float* logits);
#ifdef __cplusplus
}
#endif
#endif
