// Bandwidth-bound row kernels of the encoder path: fused embedding gather +
// LayerNorm, LayerNorm, masked mean-pool + L2 normalise, CLS pooling and the
// classification head.  fp32 statistics everywhere (layer_norm_eps = 1e-12 is
// below bf16 resolution); one warp per token row, rows held in registers.
#pragma once
#include "ptx.cuh"

namespace kj {

constexpr int kRowThreads = 256;  // 8 rows per CTA

// 4 consecutive activations as fp32, from either fp32 or bf16 storage.
__device__ __forceinline__ float4 load4f(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4f(const __nv_bfloat16* p) {
    const uint2 w = *reinterpret_cast<const uint2*>(p);
    return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16),
                       __uint_as_float(w.y & 0xffff0000u));
}

// LayerNorm over H held as float4 chunks: biased variance, eps inside the sqrt
// (reference: kjarni-transformers/src/cpu/normalization/layer_norm.rs:37-134).
template <int NV>  // float4 chunks per lane
__device__ __forceinline__ void warp_layernorm_store(float4 (&v)[NV], int H, int lane, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, float* __restrict__ out32,
                                                     __nv_bfloat16* __restrict__ out16) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
        if ((lane + 32 * i) * 4 < H) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / static_cast<float>(H);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
        if ((lane + 32 * i) * 4 < H) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    const float inv_std = 1.0f / sqrtf(warp_sum(q) / static_cast<float>(H) + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 4;
        if (c < H) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
            const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + c));
            float4 y;
            y.x = (v[i].x - mean) * inv_std * g.x + bt.x;
            y.y = (v[i].y - mean) * inv_std * g.y + bt.y;
            y.z = (v[i].z - mean) * inv_std * g.z + bt.z;
            y.w = (v[i].w - mean) * inv_std * g.w + bt.w;
            if (out32) *reinterpret_cast<float4*>(out32 + c) = y;
            if (out16) {
                uint2 w;
                w.x = pack_bf16(y.x, y.y);
                w.y = pack_bf16(y.z, y.w);
                *reinterpret_cast<uint2*>(out16 + c) = w;
            }
        }
    }
}

struct EmbedParams {
    const uint32_t* ids;       // [B*S]
    const uint32_t* type_ids;  // [B*S] or nullptr (row 0 of the type table is added to every token)
    const float* word;         // [vocab, H]
    const float* pos;          // [max_pos, H] or nullptr
    const float* type;         // [type_vocab, H] or nullptr
    const float* gamma;
    const float* beta;
    float* x32;                // [B*S, H]
    __nv_bfloat16* x16;        // [B*S, H]
    int* err_flag;             // set to 1 on a token-type id out of range (the reference panics)
    int M, S, H, vocab, max_pos, type_vocab, pos_offset;
    float eps;
};

// Embeddings::forward + embed LayerNorm (reference: cpu/embeddings/mod.rs:181-326,
// cpu/encoder/transformer_encoder.rs:303-305): word[id] (zero row if id >= vocab)
// + pos[offset + s] (skipped beyond the table) + type[tt] (row 0 if no type ids).
template <int NV>
__global__ void __launch_bounds__(kRowThreads) embed_layernorm_kernel(EmbedParams p) {
    pdl_wait();               // launched with programmatic stream serialization: x16 may still be read by the previous micro-batch
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (kRowThreads / 32) + warp;
    if (row >= p.M) return;
    const int s = row % p.S;
    const uint32_t id = p.ids[row];
    const bool has_word = id < static_cast<uint32_t>(p.vocab);
    const int pidx = p.pos_offset + s;
    const bool has_pos = p.pos != nullptr && pidx < p.max_pos;
    uint32_t tt = 0;
    const bool has_type = p.type != nullptr && p.type_vocab > 0;
    if (has_type && p.type_ids != nullptr) {
        tt = p.type_ids[row];
        if (tt >= static_cast<uint32_t>(p.type_vocab)) {
            if (lane == 0) *p.err_flag = 1;
            tt = 0;
        }
    }
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < p.H) {
            if (has_word) a = __ldg(reinterpret_cast<const float4*>(p.word + static_cast<size_t>(id) * p.H + c));
            if (has_pos) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(pidx) * p.H + c));
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            if (has_type) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.type + static_cast<size_t>(tt) * p.H + c));
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
        }
        v[i] = a;
    }
    warp_layernorm_store<NV>(v, p.H, lane, p.gamma, p.beta, p.eps, p.x32 ? p.x32 + static_cast<size_t>(row) * p.H : nullptr,
                             p.x16 ? p.x16 + static_cast<size_t>(row) * p.H : nullptr);
}

// x = LN(y) where y already holds residual + projection (+bias) from the GEMM epilogue.
template <int NV>
__global__ void __launch_bounds__(kRowThreads)
layernorm_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                 float* __restrict__ x32, __nv_bfloat16* __restrict__ x16, int M, int H) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (kRowThreads / 32) + warp;
    if (row >= M) return;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 4;
        v[i] = (c < H) ? __ldcs(reinterpret_cast<const float4*>(y + static_cast<size_t>(row) * H + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    warp_layernorm_store<NV>(v, H, lane, gamma, beta, eps, x32 ? x32 + static_cast<size_t>(row) * H : nullptr,
                             x16 ? x16 + static_cast<size_t>(row) * H : nullptr);
}

// Pooling modes (reference: kjarni-transformers/src/pooling/mod.rs:11-68 and
// cpu/encoder/traits.rs:204-225,529-536).
enum PoolMode : int { POOL_MEAN = 0, POOL_CLS = 1, POOL_MAX = 2, POOL_LAST = 3 };

// One CTA per sequence, one thread per 4 hidden columns; fused optional L2 normalise.
template <typename TIn>
__global__ void __launch_bounds__(256)
pool_l2_kernel(const TIn* __restrict__ x, const float* __restrict__ mask, float* __restrict__ out, int S, int H, int mode,
               int normalize) {
    const int b = blockIdx.x;
    const TIn* xb = x + static_cast<size_t>(b) * S * H;
    const float* mb = mask ? mask + static_cast<size_t>(b) * S : nullptr;
    __shared__ float red[8];
    __shared__ float s_count;
    __shared__ int s_last;
    const int tid = threadIdx.x;
    if (tid < 32) {
        float cnt = 0.f;
        int last = -1;
        for (int s = tid; s < S; s += 32) {
            const float m = mb ? mb[s] : 1.0f;
            cnt += m;
            if (m > 0.0f) last = s;
        }
        cnt = warp_sum(cnt);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        if (tid == 0) {
            s_count = cnt;
            s_last = last < 0 ? 0 : last;
        }
    }
    __syncthreads();
    const float count = s_count;
    float sq = 0.0f;
    float4 acc[2];  // supports H <= 2048
    for (int it = 0; it < 2; ++it) {
        const int c = (tid + it * 256) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < H) {
            if (mode == POOL_CLS || (mode == POOL_MEAN && count == 0.0f)) {
                a = load4f(xb + c);
            } else if (mode == POOL_LAST) {
                a = load4f(xb + static_cast<size_t>(s_last) * H + c);
            } else if (mode == POOL_MEAN) {
                for (int s = 0; s < S; ++s) {
                    const float m = mb ? mb[s] : 1.0f;
                    if (m != 0.0f) {
                        const float4 t = load4f(xb + static_cast<size_t>(s) * H + c);
                        a.x += t.x * m; a.y += t.y * m; a.z += t.z * m; a.w += t.w * m;
                    }
                }
                a.x /= count; a.y /= count; a.z /= count; a.w /= count;
            } else {  // POOL_MAX: masked positions become -1e9 (MASK_VALUE), result clamped below by it
                a = make_float4(-1e9f, -1e9f, -1e9f, -1e9f);
                for (int s = 0; s < S; ++s) {
                    const float m = mb ? mb[s] : 1.0f;
                    if (m != 0.0f) {
                        const float4 t = load4f(xb + static_cast<size_t>(s) * H + c);
                        a.x = fmaxf(a.x, t.x); a.y = fmaxf(a.y, t.y); a.z = fmaxf(a.z, t.z); a.w = fmaxf(a.w, t.w);
                    }
                }
            }
            sq += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
        }
        acc[it] = a;
    }
    float scale = 1.0f;
    if (normalize) {
        sq = warp_sum(sq);
        if ((tid & 31) == 0) red[tid >> 5] = sq;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += red[i];
        const float nrm = sqrtf(tot);
        if (nrm > 0.0f) scale = 1.0f / nrm;  // l2_normalize_inplace: only if norm > 0 (NaN norm leaves the row as is)
    }
    for (int it = 0; it < 2; ++it) {
        const int c = (tid + it * 256) * 4;
        if (c < H) {
            float4 a = acc[it];
            if (scale != 1.0f) { a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale; }
            *reinterpret_cast<float4*>(out + static_cast<size_t>(b) * H + c) = a;
        }
    }
}

// Masked mean pool + optional L2 normalise, the Embedder's default (reference: pooling/mod.rs:11-34,
// cpu/encoder/traits.rs:529-536), with the rows of a sequence spread over the 8 warps: warp w sums rows w, w+8, ...
// with coalesced 8-byte-per-lane loads (many independent loads in flight), partial sums meet in shared memory.
// One CTA per sequence; H <= 1024 and a multiple of 4.  Same arithmetic as pool_l2_kernel up to summation order.
// 16 warps per sequence: at S = 128 every warp issues its 8 rows in two trips of four, so a CTA is two memory round trips long
constexpr int kPoolWarps = 16;
constexpr int kPoolThreads = kPoolWarps * 32;
template <int NCH>  // float4 column chunks per lane: ceil(H / 128)
__global__ void __launch_bounds__(kPoolThreads)
mean_pool_l2_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ mask, float* __restrict__ out, int S, int H,
                    int normalize) {
    pdl_wait();  // launched with programmatic stream serialization
    pdl_launch_dependents();
    extern __shared__ float s_part[];  // [kPoolWarps][H]
    __shared__ float s_cnt[kPoolWarps];
    __shared__ float red[kPoolWarps];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const __nv_bfloat16* xb = x + static_cast<size_t>(b) * S * H;
    const float* mb = mask ? mask + static_cast<size_t>(b) * S : nullptr;
    float4 acc[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float cnt = 0.f;
    // four rows per trip with every load issued before the first use: the kernel is a latency chain (one CTA per sequence),
    // so memory-level parallelism is what sets its time.  Rows are loaded unconditionally (they exist even when masked) but
    // only accumulated when the mask is non-zero: a masked row may hold NaN (no-alloc convention) and NaN * 0 is NaN.
    constexpr int kU = 4;
    for (int s0 = warp; s0 < S; s0 += kPoolWarps * kU) {
        float m[kU];
        float4 t[kU][NCH];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int sr = s0 + kPoolWarps * u;
            m[u] = sr < S ? (mb ? mb[sr] : 1.0f) : 0.0f;
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const int c = (lane + 32 * i) * 4;
                t[u][i] = (sr < S && c < H) ? load4f(xb + static_cast<size_t>(sr) * H + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            cnt += m[u];
            if (m[u] != 0.0f) {
#pragma unroll
                for (int i = 0; i < NCH; ++i) {
                    acc[i].x += t[u][i].x * m[u]; acc[i].y += t[u][i].y * m[u]; acc[i].z += t[u][i].z * m[u]; acc[i].w += t[u][i].w * m[u];
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = (lane + 32 * i) * 4;
        if (c < H) *reinterpret_cast<float4*>(s_part + warp * H + c) = acc[i];
    }
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    float count = 0.f;
#pragma unroll
    for (int w = 0; w < kPoolWarps; ++w) count += s_cnt[w];
    // thread t owns columns 4t .. 4t+3
    const int c = threadIdx.x * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    float sq = 0.f;
    if (c < H) {
        if (count == 0.0f) {
            a = load4f(xb + c);  // no valid token: the token-0 row (pooling/mod.rs:24-31)
        } else {
#pragma unroll
            for (int w = 0; w < kPoolWarps; ++w) {
                const float4 t = *reinterpret_cast<const float4*>(s_part + w * H + c);
                a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
            }
            a.x /= count; a.y /= count; a.z /= count; a.w /= count;
        }
        sq = (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
    }
    float scale = 1.0f;
    if (normalize) {
        sq = warp_sum(sq);
        if (lane == 0) red[warp] = sq;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < kPoolWarps; ++i) tot += red[i];
        const float nrm = sqrtf(tot);
        if (nrm > 0.0f) scale = 1.0f / nrm;
    }
    if (c < H) {
        if (scale != 1.0f) { a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale; }
        *reinterpret_cast<float4*>(out + static_cast<size_t>(b) * H + c) = a;
    }
}

// Classification head on the CLS row (reference: cpu/encoder/classifier.rs:210-258):
//   z = x[b,0,:]; if pre: z = act(W_pre z + b_pre) (tanh | relu); logits = W_cls z + b_cls.
// fp32 weights and math so the argmax stage matches the fp32 oracle bit-for-bit up to summation order.
// Two kernels so that every SM takes part whatever the batch is (round 1 ran ONE CTA per 8 sequences with a scalar, latency-
// bound weight loop: 112 us for the 1000 pairs of a rerank call, 12-14 % of the C2 / C3 steps):
//   head_dense_kernel  grid (ceil(B / 8), H / 64): 64 rows of W_pre x 8 sequences per CTA, a warp's W_pre row is fetched with up to
//                      eight independent 16-byte loads per lane before the FMAs start; z1 [B,H] fp32 to global memory
//   head_cls_kernel    one warp per (sequence, label) dot product
enum HeadAct : int { HEAD_NONE = 0, HEAD_TANH = 1, HEAD_RELU = 2 };
constexpr int kHeadSeqs = 8;
constexpr int kHeadRows = 64;   // W_pre rows per CTA (8 per warp)
template <typename TIn>
struct HeadParams {
    const TIn* x;  // [B, S, H] last hidden state (bf16 from the encoder, fp32 from the debug hook)
    const float* w_pre; const float* b_pre;  // [H,H], [H] or nullptr
    const float* w_cls; const float* b_cls;  // [C,H], [C]
    float* z1;       // [B, H] fp32 scratch (pre-classifier output); unused when w_pre == nullptr
    float* logits;   // [B, C]
    int B, S, H, C, act;
};
template <typename TIn>
__global__ void __launch_bounds__(256) head_dense_kernel(HeadParams<TIn> p) {
    extern __shared__ __align__(16) float hs[];  // z0[kHeadSeqs][H]
    float* z0 = hs;
    const int b0 = blockIdx.x * kHeadSeqs;
    const int nb = min(kHeadSeqs, p.B - b0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.H, H4 = H >> 2;
    for (int i = tid; i < kHeadSeqs * H; i += 256) {
        const int s = i / H, c = i % H;
        z0[i] = s < nb ? static_cast<float>(p.x[static_cast<size_t>(b0 + s) * p.S * H + c]) : 0.0f;
    }
    __syncthreads();
    constexpr int kMaxV = 8;  // H <= 1024: at most 8 float4 per lane
    for (int rr = 0; rr < kHeadRows / 8; ++rr) {
        const int r = blockIdx.y * kHeadRows + rr * 8 + warp;
        if (r >= H) break;  // warp-uniform
        const float4* w4 = reinterpret_cast<const float4*>(p.w_pre + static_cast<size_t>(r) * H);
        float4 wv[kMaxV];
#pragma unroll
        for (int i = 0; i < kMaxV; ++i) {
            const int c4 = lane + 32 * i;
            wv[i] = c4 < H4 ? __ldg(w4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float acc[kHeadSeqs];
#pragma unroll
        for (int s = 0; s < kHeadSeqs; ++s) acc[s] = 0.0f;
#pragma unroll
        for (int i = 0; i < kMaxV; ++i) {
            const int c4 = lane + 32 * i;
            if (c4 < H4) {
#pragma unroll
                for (int s = 0; s < kHeadSeqs; ++s) {
                    const float4 z = *reinterpret_cast<const float4*>(z0 + s * H + 4 * c4);
                    acc[s] = fmaf(wv[i].x, z.x, acc[s]);
                    acc[s] = fmaf(wv[i].y, z.y, acc[s]);
                    acc[s] = fmaf(wv[i].z, z.z, acc[s]);
                    acc[s] = fmaf(wv[i].w, z.w, acc[s]);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < kHeadSeqs; ++s) acc[s] = warp_sum(acc[s]);
        if (lane == 0) {
            const float bb = p.b_pre ? p.b_pre[r] : 0.0f;
#pragma unroll
            for (int s = 0; s < kHeadSeqs; ++s) {
                if (s < nb) {
                    float v = acc[s] + bb;
                    if (p.act == HEAD_TANH) v = tanhf(v);
                    else if (p.act == HEAD_RELU) v = fmaxf(v, 0.0f);
                    p.z1[static_cast<size_t>(b0 + s) * H + r] = v;
                }
            }
        }
    }
}
// logits[b, cls] = W_cls[cls] . z[b] + b_cls[cls]; z = z1[b] (pre-classifier ran) or the CLS row of x.  One warp per output.
template <typename TIn>
__global__ void __launch_bounds__(256) head_cls_kernel(HeadParams<TIn> p) {
    const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (o >= p.B * p.C) return;
    const int b = o / p.C, cls = o % p.C;
    const float* w = p.w_cls + static_cast<size_t>(cls) * p.H;
    float acc = 0.0f;
    if (p.w_pre != nullptr) {
        const float* z = p.z1 + static_cast<size_t>(b) * p.H;
        for (int c = lane; c < p.H; c += 32) acc = fmaf(__ldg(w + c), z[c], acc);
    } else {
        const TIn* z = p.x + static_cast<size_t>(b) * p.S * p.H;
        for (int c = lane; c < p.H; c += 32) acc = fmaf(__ldg(w + c), static_cast<float>(z[c]), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) p.logits[o] = acc + (p.b_cls ? p.b_cls[cls] : 0.0f);
}
template <typename TIn>
inline void launch_cls_head(const HeadParams<TIn>& p, cudaStream_t st) {
    if (p.w_pre != nullptr) {
        const dim3 grid((p.B + kHeadSeqs - 1) / kHeadSeqs, (p.H + kHeadRows - 1) / kHeadRows);
        head_dense_kernel<TIn><<<grid, 256, static_cast<size_t>(kHeadSeqs) * p.H * sizeof(float), st>>>(p);
    }
    head_cls_kernel<TIn><<<(p.B * p.C + 7) / 8, 256, 0, st>>>(p);
}

// bf16 -> fp32 (the KJC_OUT_HIDDEN output).
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t n4) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n4) *reinterpret_cast<float4*>(out + 4 * i) = load4f(in + 4 * i);
}

// fp32 -> bf16 weight pre-pack (done once at load).
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}

}  // namespace kj
