// Fused short-sequence encoder self-attention (S <= 512, head_dim 32 or 64):
// scores = Q K^T / sqrt(d), key-padding mask, row softmax, P V, heads merged on store.
// Replaces EncoderSelfAttention::forward{,_noalloc} between the QKV projection
// and the output projection (reference: kjarni-transformers/src/cpu/encoder/
// encoder_self_attention.rs:213-298 = split heads, matmul_4d (faer), scale :244-249,
// apply_padding_mask :311-325 / utils/masks.rs:7-36, softmax activations.rs:259-279,
// P.V, permute_merge_heads :384-425).  Nothing [B,heads,S,S]-sized ever touches HBM.
//
// One CTA per (sequence, head); 8 warps x 16 query rows per pass; K, V, Q of the
// head staged once in shared memory (padded rows: conflict-free ldmatrix);
// bf16 mma.sync m16n8k16 with fp32 accumulation, online softmax in fp32 registers.
#pragma once
#include "ptx.cuh"

namespace kj {

struct AttnParams {
    const __nv_bfloat16* qkv;  // [B*S, 3H]  (Q | K | V), head h at columns h*d
    const float* mask;         // [B, S] 1 = token, 0 = padding; nullptr = all ones
    __nv_bfloat16* ctx;        // [B*S, H]
    int B, S, H, heads;
    float scale_log2e;         // (1/sqrt(d)) * log2(e)
    int nan_if_all_masked;     // 1: no-alloc (-inf) convention, a fully padded sequence yields NaN
    int max_ctas = 0;          // persistent tcgen05 kernel: CTAs to launch (0 = one per SM)
    int dbg = 0;               // probes (KJC_ATTN_DBG through kjc_dbg_attention): 1 = TMA loads only (no MMA, no softmax)
    unsigned long long* trace = nullptr;  // optional [gridDim.x][64] %globaltimer stamps (KJC_ATTN_TRACE, dbg_attention)
    int ld_qkv = 0;            // row pitch of qkv in elements (0 = 3H); attention_ts only: the encoder lays qkv rows over the FFN rows (pitch I)
};

constexpr int kAttnThreads = 256;
constexpr int kAttnKeyBlock = 64;

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int D>
__global__ void __launch_bounds__(kAttnThreads) attention_kernel(AttnParams p) {
    constexpr int kPitch = D * 2 + 16;  // bytes per smem row; +16 keeps ldmatrix bank-conflict free
    constexpr int kChunksPerRow = D / 8;
    extern __shared__ __align__(16) uint8_t smem_attn[];
    const int S = p.S;
    const int s_pad = (S + kAttnKeyBlock - 1) / kAttnKeyBlock * kAttnKeyBlock;
    uint8_t* sq = smem_attn;
    uint8_t* sk = sq + static_cast<size_t>(s_pad) * kPitch;
    uint8_t* sv = sk + static_cast<size_t>(s_pad) * kPitch;
    float* smask = reinterpret_cast<float*>(sv + static_cast<size_t>(s_pad) * kPitch);  // [s_pad] additive code
    __shared__ int s_valid;

    constexpr float kMaskedLog2 = -1.0e9f * 1.4426950408889634f;
    const int b = blockIdx.x / p.heads;
    const int h = blockIdx.x % p.heads;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s_valid = 0;
    __syncthreads();

    // ---- stage Q, K, V head slices (16-byte chunks), zero-fill rows >= S
    const size_t ld = static_cast<size_t>(3) * p.H;
    const __nv_bfloat16* base = p.qkv + static_cast<size_t>(b) * S * ld + static_cast<size_t>(h) * D;
    const uint32_t sq_base = smem_u32(sq), sk_base = smem_u32(sk), sv_base = smem_u32(sv);
    for (int i = tid; i < s_pad * kChunksPerRow * 3; i += kAttnThreads) {
        const int which = i / (s_pad * kChunksPerRow);
        const int r = (i / kChunksPerRow) % s_pad;
        const int c = i % kChunksPerRow;
        const bool in = r < S;
        const __nv_bfloat16* src = base + static_cast<size_t>(in ? r : 0) * ld + which * p.H + c * 8;
        const uint32_t dst = (which == 0 ? sq_base : (which == 1 ? sk_base : sv_base)) + r * kPitch + c * 16;
        // async 16-byte copies (no register staging, all in flight at once); rows >= S are zero-filled (src-size 0)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(in ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    int local_valid = 0;
    for (int j = tid; j < s_pad; j += kAttnThreads) {
        float code;  // 0 = keep; otherwise the value the score is replaced by: -1e9*log2e (padding) or -inf (beyond S)
        if (j >= S) code = -INFINITY;
        else {
            const bool keep = (p.mask == nullptr) || (p.mask[static_cast<size_t>(b) * S + j] != 0.0f);
            code = keep ? 0.0f : kMaskedLog2;
            local_valid += keep ? 1 : 0;
        }
        smask[j] = code;
    }
    if (local_valid) atomicAdd(&s_valid, local_valid);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const bool poison = p.nan_if_all_masked && (s_valid == 0);

    const uint32_t sq_u = sq_base, sk_u = sk_base, sv_u = sv_base;

    for (int q0 = warp * 16; q0 < S; q0 += (kAttnThreads / 32) * 16) {
        // Q fragments for this warp's 16 rows
        uint32_t qf[D / 16][4];
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            const uint32_t addr = sq_u + (q0 + (lane & 15)) * kPitch + ks * 32 + (lane >> 4) * 16;
            ldmatrix_x4(addr, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
        }
        float o[D / 8][4];
#pragma unroll
        for (int n = 0; n < D / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.0f;
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;

        for (int kb = 0; kb < s_pad; kb += kAttnKeyBlock) {
            float sc[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.0f;
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks) {
#pragma unroll
                for (int np = 0; np < 4; ++np) {  // pairs of 8-key tiles
                    const int key = kb + np * 16 + ((lane >> 4) << 3) + (lane & 7);
                    const uint32_t addr = sk_u + key * kPitch + ks * 32 + ((lane >> 3) & 1) * 16;
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4(addr, b0, b1, b2, b3);
                    mma_bf16_16816(sc[2 * np], qf[ks], b0, b1);
                    mma_bf16_16816(sc[2 * np + 1], qf[ks], b2, b3);
                }
            }
            // scale + mask, block row max
            float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int j = kb + n * 8 + (lane & 3) * 2;
                const float2 cm = *reinterpret_cast<const float2*>(smask + j);
                const float c0 = cm.x, c1 = cm.y;
                sc[n][0] = c0 == 0.0f ? sc[n][0] * p.scale_log2e : c0;
                sc[n][1] = c1 == 0.0f ? sc[n][1] * p.scale_log2e : c1;
                sc[n][2] = c0 == 0.0f ? sc[n][2] * p.scale_log2e : c0;
                sc[n][3] = c1 == 0.0f ? sc[n][3] * p.scale_log2e : c1;
                bm0 = fmaxf(bm0, fmaxf(sc[n][0], sc[n][1]));
                bm1 = fmaxf(bm1, fmaxf(sc[n][2], sc[n][3]));
            }
            bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
            bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
            bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
            bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
            // key 0 always exists, so the running max is finite from the first block on;
            // a later block made only of non-existent keys keeps the old max.
            const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);
            const float corr0 = ex2_approx(m0 - nm0), corr1 = ex2_approx(m1 - nm1);
            m0 = nm0;
            m1 = nm1;
            float rs0 = 0.0f, rs1 = 0.0f;
            uint32_t pf[4][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float p0 = ex2_approx(sc[n][0] - m0), p1 = ex2_approx(sc[n][1] - m0);
                const float p2 = ex2_approx(sc[n][2] - m1), p3 = ex2_approx(sc[n][3] - m1);
                rs0 += p0 + p1;
                rs1 += p2 + p3;
                pf[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p0, p1);
                pf[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p2, p3);
            }
            l0 = l0 * corr0 + rs0;
            l1 = l1 * corr1 + rs1;
#pragma unroll
            for (int n = 0; n < D / 8; ++n) {
                o[n][0] *= corr0; o[n][1] *= corr0;
                o[n][2] *= corr1; o[n][3] *= corr1;
            }
            // O += P V
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {  // 16 keys per step
#pragma unroll
                for (int np = 0; np < D / 16; ++np) {  // pairs of 8-dim tiles
                    const int key = kb + kk * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
                    const uint32_t addr = sv_u + key * kPitch + np * 32 + (lane >> 4) * 16;
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4_trans(addr, b0, b1, b2, b3);
                    mma_bf16_16816(o[2 * np], pf[kk], b0, b1);
                    mma_bf16_16816(o[2 * np + 1], pf[kk], b2, b3);
                }
            }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
        if (poison) inv0 = inv1 = __int_as_float(0x7fc00000);
        const int r0 = q0 + (lane >> 2), r1 = r0 + 8;
        __nv_bfloat16* out = p.ctx + static_cast<size_t>(b) * S * p.H + static_cast<size_t>(h) * D + (lane & 3) * 2;
#pragma unroll
        for (int n = 0; n < D / 8; ++n) {
            if (r0 < S) *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(r0) * p.H + n * 8) = pack_bf16(o[n][0] * inv0, o[n][1] * inv0);
            if (r1 < S) *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(r1) * p.H + n * 8) = pack_bf16(o[n][2] * inv1, o[n][3] * inv1);
        }
    }
}

inline size_t attention_smem_bytes(int S, int D) {
    const int s_pad = (S + kAttnKeyBlock - 1) / kAttnKeyBlock * kAttnKeyBlock;
    return static_cast<size_t>(3) * s_pad * (D * 2 + 16) + static_cast<size_t>(s_pad) * 4;
}

}  // namespace kj
