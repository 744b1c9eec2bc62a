// What does a REALISTIC tcgen05.mma stream cost per instruction (sm_100a)?  mma_issue.cu measured N/2 clk per M = 128, K = 16 MMA
// with one zero-filled operand pair re-read by every instruction; the GEMM kernels measure 130-200 clk.  This probe walks the
// variables in between, one at a time, on the phase-2 shape of the chained kernels (A = a resident 128 x 384 tile of six
// 16 KB k-blocks, B = a ring of N x 64 stages, 24 MMAs per output tile, two TMEM accumulators):
//   VARY  0: every MMA reads the same A / B k-block         1: A walks the six k-blocks, B walks the ring
//   DATA  0: zero-filled operands                           1: random bf16 bit patterns (operand toggling / power)
//   CMT   0: one commit at the end                          1: a tcgen05.commit per k-block (+ one per tile), as the kernels issue
//   HS    0: no waits                                       1: an (already complete) mbarrier try_wait + tcgen05.fence per k-block
//   TS    0: A from shared memory                           1: A from tensor memory
// Prints clk per MMA (issue only / issue + execute), 148 CTAs, one issuing warp each, warp-uniform issue with one elect per k-block.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_of(uint32_t M, uint32_t N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24); }

constexpr int kTiles = 48;  // output tiles per CTA, 24 MMAs each
constexpr int kABytes = 6 * 16384;

template <int N, int VARY, int DATA, int CMT, int HS, int TS>
__global__ void __launch_bounds__(128, 1) probe(unsigned long long* cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tbase_s;
    __shared__ uint64_t bar_done, bar_sink, bar_ready;
    constexpr int kBStage = N * 128;
    constexpr int kStages = (200 * 1024 - kABytes) / kBStage;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (kABytes + kStages * kBStage) / 4; i += blockDim.x) {
        uint32_t v = 0;
        if (DATA) {  // two bf16 of magnitude ~1 with pseudo-random mantissas and signs
            uint32_t h = (i + blockIdx.x * 7919u) * 2654435761u;
            v = (0x3f80u | ((h >> 3) & 0x7fu) | ((h >> 11) & 0x8000u)) | ((0x3f80u | ((h >> 17) & 0x7fu) | ((h >> 9) & 0x8000u)) << 16);
        }
        reinterpret_cast<uint32_t*>(smem)[i] = v;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(&bar_sink)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_ready)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase_s, 0);
    constexpr uint32_t idesc = idesc_of(128, N);
    if (warp == 1) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + kABytes);
        const unsigned long long t0 = clock64();
        int it = 0;
#pragma unroll 1
        for (int t = 0; t < kTiles; ++t) {
            const uint32_t d = TS ? tb + 192 + (t & 1) * 0 : tb + (t & 1) * 256;  // TS: x' in [0,192), one accumulator region behind it (N <= 256 only with one buffer)
#pragma unroll 1
            for (int kb = 0; kb < 6; ++kb, ++it) {
                if (HS == 1) {
                    uint32_t ok = 0;
                    while (!ok) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar_ready)), "r"(1) : "memory");
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint64_t da = desc_sw128(a0 + (VARY ? kb * 16384 : 0));
                const uint64_t db = desc_sw128(b0 + (VARY ? (it % kStages) * kBStage : 0));
                const uint32_t ta = tb + (VARY ? kb * 32 : 0);
                constexpr int kFirst = HS >= 2 ? HS : 4;  // HS = 2 / 3: the NEXT k-block's wait sits behind the first 2 / 3 MMAs of this one
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < kFirst; ++k) {
                        if (TS) umma_ts(d, ta + 8 * k, db + 2 * k, idesc, (kb | k) != 0);
                        else umma(d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    }
                    if (HS < 2 && CMT) {
                        commit(&bar_sink);
                        if (kb == 5) commit(&bar_sink);
                    }
                }
                __syncwarp();
                if (HS >= 2) {
                    uint32_t ok = 0;
                    while (!ok) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar_ready)), "r"(1) : "memory");
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
#pragma unroll
                        for (int k = kFirst; k < 4; ++k) {
                            if (TS) umma_ts(d, ta + 8 * k, db + 2 * k, idesc, 1);
                            else umma(d, da + 2 * k, db + 2 * k, idesc, 1);
                        }
                        if (CMT) {
                            commit(&bar_sink);
                            if (kb == 5) commit(&bar_sink);
                        }
                    }
                    __syncwarp();
                }
            }
        }
        if (elect_one()) commit(&bar_done);
        __syncwarp();
        const unsigned long long t1 = clock64();
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar_done)), "r"(0) : "memory");
        const unsigned long long t2 = clock64();
        if (lane == 0) { cyc[blockIdx.x * 2] = t1 - t0; cyc[blockIdx.x * 2 + 1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

template <int N, int VARY, int DATA, int CMT, int HS, int TS>
void run() {
    unsigned long long* cyc;
    cudaMalloc(&cyc, 148 * 16);
    cudaFuncSetAttribute(probe<N, VARY, DATA, CMT, HS, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int r = 0; r < 3; ++r) probe<N, VARY, DATA, CMT, HS, TS><<<148, 128, 200 * 1024>>>(cyc);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[296];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double a = 0, b = 0;
    for (int i = 0; i < 148; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
    const double n = 148.0 * kTiles * 24;
    printf("N=%3d A=%s vary=%d data=%s commit/kb=%d waits=%d: issue %.1f clk/MMA, issue+execute %.1f clk/MMA  (%s)\n", N, TS ? "tmem" : "smem", VARY, DATA ? "random" : "zero  ", CMT, HS,
           a / n, b / n, cudaGetErrorString(e));
    cudaFree(cyc);
}

template <int N>
void sweep() {
    run<N, 0, 0, 0, 0, 0>();
    run<N, 1, 0, 0, 0, 0>();
    run<N, 1, 1, 0, 0, 0>();
    run<N, 1, 1, 1, 0, 0>();
    run<N, 1, 1, 1, 1, 0>();
    run<N, 1, 1, 1, 2, 0>();
    run<N, 1, 1, 1, 3, 0>();
    run<N, 0, 1, 0, 0, 0>();
}

int main() {
    sweep<128>();
    sweep<192>();
    sweep<256>();
    run<128, 1, 1, 1, 1, 1>();
    run<192, 1, 1, 1, 1, 1>();
    run<256, 1, 1, 1, 1, 1>();
    run<128, 1, 0, 0, 0, 1>();
    run<256, 1, 0, 0, 0, 1>();
    return 0;
}
