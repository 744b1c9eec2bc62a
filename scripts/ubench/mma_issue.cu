// Raw tcgen05.mma issue/execute rate vs N and vs how the issuing code is written (sm_100a).
//   style 0: everything under `if (lane == 0)` (how the round-1 kernels issue)          style 1: warp-uniform code, elect_one only around the MMA
// A, B tiles: zero-filled smem (K-major, 128B swizzle).  Prints clk per MMA (M = 128, K = 16).
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
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint64_t desc_atom(uint32_t addr, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(sbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_of(uint32_t M, uint32_t N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24); }

#define NMMA 2048
template <int N, int STYLE, int TS, int BMN = 0, int CE = 0>
__global__ void __launch_bounds__(128, 1) probe(unsigned long long* cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tbase_s;
    __shared__ uint64_t bar;
    __shared__ uint64_t bar2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 100000;" ::"r"(smem_u32(&bar2)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tbase_s;
    constexpr uint32_t idesc = idesc_of(128, N) | (BMN ? (1u << 16) : 0u);
    if (warp == 1) {
        const unsigned long long t0 = clock64();
        if (STYLE == 0) {
            if (lane == 0) {
                const uint64_t da = desc_sw128(smem_u32(smem)), db = desc_sw128(smem_u32(smem + 16384));
#pragma unroll 1
                for (int i = 0; i < NMMA / 4; ++i) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (TS) umma_ts(tb + 256, tb + 8 * k, db + 2 * k, idesc, 1);
                        else umma(tb + 256, da + 2 * k, db + 2 * k, idesc, 1);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
        } else {
            const uint64_t da = desc_sw128(smem_u32(smem)), db = BMN ? desc_atom(smem_u32(smem + 16384), N == 32 ? 512 : 1024, N == 32 ? 4u : 2u) : desc_sw128(smem_u32(smem + 16384));
            const uint32_t tbu = __shfl_sync(0xffffffffu, tb, 0);
#pragma unroll 1
            for (int i = 0; i < NMMA / 4; ++i) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (elect_one()) {
                        const uint64_t dbk = BMN ? db + ((k * 16 * (N == 32 ? 64 : 128)) >> 4) : db + 2 * k;
                        if (TS) umma_ts(tbu + 256, tbu + 8 * k, dbk, idesc, 1);
                        else umma(tbu + 256, da + 2 * k, dbk, idesc, 1);
                    }
                    if (CE == 1 || (CE == 2 && (k & 1)) || (CE == 4 && k == 3))
                        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
                }
            }
            if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            __syncwarp();
        }
        const unsigned long long t1 = clock64();  // issue done
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        const unsigned long long t2 = clock64();  // execution done
        if (lane == 0) { cyc[blockIdx.x * 2] = t1 - t0; cyc[blockIdx.x * 2 + 1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

template <int N, int STYLE, int TS, int BMN = 0, int CE = 0>
void run() {
    unsigned long long* cyc;
    cudaMalloc(&cyc, 148 * 16);
    cudaFuncSetAttribute(probe<N, STYLE, TS, BMN, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int r = 0; r < 2; ++r) probe<N, STYLE, TS, BMN, CE><<<148, 128, 65536>>>(cyc);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[296];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double a = 0, b = 0;
    for (int i = 0; i < 148; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
    printf("N=%3d %s %s %s commit/%d: issue %.1f clk/MMA, issue+execute %.1f clk/MMA  (%s)\n", N, TS ? "A=tmem" : "A=smem", BMN ? "B=MN-major" : "B=K-major ", STYLE ? "uniform+elect" : "if(lane==0)  ", CE, a / 148 / NMMA,
           b / 148 / NMMA, cudaGetErrorString(e));
    cudaFree(cyc);
}


// MMA issue rate of warp 1 while warps 4..19 run "noise": 0 none, 1 MUFU+FFMA loop, 2 tcgen05.ld loop, 3 FFMA only, 4 tcgen05.ld + MUFU (softmax-like)
template <int N, int NOISE>
__global__ void __launch_bounds__(640, 1) probe_noise(unsigned long long* cyc, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tbase_s;
    __shared__ uint64_t bar;
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        done = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tbase_s;
    constexpr uint32_t idesc = idesc_of(128, N) | (1u << 16);
    if (warp == 1) {
        const uint64_t db = desc_atom(smem_u32(smem + 16384), N == 32 ? 512 : 1024, N == 32 ? 4u : 2u);
        const uint32_t tbu = __shfl_sync(0xffffffffu, tb, 0);
        const unsigned long long t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < NMMA / 8; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (elect_one()) umma_ts(tbu + 256, tbu + 8 * k, db + ((k * 16 * (N == 32 ? 64 : 128)) >> 4), idesc, 1);
        }
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        const unsigned long long t1 = clock64();
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        const unsigned long long t2 = clock64();
        if (lane == 0) { cyc[blockIdx.x * 2] = t1 - t0; cyc[blockIdx.x * 2 + 1] = t2 - t0; done = 1; }
    } else if (warp >= 4 && NOISE != 0) {
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = 0.5f + i * 0.01f + lane * 1e-5f;
        const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16);
        int it = 0;
        while (!done) {
            ++it;
            if (NOISE == 1 || NOISE == 3) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        a[i] = fmaf(a[i], 0.999f, 0.001f);
                        if (NOISE == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                    }
            }
            if (NOISE == 5 || NOISE == 6 || NOISE == 7) {
                uint32_t ok;
                if (NOISE == 5) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(1) : "memory");
                if (NOISE == 6) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(1), "r"(10000000) : "memory");
                if (NOISE == 7) asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(1) : "memory");
                a[0] += ok;
            }
            if (NOISE == 2 || NOISE == 4) {
                uint32_t v[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(tl + ((it * 32) & 127))
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float f = __uint_as_float(v[i] & 0x3fffffffu) * 1e-30f + __uint_as_float(v[i + 16] & 0x3fffffffu) * 1e-30f;
                    if (NOISE == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f)); float g = f * 0.5f; asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(g)); f += g; }
                    a[i] += f;
                }
            }
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += a[i];
        sink[blockIdx.x * 640 + threadIdx.x] = s;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
template <int N, int NOISE>
void run_noise(const char* what) {
    unsigned long long* cyc;
    float* sink;
    cudaMalloc(&cyc, 148 * 16);
    cudaMalloc(&sink, 148 * 640 * 4);
    cudaFuncSetAttribute(probe_noise<N, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int r = 0; r < 2; ++r) probe_noise<N, NOISE><<<148, 640, 65536>>>(cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[296];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double a = 0, b = 0;
    for (int i = 0; i < 148; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
    printf("N=%3d TS MN-major, 16 noise warps [%s]: issue %.1f clk/MMA, issue+execute %.1f clk/MMA (%s)\n", N, what, a / 148 / NMMA, b / 148 / NMMA, cudaGetErrorString(e));
    cudaFree(cyc);
    cudaFree(sink);
}

int main() {
    run_noise<32, 0>("none"); run_noise<32, 5>("try_wait spin"); run_noise<32, 6>("try_wait spin, 10 ms suspend hint"); run_noise<32, 7>("test_wait spin"); run_noise<32, 3>("FFMA"); run_noise<32, 1>("FFMA+MUFU"); run_noise<32, 2>("tcgen05.ld"); run_noise<32, 4>("tcgen05.ld+MUFU");

    run<32, 1, 1, 1, 4>(); run<32, 1, 1, 1, 2>(); run<32, 1, 1, 1, 1>(); run<128, 1, 0, 0, 4>(); run<128, 1, 0, 0, 2>(); run<192, 1, 0, 0, 4>(); run<256, 1, 0, 0, 4>();
    run<32, 1, 1, 1>(); run<32, 1, 0, 1>(); run<64, 1, 1, 1>(); run<64, 1, 0, 1>();
    run<32, 0, 0>(); run<32, 1, 0>(); run<32, 0, 1>(); run<32, 1, 1>();
    run<64, 0, 0>(); run<64, 1, 0>();
    run<128, 0, 0>(); run<128, 1, 0>(); run<128, 1, 1>();
    run<192, 0, 0>(); run<192, 1, 0>();
    run<256, 0, 0>(); run<256, 1, 0>(); run<256, 1, 1>();
    return 0;
}
