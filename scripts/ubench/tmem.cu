// tcgen05.ld throughput per SM sub-partition and its overlap with MUFU work (sm_100a).
// MODE 0: ld.32x32b.x32 + wait::ld only; 1: 32 MUFU.EX2 only; 2: ld + wait + 32 MUFU on the loaded values (the softmax pass-2 pattern);
// 3: ld (no wait until the end of 4 loads) ; 4: ld x32 + 32 FMNMX (pass-1 pattern)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
#define ITER 128
template <int MODE>
__global__ void probe(float* out, unsigned long long* cyc) {
    __shared__ uint32_t tbase_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tbase_s + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(0.5f + i * 0.01f);
    __syncthreads();
    const unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        const uint32_t col = ((it + (warp >> 2)) * 32) & 511;
        if (MODE == 0 || MODE == 2 || MODE == 4) {
            tmem_ld_32x32(tb + col, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (MODE == 3) {
            tmem_ld_32x32(tb + col, v);
            if ((it & 3) == 3) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float f = __uint_as_float(v[i]) * 1e-30f;
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f));
                acc += f;
            }
        }
        if (MODE == 4) {
            float m = __uint_as_float(v[0]);
#pragma unroll
            for (int i = 1; i < 32; ++i) m = fmaxf(m, __uint_as_float(v[i]));
            acc = fmaxf(acc, m);
        }
        if (MODE == 0 || MODE == 3) acc += __uint_as_float(v[it & 31]);
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const unsigned long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase_s) : "memory");
}

template <int MODE>
void run(const char* name) {
    float* out;
    unsigned long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    for (int wps : {1, 2, 4}) {
        const int threads = wps * 128;
        probe<MODE><<<148, threads>>>(out, cyc);
        probe<MODE><<<148, threads>>>(out, cyc);
        cudaDeviceSynchronize();
        unsigned long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += h[i];
        avg /= 148;
        printf("%-34s warps/sched=%d: %8.0f clk -> %.1f clk per (32x32 chunk) per scheduler\n", name, wps, avg, avg / (double(wps) * ITER));
    }
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("tcgen05.ld x32 + wait");
    run<3>("tcgen05.ld x32, wait every 4");
    run<1>("32 MUFU.EX2 + 32 FADD");
    run<2>("ld + wait + 32 MUFU + 32 FADD");
    run<4>("ld + wait + 32 FMNMX");
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
