// Instruction-throughput probes for the softmax / epilogue inner loops (sm_100a): cycles per warp-instruction per SM sub-partition
// for MUFU.EX2, the f32x2 -> bf16x2 pack (F2FP), FFMA, FADD, FMNMX and the integer-rounding alternative to F2FP, at 1 / 2 / 4 warps
// per scheduler, plus tcgen05.ld throughput.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define ITER 256
template <int OP>
__global__ void probe(float* out, unsigned long long* cyc, float seed) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
    uint32_t acc = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 1) a[i] = fmaf(a[i], 0.999f, 0.001f);
            if (OP == 2) a[i] = a[i] + 0.001f;
            if (OP == 3) a[i] = fmaxf(a[i], a[(i + 1) & 15]);
            if (OP == 4) {  // pack pair (i, i^1) -> bf16x2, fed back so it is not hoisted
                uint32_t r;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[i ^ 1]));
                acc ^= r;
            }
            if (OP == 5) {  // integer rounding pack: (bits + 0x8000) >> 16 pairs via PRMT
                uint32_t x = __float_as_uint(a[i]) + 0x8000u, y = __float_as_uint(a[i ^ 1]) + 0x8000u;
                acc ^= __byte_perm(x, y, 0x7632);
            }
            if (OP == 6) {  // the pass-2 mix per element pair: 2 FFMA + 2 EX2 + 2 FADD + 1 pack
                float f0 = fmaf(a[i], 0.5f, -1.0f), f1 = fmaf(a[i ^ 1], 0.5f, -1.0f);
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f0));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f1));
                a[(i + 2) & 15] += f0 + f1;
                uint32_t r;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(f0), "f"(f1));
                acc ^= r;
            }
        }
    }
    const unsigned long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter_instr) {
    float* out;
    unsigned long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    for (int warps_per_sched : {1, 2, 4}) {
        const int threads = warps_per_sched * 4 * 32;
        probe<OP><<<148, threads>>>(out, cyc, 0.5f);
        probe<OP><<<148, threads>>>(out, cyc, 0.5f);
        cudaDeviceSynchronize();
        unsigned long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += h[i];
        avg /= 148;
        // warp-instructions issued per scheduler = warps_per_sched * ITER * 16 * per_iter_instr
        printf("%-28s warps/sched=%d: %8.0f clk -> %.2f clk per warp-instr per scheduler (%.2f per element-op)\n", name, warps_per_sched, avg,
               avg / (double(warps_per_sched) * ITER * 16 * per_iter_instr), avg / (double(warps_per_sched) * ITER * 16));
    }
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("MUFU.EX2", 1);
    run<1>("FFMA", 1);
    run<2>("FADD", 1);
    run<3>("FMNMX", 1);
    run<4>("F2FP pack bf16x2", 1);
    run<5>("IADD+IADD+PRMT pack", 3);
    run<6>("softmax mix (7 instr / pair)", 7);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
