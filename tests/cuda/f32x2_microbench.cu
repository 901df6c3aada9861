// f32x2_microbench.cu — issue/pipe throughput of the packed fp32 instructions (add/mul/fma.rn.f32x2) against
// their scalar forms on sm_100a.  Decides whether the FFT FIR kernel should process two independent transforms
// per thread in packed registers.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a f32x2_microbench.cu
#include <cuda_runtime.h>

#include <cstdio>

constexpr int kChains = 8;
constexpr int kIters = 4096;

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <int MODE>  // 0: scalar FFMA, 1: FFMA2, 2: FADD2, 3: mix FFMA2 + scalar FFMA 1:1
__global__ void bench(float* out, float s) {
    unsigned long long p[kChains];
    float q[kChains];
    for (int i = 0; i < kChains; i++) {
        q[i] = s * (threadIdx.x + i);
        p[i] = ((unsigned long long)__float_as_uint(q[i]) << 32) | __float_as_uint(q[i] + 1.f);
    }
    const unsigned long long m = ((unsigned long long)__float_as_uint(0.999f) << 32) | __float_as_uint(1.001f);
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < kChains; i++) {
            if (MODE == 0) q[i] = fma1(q[i], 0.999f, s);
            if (MODE == 1) p[i] = fma2(p[i], m, m);
            if (MODE == 2) p[i] = add2(p[i], m);
            if (MODE == 3) { p[i] = fma2(p[i], m, m); q[i] = fma1(q[i], 0.999f, s); }
        }
    }
    float acc = 0.f;
    for (int i = 0; i < kChains; i++) acc += q[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int per_iter_instr) {
    float* d;
    cudaMalloc(&d, 148 * 4 * 1024 * 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int threads : {128, 256, 512, 1024}) {
        bench<MODE><<<148, threads>>>(d, 1e-3f);
        cudaEventRecord(a);
        bench<MODE><<<148, threads>>>(d, 1e-3f);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        const double winstr = (double)threads / 32 * kIters * kChains * per_iter_instr;  // warp-instructions per SM
        printf("%-28s threads/SM %4d: %.3f ms, %.2f warp-instr/ns/SM (x1.9 GHz: %.2f per clk per SM)\n", name, threads, ms,
               winstr / (ms * 1e6), winstr / (ms * 1e6) / 1.9);
    }
    cudaFree(d);
}

int main() {
    run<0>("scalar FFMA", 1);
    run<1>("FFMA2", 1);
    run<2>("FADD2", 1);
    run<3>("FFMA2 + FFMA interleaved", 2);
    return 0;
}
