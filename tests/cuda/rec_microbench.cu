// Microbenchmark of the lane = channel recurrence loop (fused_chain.cu recurrence_row / DF1Core):
// cycles per sample step for one warp alone on an SM, and next to warps that hammer shared memory / the FP pipe.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ int swz(int m) { return (m & ~3) | ((m & 3) ^ ((m >> 3) & 3)); }
struct DF1Core {
    float a1, a2, y1, y2;
    __device__ __forceinline__ float step(float p) { float out = sub(sub(p, mul(a1, y1)), mul(a2, y2)); y2 = y1; y1 = out; return out; }
};
__device__ __forceinline__ void recurrence_row(DF1Core& core, float4* r, int valid_f4) {
    const int last = valid_f4 - 1;
    float4 a = r[swz(0)];
    float4 b = r[swz(min(1, last))];
#pragma unroll 4
    for (int m = 0; m < valid_f4; m++) {
        const float4 nxt = r[swz(min(m + 2, last))];
        float4 x = a;
        x.x = core.step(x.x); x.y = core.step(x.y); x.z = core.step(x.z); x.w = core.step(x.w);
        r[swz(m)] = x;
        a = b; b = nxt;
    }
}
// registers-only variant: no shared memory in the loop at all
__device__ __forceinline__ float chain_only(DF1Core& core, float p, int n) {
    float acc = 0.f;
    for (int i = 0; i < n; i++) acc += core.step(p);
    return acc;
}
constexpr int S = 256, ROW = S + 4, G = 16;
__global__ void bench(int mode, int reps, long long* out, float* sink) {
    extern __shared__ float4 sm4[];
    float* tile = reinterpret_cast<float*>(sm4);
    float4* junk = sm4 + (G * ROW) / 4 + 64;
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    for (int i = t; i < G * ROW; i += blockDim.x) tile[i] = 1e-3f * (float)(i % 97);
    __syncthreads();
    if (w == 0) {
        DF1Core core{-1.8f, 0.83f, 0.f, 0.f};
        long long t0 = clock64();
        float acc = 0.f;
        for (int r = 0; r < reps; r++) {
            if (mode == 3) acc += chain_only(core, 0.001f * r, S);
            else if (lane < G) recurrence_row(core, reinterpret_cast<float4*>(tile + lane * ROW), S / 4);
        }
        long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
        sink[blockIdx.x * 32 + lane] = core.y1 + acc;
    } else if (mode == 1) {  // other warps: heavy shared-memory traffic (like staging + tile transposes)
        float4 v = make_float4(t, 1, 2, 3);
        for (int r = 0; r < reps * 40; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) junk[(k * 256 + t) & 2047] = v;
#pragma unroll
            for (int k = 0; k < 8; k++) { float4 q = junk[(k * 256 + t + 32) & 2047]; v.x += q.x; v.y += q.y; }
        }
        sink[4096 + blockIdx.x * 512 + t] = v.x + v.y;
    } else if (mode == 2 || (mode == 4 && (w & 3) != 0) || (mode == 5 && (w & 3) == 0)) {  // other warps: FP work only, lots of ILP
        float a[16];
        for (int k = 0; k < 16; k++) a[k] = t + k;
        for (int r = 0; r < reps * 200; r++) {
#pragma unroll
            for (int k = 0; k < 16; k++) a[k] = __fmaf_rn(a[k], 1.0001f, 0.5f);
        }
        float s = 0;
        for (int k = 0; k < 16; k++) s += a[k];
        sink[4096 + blockIdx.x * 512 + t] = s;
    }
}
int main() {
    long long* d; float* sink;
    cudaMalloc(&d, 8); cudaMalloc(&sink, (4096 + 148 * 512) * 4);
    const int reps = 200;
    const int smem = (G * ROW + 256) * 4 + 2048 * 16;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[] = {"R warp alone", "R + 8 warps of smem traffic", "R + 8 warps of FP work", "register-only chain (no smem), alone",
                           "R (warp 0) + FP work on warps with wid%4 != 0 (9 of 12)", "R (warp 0) + FP work on warps with wid%4 == 0 (2 of 12)"};
    for (int mode = 0; mode < 6; mode++) {
        int threads = (mode == 1 || mode == 2) ? 288 : (mode >= 4 ? 384 : 32);
        bench<<<148, threads, smem>>>(mode, reps, d, sink);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("%-40s %.1f cycles per sample step\n", names[mode], (double)h / (reps * S));
    }
    return 0;
}
