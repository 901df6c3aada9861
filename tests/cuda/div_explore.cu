// Exploration tool: mismatch counts of candidate constant-division sequences vs IEEE, per exponent bucket.
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ float varA(float a, float b, float rh, float rl) {
    float q0 = __fmul_rn(a, rh); float rem = __fmaf_rn(-q0, b, a); return __fmaf_rn(rem, rh, q0);
}
__device__ __forceinline__ float varB(float a, float b, float rh, float rl) {
    float q0 = __fmaf_rn(a, rh, __fmul_rn(a, rl)); float rem = __fmaf_rn(-q0, b, a); return __fmaf_rn(rem, rh, q0);
}
__device__ __forceinline__ float varC(float a, float b, float rh, float rl) {  // two corrections
    float q0 = __fmul_rn(a, rh); float rem = __fmaf_rn(-q0, b, a); float q1 = __fmaf_rn(rem, rh, q0);
    float rem1 = __fmaf_rn(-q1, b, a); return __fmaf_rn(rem1, rh, q1);
}
template <int V>
__global__ void check(float b, float rh, float rl, unsigned long long* hist) {
    unsigned long long base = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 256ull;
    for (int k = 0; k < 256; k++) {
        const unsigned bits = (unsigned)(base + k);
        const float a = __uint_as_float(bits);
        float q = V == 0 ? varA(a, b, rh, rl) : V == 1 ? varB(a, b, rh, rl) : varC(a, b, rh, rl);
        const float ref = __fdiv_rn(a, b);
        if (__float_as_uint(q) != __float_as_uint(ref) && !(q != q && ref != ref)) atomicAdd(&hist[(bits >> 23) & 0xFF], 1ull);
    }
}
int main(int argc, char** argv) {
    unsigned long long *d, h[256];
    cudaMalloc(&d, 256 * 8);
    for (int i = 1; i < argc; i++) {
        const float b = strtof(argv[i], nullptr);
        const double rd = 1.0 / (double)b;
        const float rh = (float)rd, rl = (float)(rd - (double)rh);
        for (int v = 0; v < 3; v++) {
            cudaMemset(d, 0, 256 * 8);
            if (v == 0) check<0><<<65536, 256>>>(b, rh, rl, d);
            if (v == 1) check<1><<<65536, 256>>>(b, rh, rl, d);
            if (v == 2) check<2><<<65536, 256>>>(b, rh, rl, d);
            cudaDeviceSynchronize();
            cudaMemcpy(h, d, 256 * 8, cudaMemcpyDeviceToHost);
            unsigned long long tot = 0, mid = 0; int lo = 999, hi = -1;
            for (int e = 0; e < 256; e++) { tot += h[e]; if (e >= 30 && e <= 220) mid += h[e]; if (h[e]) { if (e < lo) lo = e; if (e > hi) hi = e; } }
            printf("b=%.9g var%c total=%llu  exp[30..220]=%llu  exp-range-with-mismatch=[%d,%d]  e0=%llu e1=%llu e2=%llu e254=%llu e255=%llu\n", b, 'A' + v, tot, mid, lo, hi, h[0], h[1], h[2], h[254], h[255]);
        }
    }
    return 0;
}
