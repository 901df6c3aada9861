// Exhaustive check of exact_math.cuh's div_const against IEEE division (__fdiv_rn): for every
// divisor on the command line and ALL 2^32 f32 dividends, whenever the acceptance rule (min|q| >= 2^-90, max|a| <= 2^90) holds
// its result must be bit-identical (NaNs compared as NaNs).  Prints "divisor mismatches flagged".
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../dsp_stuff_b200/csrc/exact_math.cuh"

__global__ void check(float b, float r, unsigned long long* mism, unsigned long long* flagged) {
    unsigned long long base = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 256ull;
    unsigned long long m = 0, f = 0;
    for (int k = 0; k < 256; k++) {
        const unsigned bits = (unsigned)(base + k);
        const float a = __uint_as_float(bits);
        float mn = dspb::kDivHi, mx = 0.0f;
        const float q = dspb::div_const(a, dspb::ConstDiv{b, r}, mn, mx);
        const float ref = __fdiv_rn(a, b);
        const bool bad = !dspb::div_const_accept(mn, mx);
        if (bad) f++;
        else if (__float_as_uint(q) != __float_as_uint(ref) && !(q != q && ref != ref)) m++;
    }
    if (m) atomicAdd(mism, m);
    if (f) atomicAdd(flagged, f);
}

int main(int argc, char** argv) {
    unsigned long long *d, h[2];
    cudaMalloc(&d, 16);
    int rc = 0;
    for (int i = 1; i < argc; i++) {
        const float b = strtof(argv[i], nullptr);
        const float r = 1.0f / b;
        cudaMemset(d, 0, 16);
        check<<<65536, 256>>>(b, r, d, d + 1);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error\n"); return 2; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%.9g %llu %llu\n", b, h[0], h[1]);
        if (h[0]) rc = 1;
    }
    return rc;
}
