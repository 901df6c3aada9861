// fir_fft.cu — overlap-save FFT convolution for the Fir node (nodes/fir.rs:179-225), sm_100a.
//
// One CTA convolves one segment of TWO channels at once: z = xA + i*xB, Z = FFT(z), Y = Z .* H,
// y = IFFT(Y); because h is real, Re(y) = xA * h and Im(y) = xB * h.  F = 8192 complex points live in
// (padded, bank-conflict-free) shared memory; the transform is decimation-in-frequency with radices
// 8, 8, 8, 16 forward and the exact mirror (decimation-in-time, conjugate twiddles) backward, so the
// forward output order is irrelevant: H is produced ONCE per tap set by running the very same forward
// passes (in f64) on the zero-padded impulse response and is stored in that same order, pre-scaled
// by 1/F.  The last forward pass, the spectrum product and the first inverse pass happen in
// registers.  The first pass reads the input window straight from global memory and the last pass
// writes the F-N+1 valid outputs straight back (coalesced 4-byte accesses: the window start is not
// 16-byte aligned because N-1 is odd).
//
// Accuracy: f32 butterflies, twiddles from an f64-computed table: ~1e-7 of the signal rms, inside
// the 1e-5 / -100 dBFS parity bar against the reference's f64 accumulation (tests/test_gpu_fir.py).
#include <cuda_runtime.h>

#include <cmath>
#include <type_traits>
#include <utility>
#include <vector>

#include "plan.h"

namespace dspb {
namespace {

constexpr int kF = 8192;       // complex FFT size
constexpr int kLog2F = 13;
constexpr int kNT = 256;       // threads
constexpr int kPadded = kF + kF / 16;

template <typename T>
struct C2 {
    T x, y;
};
template <typename T> __device__ __forceinline__ C2<T> operator+(C2<T> a, C2<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> __device__ __forceinline__ C2<T> operator-(C2<T> a, C2<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> __device__ __forceinline__ C2<T> cmul(C2<T> a, C2<T> w) { return {a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x}; }
template <typename T> __device__ __forceinline__ C2<T> cmulc(C2<T> a, C2<T> w) { return {a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y}; }  // a * conj(w)

template <int N, class Fn, int... I>
__device__ __forceinline__ void static_for_impl(Fn&& f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class Fn>
__device__ __forceinline__ void static_for(Fn&& f) {
    static_for_impl<N>(static_cast<Fn&&>(f), std::make_integer_sequence<int, N>{});
}

// a * w_SZ^I  (forward: w = exp(-2 pi i / SZ); INV: conjugate), I < SZ/2, SZ in {2,4,8,16}
template <int SZ, int I, bool INV, typename T>
__device__ __forceinline__ C2<T> tw(C2<T> a) {
    if constexpr (I == 0) {
        return a;
    } else if constexpr (4 * I == SZ) {
        if constexpr (INV) return {-a.y, a.x};
        else return {a.y, -a.x};
    } else if constexpr (8 * I == SZ) {
        const T h = T(0.70710678118654752440);
        if constexpr (INV) return {(a.x - a.y) * h, (a.x + a.y) * h};
        else return {(a.x + a.y) * h, (a.y - a.x) * h};
    } else if constexpr (8 * I == 3 * SZ) {
        const T h = T(0.70710678118654752440);
        if constexpr (INV) return {-(a.x + a.y) * h, (a.x - a.y) * h};
        else return {(a.y - a.x) * h, -(a.x + a.y) * h};
    } else {
        static_assert(SZ == 16, "generic twiddles are only tabulated for SZ = 16");
        constexpr double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
        constexpr double cs[4][2] = {{c1, s1}, {s1, c1}, {-s1, c1}, {-c1, s1}};  // I = 1, 3, 5, 7
        const T c = T(cs[(I - 1) / 2][0]), s = T(cs[(I - 1) / 2][1]);
        if constexpr (INV) return {a.x * c - a.y * s, a.y * c + a.x * s};
        else return {a.x * c + a.y * s, a.y * c - a.x * s};
    }
}

// in-register radix-2 decimation-in-frequency DFT of R points: natural in, bit-reversed out
template <int R, typename T>
__device__ __forceinline__ void fft_dif(C2<T> (&v)[R]) {
    static_for<4>([&](auto st) {
        constexpr int sz = R >> decltype(st)::value;
        if constexpr (sz >= 2) {
            constexpr int half = sz / 2;
            static_for<R / sz>([&](auto bk) {
                static_for<half>([&](auto ii) {
                    constexpr int i = decltype(ii)::value, o = decltype(bk)::value * sz;
                    const C2<T> a = v[o + i], b = v[o + i + half];
                    v[o + i] = a + b;
                    v[o + i + half] = tw<sz, i, false, T>(a - b);
                });
            });
        }
    });
}
// exact mirror: bit-reversed in, natural out, conjugate twiddles, unnormalised (gain R)
template <int R, typename T>
__device__ __forceinline__ void ifft_dit(C2<T> (&v)[R]) {
    static_for<4>([&](auto st) {
        constexpr int sz = 2 << decltype(st)::value;
        if constexpr (sz <= R) {
            constexpr int half = sz / 2;
            static_for<R / sz>([&](auto bk) {
                static_for<half>([&](auto ii) {
                    constexpr int i = decltype(ii)::value, o = decltype(bk)::value * sz;
                    const C2<T> a = v[o + i], b = tw<sz, i, true, T>(v[o + i + half]);
                    v[o + i] = a + b;
                    v[o + i + half] = a - b;
                });
            });
        }
    });
}

__host__ __device__ constexpr int bitrev3(int s) { return ((s & 1) << 2) | (s & 2) | ((s >> 2) & 1); }
__device__ __forceinline__ int pad(int p) { return p + (p >> 4); }

// twiddles w^q, q = 1..7, from three table look-ups (w, w^2, w^4) and four products
template <typename T>
__device__ __forceinline__ void twiddles8(const C2<T>* __restrict__ W, int k1, C2<T> (&w)[8]) {
    w[1] = W[k1 & (kF - 1)];
    w[2] = W[(2 * k1) & (kF - 1)];
    w[4] = W[(4 * k1) & (kF - 1)];
    w[3] = cmul(w[1], w[2]);
    w[5] = cmul(w[1], w[4]);
    w[6] = cmul(w[2], w[4]);
    w[7] = cmul(w[3], w[4]);
}

// forward radix-8 pass on the shared array: sub-transform size M, L = M/8
template <int M, typename T>
__device__ __forceinline__ void fwd_pass8(C2<T>* a, const C2<T>* __restrict__ W, int t) {
    constexpr int L = M / 8;
#pragma unroll 1
    for (int k = 0; k < kF / 8 / kNT; k++) {
        const int u = t + kNT * k, b = u / L, j = u % L, base = b * M + j;
        C2<T> v[8], w[8];
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = a[pad(base + r * L)];
        twiddles8(W, j * (kF / M), w);
        fft_dif<8>(v);
        a[pad(base)] = v[0];
#pragma unroll
        for (int s = 1; s < 8; s++) a[pad(base + bitrev3(s) * L)] = cmul(v[s], w[bitrev3(s)]);
    }
}
template <int M, typename T>
__device__ __forceinline__ void inv_pass8(C2<T>* a, const C2<T>* __restrict__ W, int t) {
    constexpr int L = M / 8;
#pragma unroll 1
    for (int k = 0; k < kF / 8 / kNT; k++) {
        const int u = t + kNT * k, b = u / L, j = u % L, base = b * M + j;
        C2<T> v[8], w[8];
        twiddles8(W, j * (kF / M), w);
        v[0] = a[pad(base)];
#pragma unroll
        for (int s = 1; s < 8; s++) v[s] = cmulc(a[pad(base + bitrev3(s) * L)], w[bitrev3(s)]);
        ifft_dit<8>(v);
#pragma unroll
        for (int r = 0; r < 8; r++) a[pad(base + r * L)] = v[r];
    }
}

// ---- the convolution kernel --------------------------------------------------------------------------
__global__ void __launch_bounds__(kNT, 3)
fir_fft_kernel(const float* __restrict__ U, long long u_stride, int hist_pad, float* __restrict__ Y, long long y_stride,
               const float2* __restrict__ Hg, const float2* __restrict__ Wg, int N, long long T, float divisor, int c_begin,
               int c_end) {
    extern __shared__ float2 smem_f2[];
    C2<float>* a = reinterpret_cast<C2<float>*>(smem_f2);
    const C2<float>* W = reinterpret_cast<const C2<float>*>(Wg);
    const C2<float>* H = reinterpret_cast<const C2<float>*>(Hg);
    const int t = threadIdx.x;
    const int V = kF - N + 1;  // valid outputs per segment
    const long long s0 = (long long)blockIdx.x * V;
    const long long w0 = s0 - (N - 1);  // call-relative index of window sample 0 (>= -hist_pad)
    const int chA = c_begin + 2 * blockIdx.y, chB = chA + 1;
    const bool hasB = chB < c_end;
    const float* rowA = U + (long long)chA * u_stride + hist_pad;
    const float* rowB = U + (long long)(hasB ? chB : chA) * u_stride + hist_pad;

    // forward pass 1 (M = F, radix 8): operands straight from global memory
    {
        constexpr int L = kF / 8;
#pragma unroll 1
        for (int k = 0; k < kF / 8 / kNT; k++) {
            const int j = t + kNT * k;
            C2<float> v[8], w[8];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const long long n = w0 + j + r * L;
                const bool in = n < T;
                v[r].x = in ? __ldg(rowA + n) : 0.0f;
                v[r].y = (in && hasB) ? __ldg(rowB + n) : 0.0f;
            }
            twiddles8(W, j, w);
            fft_dif<8>(v);
            a[pad(j)] = v[0];
#pragma unroll
            for (int s = 1; s < 8; s++) a[pad(j + bitrev3(s) * L)] = cmul(v[s], w[bitrev3(s)]);
        }
    }
    __syncthreads();
    fwd_pass8<kF / 8>(a, W, t);
    __syncthreads();
    fwd_pass8<kF / 64>(a, W, t);
    __syncthreads();
    // forward pass 4 (radix 16, no twiddles) . spectrum product . inverse pass 4, all in registers
#pragma unroll 1
    for (int k = 0; k < kF / 16 / kNT; k++) {
        const int u = t + kNT * k, base = 16 * u;
        C2<float> v[16];
#pragma unroll
        for (int s = 0; s < 16; s++) v[s] = a[pad(base) + s];
        fft_dif<16>(v);
        const float4* h4 = reinterpret_cast<const float4*>(H + base);
#pragma unroll
        for (int s = 0; s < 16; s += 2) {
            const float4 h = __ldg(h4 + s / 2);
            v[s] = cmul(v[s], C2<float>{h.x, h.y});
            v[s + 1] = cmul(v[s + 1], C2<float>{h.z, h.w});
        }
        ifft_dit<16>(v);
#pragma unroll
        for (int s = 0; s < 16; s++) a[pad(base) + s] = v[s];
    }
    __syncthreads();
    inv_pass8<kF / 64>(a, W, t);
    __syncthreads();
    inv_pass8<kF / 8>(a, W, t);
    __syncthreads();
    // inverse pass 1: results straight to global memory (only the V valid samples)
    {
        constexpr int L = kF / 8;
        float* outA = Y + (long long)chA * y_stride;
        float* outB = Y + (long long)chB * y_stride;
#pragma unroll 1
        for (int k = 0; k < kF / 8 / kNT; k++) {
            const int j = t + kNT * k;
            C2<float> v[8], w[8];
            twiddles8(W, j, w);
            v[0] = a[pad(j)];
#pragma unroll
            for (int s = 1; s < 8; s++) v[s] = cmulc(a[pad(j + bitrev3(s) * L)], w[bitrev3(s)]);
            ifft_dit<8>(v);
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int n = j + r * L;
                const long long o = s0 + n - (N - 1);
                if (n >= N - 1 && o < T) {
                    outA[o] = __fmul_rn(v[r].x, divisor);
                    if (hasB) outB[o] = __fmul_rn(v[r].y, divisor);
                }
            }
        }
    }
}

// ---- spectrum of h in the transform's own output order (f64), scaled by 1/F -----------------------------
__global__ void __launch_bounds__(kNT, 1)
fir_spectrum_kernel(const double* __restrict__ taps_rev, int N, const double2* __restrict__ Wd, float2* __restrict__ Hout) {
    extern __shared__ double2 smem_d2[];
    C2<double>* a = reinterpret_cast<C2<double>*>(smem_d2);
    const C2<double>* W = reinterpret_cast<const C2<double>*>(Wd);
    const int t = threadIdx.x;
    for (int n = t; n < kF; n += kNT) a[pad(n)] = C2<double>{n < N ? taps_rev[N - 1 - n] : 0.0, 0.0};  // h[n] = taps[N-1-n]
    __syncthreads();
    fwd_pass8<kF>(a, W, t);
    __syncthreads();
    fwd_pass8<kF / 8>(a, W, t);
    __syncthreads();
    fwd_pass8<kF / 64>(a, W, t);
    __syncthreads();
    for (int k = 0; k < kF / 16 / kNT; k++) {
        const int u = t + kNT * k, base = 16 * u;
        C2<double> v[16];
#pragma unroll
        for (int s = 0; s < 16; s++) v[s] = a[pad(base) + s];
        fft_dif<16>(v);
#pragma unroll
        for (int s = 0; s < 16; s++) Hout[base + s] = make_float2((float)(v[s].x / kF), (float)(v[s].y / kF));
    }
}

struct Tables {
    float2* Wf = nullptr;
    double2* Wd = nullptr;
    int device = -1;
};
Tables g_tab;

int ensure_tables() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (g_tab.Wf && g_tab.device == dev) return 0;
    std::vector<float2> wf(kF);
    std::vector<double2> wd(kF);
    for (int k = 0; k < kF; k++) {
        const double ang = -2.0 * M_PI * (double)k / (double)kF;
        wd[k] = make_double2(std::cos(ang), std::sin(ang));
        wf[k] = make_float2((float)wd[k].x, (float)wd[k].y);
    }
    cudaError_t e;
    if ((e = cudaMalloc(&g_tab.Wf, kF * sizeof(float2))) != cudaSuccess) return (int)e;
    if ((e = cudaMalloc(&g_tab.Wd, kF * sizeof(double2))) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpy(g_tab.Wf, wf.data(), kF * sizeof(float2), cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpy(g_tab.Wd, wd.data(), kF * sizeof(double2), cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
    g_tab.device = dev;
    return 0;
}

}  // namespace

int fir_fft_max_taps() { return kF / 2 + 1; }

int launch_fir_fft(const FirPlan& fp, const float* U, int64_t u_stride, float* Y, int64_t y_stride, int c_begin, int c_end,
                   int64_t T, int64_t started, cudaStream_t st, int* n_launches) {
    (void)started;
    if (fp.log2F != kLog2F || fp.n_taps > fir_fft_max_taps() || fp.n_taps < 1) return (int)cudaErrorInvalidValue;
    int rc = ensure_tables();
    if (rc) return rc;
    static bool configured = false;
    const int smem = kPadded * (int)sizeof(float2);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fir_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const int V = kF - fp.n_taps + 1;
    const long long n_seg = (T + V - 1) / V;
    const int pairs = (c_end - c_begin + 1) / 2;
    for (int p0 = 0; p0 < pairs; p0 += 65535) {
        dim3 grid((unsigned)n_seg, (unsigned)std::min(65535, pairs - p0));
        fir_fft_kernel<<<grid, kNT, smem, st>>>(U, u_stride, fp.hist_pad, Y, y_stride, fp.H, g_tab.Wf, fp.n_taps, T, fp.divisor,
                                               c_begin + 2 * p0, c_end);
        if (n_launches) *n_launches += 1;
    }
    return (int)cudaGetLastError();
}

int fir_prepare_spectrum(int log2F, const double* taps_rev_dev, int n_taps, float2* H_dev, void* stream) {
    if (log2F != kLog2F || n_taps > fir_fft_max_taps()) return (int)cudaErrorInvalidValue;
    int rc = ensure_tables();
    if (rc) return rc;
    const int smem = kPadded * (int)sizeof(double2);
    cudaError_t e = cudaFuncSetAttribute(fir_spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    fir_spectrum_kernel<<<1, kNT, smem, (cudaStream_t)stream>>>(taps_rev_dev, n_taps, g_tab.Wd, H_dev);
    return (int)cudaGetLastError();
}

}  // namespace dspb
