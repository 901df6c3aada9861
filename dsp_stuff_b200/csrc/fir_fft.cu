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
struct __align__(2 * sizeof(T)) C2 {
    T x, y;
};
__device__ __forceinline__ C2<double> operator+(C2<double> a, C2<double> b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ C2<double> operator-(C2<double> a, C2<double> b) { return {a.x - b.x, a.y - b.y}; }
// f32: one packed FADD2 per complex add/sub (sm_100 add/sub.f32x2), half the issue slots of two FADDs
__device__ __forceinline__ unsigned long long pack2(C2<float> a) {
    return (unsigned long long)__float_as_uint(a.x) | ((unsigned long long)__float_as_uint(a.y) << 32);
}
__device__ __forceinline__ C2<float> unpack2(unsigned long long r) {
    return {__uint_as_float((unsigned)r), __uint_as_float((unsigned)(r >> 32))};
}
__device__ __forceinline__ C2<float> operator+(C2<float> a, C2<float> b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pack2(a)), "l"(pack2(b)));
    return unpack2(d);
}
__device__ __forceinline__ C2<float> operator-(C2<float> a, C2<float> b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pack2(a)), "l"(pack2(b)));
    return unpack2(d);
}
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

// Twiddle source.  W[k] = exp(-2 pi i k / F) factorised as coarse[k >> 4] * fine[k & 15] (512 + 16
// entries, 4.1 KB of shared memory instead of an L2-resident 64 KB table).
constexpr int kCoarse = kF / 16, kFine = 16;
template <typename T>
struct Tw {
    const C2<T>* coarse;  // [512] exp(-2 pi i m / 512)
    const C2<T>* fine;    // [16]  exp(-2 pi i b / 8192)
    __device__ __forceinline__ C2<T> at(int k) const {
        k &= kF - 1;
        const C2<T> c = coarse[k >> 4];
        if ((k & 15) == 0) return c;
        return cmul(c, fine[k & 15]);
    }
};
// twiddles w^q, q = 1..7, from three look-ups (w, w^2, w^4) and four products.  STEP = F/M: every index is
// a multiple of STEP, so look-ups whose index is a multiple of 16 need no fine factor.
template <int STEP, typename T>
__device__ __forceinline__ void twiddles8(const Tw<T>& W, int k1, C2<T> (&w)[8]) {
    if constexpr (STEP % 16 == 0) {
        w[1] = W.coarse[(k1 >> 4) & (kCoarse - 1)];
        w[2] = W.coarse[(k1 >> 3) & (kCoarse - 1)];
        w[4] = W.coarse[(k1 >> 2) & (kCoarse - 1)];
    } else if constexpr (STEP % 8 == 0) {
        w[1] = W.at(k1);
        w[2] = W.coarse[(k1 >> 3) & (kCoarse - 1)];
        w[4] = W.coarse[(k1 >> 2) & (kCoarse - 1)];
    } else {
        w[1] = W.at(k1);
        w[2] = W.at(2 * k1);
        w[4] = W.at(4 * k1);
    }
    w[3] = cmul(w[1], w[2]);
    w[5] = cmul(w[1], w[4]);
    w[6] = cmul(w[2], w[4]);
    w[7] = cmul(w[3], w[4]);
}
template <typename T, typename TG>
__device__ __forceinline__ Tw<T> load_tables(C2<T>* sm, const TG* __restrict__ Wg, int t) {
    // sm: [512 + 16]; Wg: the full F-entry table in global memory
    for (int i = t; i < kCoarse; i += kNT) sm[i] = C2<T>{(T)Wg[16 * i].x, (T)Wg[16 * i].y};
    if (t < kFine) sm[kCoarse + t] = C2<T>{(T)Wg[t].x, (T)Wg[t].y};
    return Tw<T>{sm, sm + kCoarse};
}

// forward radix-8 pass on the shared array: sub-transform size M, L = M/8.  Butterfly u = t + 256 k of
// thread t has j = u % L; for L <= 256 that is the same for every k, so the twiddles are loop-invariant.
template <int M, typename T>
__device__ __forceinline__ void fwd_pass8(C2<T>* a, const Tw<T>& W, int t) {
    constexpr int L = M / 8;
    static_assert(L % 16 == 0, "constant padded stride needs L % 16 == 0");
    constexpr int LP = L + L / 16;  // pad(base + r*L) == pad(base) + r*LP
    constexpr bool kInvariant = (kNT % L) == 0;
    C2<T> w[8];
    if constexpr (kInvariant) twiddles8<kF / M>(W, (t % L) * (kF / M), w);
#pragma unroll 1
    for (int k = 0; k < kF / 8 / kNT; k++) {
        const int u = t + kNT * k, b = u / L, j = u % L, base = b * M + j;
        C2<T>* p = a + pad(base);
        C2<T> v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = p[r * LP];
        if constexpr (!kInvariant) twiddles8<kF / M>(W, j * (kF / M), w);
        fft_dif<8>(v);
        p[0] = v[0];
#pragma unroll
        for (int s = 1; s < 8; s++) p[bitrev3(s) * LP] = cmul(v[s], w[bitrev3(s)]);
    }
}
template <int M, typename T>
__device__ __forceinline__ void inv_pass8(C2<T>* a, const Tw<T>& W, int t) {
    constexpr int L = M / 8;
    constexpr int LP = L + L / 16;
    constexpr bool kInvariant = (kNT % L) == 0;
    C2<T> w[8];
    if constexpr (kInvariant) twiddles8<kF / M>(W, (t % L) * (kF / M), w);
#pragma unroll 1
    for (int k = 0; k < kF / 8 / kNT; k++) {
        const int u = t + kNT * k, b = u / L, j = u % L, base = b * M + j;
        C2<T>* p = a + pad(base);
        C2<T> v[8];
        if constexpr (!kInvariant) twiddles8<kF / M>(W, j * (kF / M), w);
        v[0] = p[0];
#pragma unroll
        for (int s = 1; s < 8; s++) v[s] = cmulc(p[bitrev3(s) * LP], w[bitrev3(s)]);
        ifft_dit<8>(v);
#pragma unroll
        for (int r = 0; r < 8; r++) p[r * LP] = v[r];
    }
}

// ---- the convolution kernel --------------------------------------------------------------------------
// Ne = effective tap count: N padded with zero taps so that Ne - 1 is a multiple of 4.  Then every window
// start (s0 - (Ne-1), s0 a multiple of V = F - Ne + 1) and every output run is 16-byte aligned and the
// first / last pass move two consecutive samples per 64-bit access.
// post: optional epilogue "acc = 0.0 + y; acc /= post_nf" = the fan-in average of a sink fed only by this
// node (node.rs:162-194, nodes/output.rs:223), so no separate kernel has to touch the output again.
__global__ void __launch_bounds__(kNT, 3)
fir_fft_kernel(const float* __restrict__ U, long long u_stride, int hist_pad, float* __restrict__ Y, long long y_stride,
               const float2* __restrict__ Hg, const float2* __restrict__ Wg, int Ne, long long T, float divisor, float post_nf,
               int c_begin, int c_end) {
    extern __shared__ float2 smem_f2[];
    C2<float>* a = reinterpret_cast<C2<float>*>(smem_f2);
    const int t = threadIdx.x;
    const Tw<float> W = load_tables<float>(a + kPadded, Wg, t);
    const C2<float>* H = reinterpret_cast<const C2<float>*>(Hg);
    const int V = kF - Ne + 1;  // valid outputs per segment (multiple of 4)
    const long long s0 = (long long)blockIdx.x * V;
    const long long w0 = s0 - (Ne - 1);  // call-relative index of window sample 0 (>= -hist_pad), multiple of 4
    const int chA = c_begin + 2 * blockIdx.y, chB = chA + 1;
    const bool hasB = chB < c_end;
    const float* rowA = U + (long long)chA * u_stride + hist_pad + w0;
    const float* rowB = U + (long long)(hasB ? chB : chA) * u_stride + hist_pad + w0;
    const long long lim = T - w0;  // window samples >= lim lie beyond this call's input: zeros
    __syncthreads();               // twiddle tables visible

    // forward pass 1 (M = F, radix 8): operands straight from global memory, two butterflies per step
    {
        constexpr int L = kF / 8;
#pragma unroll 1
        for (int k = 0; k < kF / 8 / kNT / 2; k++) {
            const int j = 2 * (t + kNT * k);
            C2<float> v0[8], v1[8], w[8];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int n = j + r * L;
                float2 xa = make_float2(0.f, 0.f), xb = make_float2(0.f, 0.f);
                if (n < lim) {  // T, w0 and n are even: the pair is inside or outside together
                    xa = __ldg(reinterpret_cast<const float2*>(rowA + n));
                    if (hasB) xb = __ldg(reinterpret_cast<const float2*>(rowB + n));
                }
                v0[r] = C2<float>{xa.x, xb.x};
                v1[r] = C2<float>{xa.y, xb.y};
            }
            constexpr int LP = L + L / 16;
            C2<float>* p = a + pad(j);  // j is even: j and j + 1 share the padding offset
            twiddles8<1>(W, j, w);
            fft_dif<8>(v0);
            p[0] = v0[0];
#pragma unroll
            for (int s = 1; s < 8; s++) p[bitrev3(s) * LP] = cmul(v0[s], w[bitrev3(s)]);
            twiddles8<1>(W, j + 1, w);
            fft_dif<8>(v1);
            p[1] = v1[0];
#pragma unroll
            for (int s = 1; s < 8; s++) p[1 + bitrev3(s) * LP] = cmul(v1[s], w[bitrev3(s)]);
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
        float4 h[8];
        // H is stored [k][s / 2][thread] (float4 = two consecutive bins): every request of a warp is one
        // contiguous 512-byte run.  (In natural order each thread owns a 128-byte line: 32 lines per request,
        // which alone kept the L1 data pipe 84 % busy -- ncu l1tex__data_pipe_lsu_wavefronts.)
        const float4* h4 = reinterpret_cast<const float4*>(H) + (k * 8) * kNT + t;
#pragma unroll
        for (int s = 0; s < 8; s++) h[s] = __ldg(h4 + s * kNT);
#pragma unroll
        for (int s = 0; s < 16; s++) v[s] = a[pad(base) + s];
        fft_dif<16>(v);
#pragma unroll
        for (int s = 0; s < 16; s += 2) {
            v[s] = cmul(v[s], C2<float>{h[s / 2].x, h[s / 2].y});
            v[s + 1] = cmul(v[s + 1], C2<float>{h[s / 2].z, h[s / 2].w});
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
    // inverse pass 1: results straight to global memory (only the V valid samples), two samples per store
    {
        constexpr int L = kF / 8;
        float* outA = Y + (long long)chA * y_stride + s0 - (Ne - 1);
        float* outB = Y + (long long)chB * y_stride + s0 - (Ne - 1);
        const bool post = post_nf != 0.0f;
#pragma unroll 1
        for (int k = 0; k < kF / 8 / kNT / 2; k++) {
            const int j = 2 * (t + kNT * k);
            C2<float> v0[8], v1[8], w[8];
            constexpr int LP = L + L / 16;
            const C2<float>* p = a + pad(j);
            twiddles8<1>(W, j, w);
            v0[0] = p[0];
#pragma unroll
            for (int s = 1; s < 8; s++) v0[s] = cmulc(p[bitrev3(s) * LP], w[bitrev3(s)]);
            ifft_dit<8>(v0);
            twiddles8<1>(W, j + 1, w);
            v1[0] = p[1];
#pragma unroll
            for (int s = 1; s < 8; s++) v1[s] = cmulc(p[1 + bitrev3(s) * LP], w[bitrev3(s)]);
            ifft_dit<8>(v1);
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int n = j + r * L;
                if (n >= Ne - 1 && n < lim) {
                    float2 ya = make_float2(__fmul_rn(v0[r].x, divisor), __fmul_rn(v1[r].x, divisor));
                    float2 yb = make_float2(__fmul_rn(v0[r].y, divisor), __fmul_rn(v1[r].y, divisor));
                    if (post) {
                        ya.x = __fdiv_rn(__fadd_rn(0.0f, ya.x), post_nf); ya.y = __fdiv_rn(__fadd_rn(0.0f, ya.y), post_nf);
                        yb.x = __fdiv_rn(__fadd_rn(0.0f, yb.x), post_nf); yb.y = __fdiv_rn(__fadd_rn(0.0f, yb.y), post_nf);
                    }
                    *reinterpret_cast<float2*>(outA + n) = ya;
                    if (hasB) *reinterpret_cast<float2*>(outB + n) = yb;
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
    const int t = threadIdx.x;
    const Tw<double> W = load_tables<double>(a + kPadded, Wd, t);
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
        for (int s = 0; s < 16; s++)  // [k][s / 2][thread] float4 order, see fir_fft_kernel
            Hout[((k * 8 + (s >> 1)) * kNT + t) * 2 + (s & 1)] = make_float2((float)(v[s].x / kF), (float)(v[s].y / kF));
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
static int effective_taps(int n) { return (n - 1 + 3) / 4 * 4 + 1; }  // Ne - 1 multiple of 4

int launch_fir_fft(const FirPlan& fp, const float* U, int64_t u_stride, float* Y, int64_t y_stride, int c_begin, int c_end,
                   int64_t T, int64_t started, cudaStream_t st, int* n_launches) {
    (void)started;
    if (fp.log2F != kLog2F || fp.n_taps > fir_fft_max_taps() || fp.n_taps < 1) return (int)cudaErrorInvalidValue;
    int rc = ensure_tables();
    if (rc) return rc;
    static bool configured = false;
    const int smem = (kPadded + kCoarse + kFine) * (int)sizeof(float2);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fir_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const int Ne = effective_taps(fp.n_taps);
    const int V = kF - Ne + 1;
    const long long n_seg = (T + V - 1) / V;
    const int pairs = (c_end - c_begin + 1) / 2;
    for (int p0 = 0; p0 < pairs; p0 += 65535) {
        dim3 grid((unsigned)n_seg, (unsigned)std::min(65535, pairs - p0));
        fir_fft_kernel<<<grid, kNT, smem, st>>>(U, u_stride, fp.hist_pad, Y, y_stride, fp.H, g_tab.Wf, Ne, T, fp.divisor, fp.post_nf,
                                               c_begin + 2 * p0, c_end);
        if (n_launches) *n_launches += 1;
    }
    return (int)cudaGetLastError();
}

int fir_prepare_spectrum(int log2F, const double* taps_rev_dev, int n_taps, float2* H_dev, void* stream) {
    if (log2F != kLog2F || n_taps > fir_fft_max_taps()) return (int)cudaErrorInvalidValue;
    int rc = ensure_tables();
    if (rc) return rc;
    const int smem = (kPadded + kCoarse + kFine) * (int)sizeof(double2);
    cudaError_t e = cudaFuncSetAttribute(fir_spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    fir_spectrum_kernel<<<1, kNT, smem, (cudaStream_t)stream>>>(taps_rev_dev, n_taps, g_tab.Wd, H_dev);
    return (int)cudaGetLastError();
}

}  // namespace dspb
