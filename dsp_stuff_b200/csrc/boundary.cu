// boundary.cu — the data-format steps either side of the effect path, batched over streams:
//   * stereo fold on capture: devices.rs:244-262 `do_read_2`: interleaved frames [a, b] -> mono a + b (f32 add,
//     NOT an average) before the samples enter the graph;
//   * mono -> stereo duplicate on playback: devices.rs:443-500 `do_write_2`: every mono sample fills both slots of
//     its output frame (`o.fill(x)`).
//   * the 48 kHz -> device-rate converter in front of that duplicate: devices.rs:550-556 builds dasp_signal's
//     `Converter::from_hz_to_hz(.., Sinc::new(Fixed::from([0.0; 16])), 48_000.0, target)`, do_write_2 pulls one frame per
//     stereo output frame (resample_dup_kernel below; the two crates are un-vendored: restated, parity unpinned).
// Fold and duplicate are pure HBM streams: 12 algorithmic bytes per mono channel-sample (8 + 4), 128-bit accesses.
#include <cuda_runtime.h>

#include <cstdint>

#include "plan.h"

namespace dspb {
namespace {

__global__ void __launch_bounds__(256)
fold_stereo_kernel(const float4* __restrict__ in, float2* __restrict__ out, long long n_pairs) {  // 2 frames per thread
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += stride) {
        const float4 f = __ldcs(in + i);
        __stcs(out + i, make_float2(__fadd_rn(f.x, f.y), __fadd_rn(f.z, f.w)));
    }
}
__global__ void __launch_bounds__(256)
dup_stereo_kernel(const float2* __restrict__ in, float4* __restrict__ out, long long n_pairs) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += stride) {
        const float2 m = __ldcs(in + i);
        __stcs(out + i, make_float4(m.x, m.x, m.y, m.y));
    }
}
__global__ void fold_stereo_tail(const float* in, float* out, long long first, long long n) {
    const long long i = first + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fadd_rn(in[2 * i], in[2 * i + 1]);
}
__global__ void dup_stereo_tail(const float* in, float* out, long long first, long long n) {
    const long long i = first + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) { out[2 * i] = in[i]; out[2 * i + 1] = in[i]; }
}

// Sample-rate conversion + duplicate.  Which source frames an output frame sees, the fractional position x and the 16
// window weights sinc(a) hann(a) depend only on the output index, not on the channel: the host walks dasp's Converter
// control flow once per call (engine.cpp resample_plan: f64, the same libm the oracle uses) and uploads, per output frame
// m, {pushes so far, Sinc::idx} and the weights.  A thread then evaluates Sinc::interpolate for one (channel, frame):
// v += f32(w f64(frame)) left tap, right tap, n = 0 .. max_depth - 1, exactly the reference's order and roundings.
// frames[i] of the 16-frame ring (i taken modulo 16, like dasp's ring_buffer::Fixed): source frame P - 16 + i where P
// frames have been pushed; frames pushed in earlier calls come from hist[C x 16], frames beyond the input are zeros
// (CountingSignal::next past its buffer, devices.rs:380-386).
__global__ void __launch_bounds__(256)
resample_dup_kernel(const float* __restrict__ in, long long n_in, const float* __restrict__ hist, const int2* __restrict__ meta,
                    const double* __restrict__ w, float2* __restrict__ out, long long n_out, int channels) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (m >= n_out || c >= channels) return;
    const int2 mt = meta[m];
    const int P = mt.x, idx = mt.y;
    const int depth = 8, nl = idx, nr = idx + 1;
    const int rightmost = nl + depth, leftmost = nr - depth;
    const int max_depth = rightmost >= 16 ? 16 - depth : (leftmost < 0 ? depth + leftmost : depth);
    const float* src = in + (long long)c * n_in;
    const float* hs = hist + (long long)c * 16;
    auto frame = [&](int i) -> float {
        const long long r = (long long)P - 16 + (i & 15);   // relative to this call's first input sample
        if (r >= 0) return r < n_in ? __ldg(src + r) : 0.0f;
        return hs[16 + r];                                  // r in [-16, -1]: pushed by earlier calls (zeros at the very start)
    };
    const double* wm = w + m * 16;
    float v = 0.0f;
    for (int n = 0; n < max_depth; n++) {
        v = __fadd_rn(v, __double2float_rn(__dmul_rn(wm[2 * n], (double)frame(nl - n))));
        v = __fadd_rn(v, __double2float_rn(__dmul_rn(wm[2 * n + 1], (double)frame(nr + n))));
    }
    out[(long long)c * n_out + m] = make_float2(v, v);      // o.fill(x), devices.rs:487-491
}
// the 16 most recently pushed frames after a call that pushed `pushed` frames
__global__ void resample_hist_kernel(const float* __restrict__ in, long long n_in, const float* __restrict__ hist_old,
                                     float* __restrict__ hist_new, long long pushed, int channels) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels) return;
    for (int i = 0; i < 16; i++) {
        const long long r = pushed - 16 + i;
        float f;
        if (r >= 0) f = r < n_in ? in[(long long)c * n_in + r] : 0.0f;
        else f = hist_old[(long long)c * 16 + 16 + r];
        hist_new[(long long)c * 16 + i] = f;
    }
}

int grid_for(long long n_pairs) {
    long long g = (n_pairs + 255) / 256;
    const long long cap = 148 * 16;  // a multiple of the SM count; grid-stride beyond that
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

// total_mono = channels * n_samples (rows are contiguous, so the whole batch is one flat stream)
int launch_fold_stereo(const float* interleaved, float* mono, long long total_mono, cudaStream_t st) {
    const bool aligned = (((uintptr_t)interleaved & 15) == 0) && (((uintptr_t)mono & 7) == 0);
    const long long n_pairs = aligned ? total_mono / 2 : 0;
    if (n_pairs) fold_stereo_kernel<<<grid_for(n_pairs), 256, 0, st>>>(reinterpret_cast<const float4*>(interleaved), reinterpret_cast<float2*>(mono), n_pairs);
    const long long rest = total_mono - 2 * n_pairs;
    if (rest) fold_stereo_tail<<<(unsigned)((rest + 255) / 256), 256, 0, st>>>(interleaved, mono, 2 * n_pairs, total_mono);
    return (int)cudaGetLastError();
}
int launch_dup_stereo(const float* mono, float* interleaved, long long total_mono, cudaStream_t st) {
    const bool aligned = (((uintptr_t)interleaved & 15) == 0) && (((uintptr_t)mono & 7) == 0);
    const long long n_pairs = aligned ? total_mono / 2 : 0;
    if (n_pairs) dup_stereo_kernel<<<grid_for(n_pairs), 256, 0, st>>>(reinterpret_cast<const float2*>(mono), reinterpret_cast<float4*>(interleaved), n_pairs);
    const long long rest = total_mono - 2 * n_pairs;
    if (rest) dup_stereo_tail<<<(unsigned)((rest + 255) / 256), 256, 0, st>>>(mono, interleaved, 2 * n_pairs, total_mono);
    return (int)cudaGetLastError();
}

int launch_resample_dup(const float* mono, long long n_in, const float* hist_old, float* hist_new, const void* meta, const double* w,
                        float* interleaved, long long n_out, long long pushed, int channels, cudaStream_t st) {
    if (n_out > 0) {
        for (int c0 = 0; c0 < channels; c0 += 65535) {
            const int cc = channels - c0 < 65535 ? channels - c0 : 65535;
            dim3 grid((unsigned)((n_out + 255) / 256), (unsigned)cc);
            resample_dup_kernel<<<grid, 256, 0, st>>>(mono + (long long)c0 * n_in, n_in, hist_old + (long long)c0 * 16,
                                                     reinterpret_cast<const int2*>(meta), w,
                                                     reinterpret_cast<float2*>(interleaved) + (long long)c0 * n_out, n_out, cc);
        }
    }
    resample_hist_kernel<<<(channels + 127) / 128, 128, 0, st>>>(mono, n_in, hist_old, hist_new, pushed, channels);
    return (int)cudaGetLastError();
}

}  // namespace dspb
