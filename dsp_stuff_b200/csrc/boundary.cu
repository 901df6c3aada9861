// boundary.cu — the data-format steps either side of the effect path, batched over streams:
//   * stereo fold on capture: devices.rs:244-262 `do_read_2`: interleaved frames [a, b] -> mono a + b (f32 add,
//     NOT an average) before the samples enter the graph;
//   * mono -> stereo duplicate on playback: devices.rs:443-500 `do_write_2`: every mono sample fills both slots of
//     its output frame (`o.fill(x)`).
// The 48 kHz -> device-rate sinc resampler that sits in front of the duplicate (dasp_interpolate Sinc<[f32; 16]>,
// an un-vendored dependency) is out of scope: streams here stay at the graph's 48 kHz.
// Both are pure HBM streams: 12 algorithmic bytes per mono channel-sample (8 + 4), 128-bit accesses.
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

}  // namespace dspb
