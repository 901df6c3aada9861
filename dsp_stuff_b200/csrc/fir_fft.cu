// fir_fft.cu — overlap-save FFT convolution for the Fir node (placeholder until the kernel lands:
// launch_fir_fft reports "not supported" so the engine refuses FIR_FFT instead of silently
// computing something else).
#include <cuda_runtime.h>

#include "plan.h"

namespace dspb {

int launch_fir_fft(const FirPlan&, const float*, int64_t, float*, int64_t, int, int, int64_t, int64_t, cudaStream_t, int*) {
    return (int)cudaErrorNotSupported;
}

int fir_prepare_spectrum(int, const double*, int, float2*, void*) { return 0; }

}  // namespace dspb
