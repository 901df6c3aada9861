// exact_math.cuh — f32 arithmetic that must round exactly like the reference's (rustc: IEEE, no FMA
// contraction, no reassociation).  The *_rn intrinsics are never contracted by nvcc.
#pragma once
#include <cuda_runtime.h>

namespace dspb {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dv(float a, float b) { return __fdiv_rn(a, b); }

// Correctly rounded a / b for a divisor known on the host (b > 0, r = RN(1/b) in IEEE f32).
// Markstein's sequence: q0 = RN(a*r); rem = a - q0*b (one FMA); q1 = RN(q0 + rem*r).  Measured on the
// device over all 2^32 dividends (tests/cuda/div_explore.cu): q1 == RN(a/b) for EVERY dividend whose
// biased exponent lies in [30, 220]; the only failures are in the underflow fringe (rem loses bits),
// the overflow fringe, +-inf and -0.  So the caller keeps the running min of |q1| and max of |a| over
// its chunk and accepts the fast results only if min|q1| >= 2^-90 and max|a| <= 2^90 (a chunk with
// zeros, denormal tails or infinities is redone with __fdiv_rn, which is always right).  The engine
// additionally enables this path per divisor only after verify_const_div() enumerated all 2^32
// dividends under exactly this acceptance rule with zero mismatches.
struct ConstDiv {
    float b, r;
};
constexpr float kDivLo = 8.0779357e-28f;  // 2^-90
constexpr float kDivHi = 1.2379400e27f;   // 2^90
__device__ __forceinline__ float div_const(float a, const ConstDiv d, float& mn_q, float& mx_a) {
    const float q0 = __fmul_rn(a, d.r);
    const float rem = __fmaf_rn(-q0, d.b, a);
    const float q1 = __fmaf_rn(rem, d.r, q0);
    mn_q = fminf(mn_q, fabsf(q1));
    mx_a = fmaxf(mx_a, fabsf(a));
    return q1;
}
__device__ __forceinline__ bool div_const_accept(float mn_q, float mx_a) { return mn_q >= kDivLo && mx_a <= kDivHi; }

}  // namespace dspb
