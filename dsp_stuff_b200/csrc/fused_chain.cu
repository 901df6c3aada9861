// fused_chain.cu — the fused effect-segment kernel (sm_100a).
//
// One launch runs a whole fused segment of the node graph (every node except Fir) for a range of
// channels over all samples of the call.  Replaces, for those nodes, the reference's per-node task
// loop + SimpleNode::process bodies (dsp-stuff/src/node.rs:267-352, nodes/*.rs).
//
// Geometry: a CTA owns G channels for the whole call and walks time in tiles of S = 4096/G samples,
// so every sequential dependency (IIR state, comb ring, Fuzz/SignalGen 128-sample blocks) stays
// inside one CTA.  256 threads; thread (g, j) holds 16 consecutive samples of channel g in
// registers (`acc`).  HBM traffic is 128-bit per thread (4 x LDG/STG.128 per 16 samples); the next
// tile's inputs and ring reads are prefetched with cp.async into shared memory while the current
// tile computes.  Recurrences are executed lane = channel, strictly in time order with FMA-free
// f32 arithmetic (bit-identical to the reference), after a conflict-free shared-memory transpose.
//
// Arithmetic that must match the reference bit for bit uses __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn
// so nvcc can never contract it into FMAs (rustc does not).
#include <cuda_runtime.h>

#include "plan.h"

namespace dspb {
namespace {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dv(float a, float b) { return __fdiv_rn(a, b); }

// distort.rs:53-61
__device__ __forceinline__ float clipf(float s) { return s < -1.0f ? -1.0f : (s > 1.0f ? 1.0f : s); }
// f32::signum: NaN -> NaN, else copysign(1, x)
__device__ __forceinline__ float signumf(float x) { return x != x ? x : copysignf(1.0f, x); }
__device__ __forceinline__ float clamp01(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }

enum DistortMode { HardClip, SoftClip, Tanh, RecipSoftClip, Fuzz, Sin, Atan, Square, Chebyshev4 };

__device__ __forceinline__ float shape(int mode, float x, float l) {
    if (l < 0.001f) return x;
    switch (mode) {
        case HardClip: return dv(clipf(mul(x, l)), l);
        case SoftClip: {
            float s = mul(x, l);
            if (s > 1.0f) s = 2.0f / 3.0f;
            else if (s >= -1.0f && s <= 1.0f) s = sub(s, dv(mul(mul(s, s), s), 3.0f));
            else s = -2.0f / 3.0f;
            return dv(clipf(s), l);
        }
        case Tanh: return tanhf(mul(x, l));
        case RecipSoftClip: return mul(signumf(x), sub(1.0f, dv(1.0f, add(mul(fabsf(x), l), 1.0f))));
        case Sin: return sinf(mul(x, l));
        case Atan: return atanf(mul(x, l));
        case Square: { float v = mul(x, l); return mul(mul(v, v), signumf(v)); }
        case Chebyshev4: {
            float v = mul(x, l);
            float v2 = mul(v, v);
            return add(sub(mul(8.0f, mul(v2, v2)), mul(8.0f, v2)), 1.0f);
        }
    }
    return x;
}

// max over the 8 consecutive threads (= one 128-sample reference block) of |v| under
// f32::total_cmp: non-negative floats and NaNs order like their bit patterns.
__device__ __forceinline__ float block128_max_abs(const float (&v)[kChunk]) {
    unsigned m = 0;
#pragma unroll
    for (int i = 0; i < kChunk; i++) m = max(m, __float_as_uint(fabsf(v[i])));
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 4));
    return __uint_as_float(m);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
    asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct DF1 {  // crate biquad 0.4.2 DirectForm1<f32>::run via nodes/biquad.rs:87
    float b0, b1, b2, a1, a2;
    float x1, x2, y1, y2;
    __device__ __forceinline__ void load(const float4 s) { x1 = s.x; x2 = s.y; y1 = s.z; y2 = s.w; }
    __device__ __forceinline__ float4 save() const { return make_float4(x1, x2, y1, y2); }
    __device__ __forceinline__ float step(float x) {
        float out = sub(sub(add(add(mul(b0, x), mul(b1, x1)), mul(b2, x2)), mul(a1, y1)), mul(a2, y2));
        x2 = x1; x1 = x; y2 = y1; y1 = out;
        return out;
    }
};
struct LP1 {  // nodes/low_pass.rs:36-39
    float r, omr, z;
    __device__ __forceinline__ void load(const float4 s) { z = s.x; }
    __device__ __forceinline__ float4 save() const { return make_float4(z, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ float step(float x) { z = add(mul(x, omr), mul(r, z)); return z; }
};
struct HP1 {  // nodes/high_pass.rs:36-39
    float r, omr, z;
    __device__ __forceinline__ void load(const float4 s) { z = s.x; }
    __device__ __forceinline__ float4 save() const { return make_float4(z, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ float step(float x) { z = add(mul(x, omr), mul(r, z)); return sub(x, z); }
};
struct Env {  // dasp_envelope 0.11.0 Detector<f32, Peak<FullWave>>::next via nodes/envelope.rs:50
    float ga, gr, prev;
    __device__ __forceinline__ void load(const float4 s) { prev = s.x; }
    __device__ __forceinline__ float4 save() const { return make_float4(prev, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ float step(float x) {
        float d = fabsf(x);
        float g = prev < d ? ga : gr;
        prev = add(d, mul(sub(prev, d), g));
        return prev;
    }
};

// Swizzled float4 slot of logical float4 index m inside a tile row: thread j writes its four
// float4s to 4j + (k ^ ((j>>1)&3)), which makes both the time-parallel writes (8 lanes, stride 64 B)
// and the lane = channel reads (rows padded by 4 floats) bank-conflict free.
__device__ __forceinline__ int swz(int m) { return (m & ~3) | ((m & 3) ^ ((m >> 3) & 3)); }

template <int G>
struct Geo {
    static constexpr int S = kTile / G;        // samples per channel per tile
    static constexpr int TPC = kThreads / G;   // threads per channel
    static constexpr int ROW = S + 4;          // padded tile row (floats)
};

template <int G, class R>
__device__ __forceinline__ void run_recurrence(R rec, float (&acc)[kChunk], float* tile, float4* sm_state,
                                               int g, int j, int valid_f4) {
    using Q = Geo<G>;
    const int t = threadIdx.x;
    float4* row = reinterpret_cast<float4*>(tile + g * Q::ROW);
#pragma unroll
    for (int k = 0; k < 4; k++)
        row[4 * j + (k ^ ((j >> 1) & 3))] = make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
    __syncthreads();
    if (t < G && valid_f4 > 0) {  // lane = channel, strictly sequential in time
        float4* r = reinterpret_cast<float4*>(tile + t * Q::ROW);
        rec.load(sm_state[t]);
        float4 v = r[swz(0)];
        for (int m = 0; m < valid_f4; m++) {
            float4 nx = v;
            if (m + 1 < valid_f4) nx = r[swz(m + 1)];
            v.x = rec.step(v.x);
            v.y = rec.step(v.y);
            v.z = rec.step(v.z);
            v.w = rec.step(v.w);
            r[swz(m)] = v;
            v = nx;
        }
        sm_state[t] = rec.save();
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float4 v = row[4 * j + (k ^ ((j >> 1) & 3))];
        acc[4 * k] = v.x; acc[4 * k + 1] = v.y; acc[4 * k + 2] = v.z; acc[4 * k + 3] = v.w;
    }
}

template <int G>
__global__ void __launch_bounds__(kThreads, 2)
fused_kernel(const __grid_constant__ Program prog, int c_begin, int c_end, long long T, int n_states) {
    using Q = Geo<G>;
    extern __shared__ float4 smem4[];
    const int t = threadIdx.x;
    const int g = t / Q::TPC;
    const int j = t % Q::TPC;
    const int ch = c_begin + blockIdx.x * G + g;
    const bool ch_ok = ch < c_end;

    // shared memory carve-up
    float4* sm_state = smem4;                                  // [kMaxStates][G]
    float* tile = reinterpret_cast<float*>(sm_state + kMaxStates * G);
    float4* stage = reinterpret_cast<float4*>(tile + (prog.needs_tile ? G * Q::ROW : 0));  // [2][n_prefetch][4][256]
    float4* vregs = stage + 2 * prog.n_prefetch * 4 * kThreads;  // [n_vregs][4][256]

    for (int i = t; i < n_states * G; i += kThreads) {
        int s = i / G, c = c_begin + blockIdx.x * G + (i % G);
        sm_state[s * G + (i % G)] = c < c_end ? reinterpret_cast<const float4*>(prog.states[s])[c] : make_float4(0, 0, 0, 0);
    }

    const long long n_tiles = (T + Q::S - 1) / Q::S;
    const bool fast_ring_all = true;
    (void)fast_ring_all;

    auto issue_prefetch = [&](long long tile_i, int parity) {
        const long long n0 = tile_i * Q::S + (long long)j * kChunk;
        if (ch_ok && tile_i < n_tiles && n0 < T) {
#pragma unroll
            for (int s = 0; s < kMaxPrefetch; s++) {
                if (s >= prog.n_prefetch) break;
                const float* src;
                if (prog.pf_buf[s] >= 0) {
                    const BufDesc& b = prog.bufs[prog.pf_buf[s]];
                    src = b.base + (long long)ch * b.row_stride + n0;
                } else {
                    const RingDesc& r = prog.rings[prog.pf_ring[s]];
                    src = r.base + (long long)ch * r.D + (r.pos + n0) % r.D;
                }
                float4* dst = stage + ((parity * prog.n_prefetch + s) * 4) * kThreads + t;
#pragma unroll
                for (int k = 0; k < 4; k++) cp_async16(dst + k * kThreads, src + 4 * k);
            }
        }
        cp_async_commit();
    };

    issue_prefetch(0, 0);
    __syncthreads();

    for (long long tile_i = 0; tile_i < n_tiles; tile_i++) {
        const int parity = (int)(tile_i & 1);
        const long long n0 = tile_i * Q::S + (long long)j * kChunk;
        const bool active = ch_ok && n0 < T;
        const long long rem = T - tile_i * Q::S;
        const int valid_f4 = (int)((rem < Q::S ? rem : Q::S) / 4);

        __syncthreads();  // ring / vreg / tile hazards across tiles
        issue_prefetch(tile_i + 1, parity ^ 1);
        cp_async_wait<1>();

        float acc[kChunk];
#pragma unroll
        for (int i = 0; i < kChunk; i++) acc[i] = 0.0f;

        for (int ip = 0; ip < prog.n_ops; ip++) {
            const Op& op = prog.ops[ip];
            const int code = op.code;

            // tile-valued parameters (connected control ports)
            auto load_param = [&](int which, float (&P)[kChunk]) {
                if (op.pflags & (1 << which)) {
                    const float4* v = vregs + (op.pv[which] * 4) * kThreads + t;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        float4 q = v[k * kThreads];
                        P[4 * k] = q.x; P[4 * k + 1] = q.y; P[4 * k + 2] = q.z; P[4 * k + 3] = q.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < kChunk; i++) P[i] = op.p[which];
                }
            };

            switch (code) {
                case OP_ZERO: {
#pragma unroll
                    for (int i = 0; i < kChunk; i++) acc[i] = 0.0f;
                } break;
                case OP_LOADG:
                case OP_ADDG:
                case OP_COPYG: {
                    float v[kChunk];
                    if (op.aux) {  // staged by cp.async
                        const float4* s = stage + ((parity * prog.n_prefetch + (op.aux - 1)) * 4) * kThreads + t;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            float4 q = s[k * kThreads];
                            v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
                        }
                    } else if (active) {
                        const BufDesc& b = prog.bufs[op.buf];
                        const float4* p = reinterpret_cast<const float4*>(b.base + (long long)ch * b.row_stride + n0);
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            float4 q = ldg_stream(p + k);
                            v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < kChunk; i++) v[i] = 0.0f;
                    }
                    if (!active) {
#pragma unroll
                        for (int i = 0; i < kChunk; i++) v[i] = 0.0f;
                    }
#pragma unroll
                    for (int i = 0; i < kChunk; i++)
                        acc[i] = code == OP_COPYG ? v[i] : add(code == OP_LOADG ? 0.0f : acc[i], v[i]);
                } break;
                case OP_LOADV:
                case OP_ADDV:
                case OP_COPYV:
                case OP_ADD: {
                    const float4* s = vregs + (op.vreg * 4) * kThreads + t;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        float4 q = s[k * kThreads];
                        if (code == OP_COPYV) {
                            acc[4 * k] = q.x; acc[4 * k + 1] = q.y; acc[4 * k + 2] = q.z; acc[4 * k + 3] = q.w;
                        } else {
                            const bool ld = code == OP_LOADV;
                            acc[4 * k] = add(ld ? 0.0f : acc[4 * k], q.x);
                            acc[4 * k + 1] = add(ld ? 0.0f : acc[4 * k + 1], q.y);
                            acc[4 * k + 2] = add(ld ? 0.0f : acc[4 * k + 2], q.z);
                            acc[4 * k + 3] = add(ld ? 0.0f : acc[4 * k + 3], q.w);
                        }
                    }
                } break;
                case OP_DIVC: {
                    const float nf = op.p[0];
#pragma unroll
                    for (int i = 0; i < kChunk; i++) acc[i] = dv(acc[i], nf);
                } break;
                case OP_SAVEV: {
                    float4* s = vregs + (op.vreg * 4) * kThreads + t;
#pragma unroll
                    for (int k = 0; k < 4; k++) s[k * kThreads] = make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
                } break;
                case OP_STOREG: {
                    if (active) {
                        const BufDesc& b = prog.bufs[op.buf];
                        float4* p = reinterpret_cast<float4*>(b.base + (long long)ch * b.row_stride + n0);
#pragma unroll
                        for (int k = 0; k < 4; k++) stg_stream(p + k, make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]));
                    }
                } break;
                case OP_MODMAP: {
                    const float lo = op.p[0], span = sub(op.p[1], op.p[0]);
#pragma unroll
                    for (int i = 0; i < kChunk; i++) {
                        float y = dv(add(acc[i], 1.0f), 2.0f);
                        acc[i] = add(lo, mul(span, clamp01(y)));
                    }
                } break;
                case OP_GAIN: {
                    float P[kChunk];
                    load_param(0, P);
#pragma unroll
                    for (int i = 0; i < kChunk; i++) acc[i] = mul(acc[i], P[i]);
                } break;
                case OP_DISTORT: {
                    float P[kChunk];
                    load_param(0, P);
                    const int mode = op.mode;
                    if (mode == Fuzz) {  // nodes/distort.rs:146-172, per 128-sample reference block
                        const float mx = block128_max_abs(acc);
                        float z[kChunk];
#pragma unroll
                        for (int i = 0; i < kChunk; i++) {
                            float q = dv(clipf(mul(acc[i], P[i])), mx);
                            z[i] = copysignf(sub(1.0f, expf(copysignf(q, -1.0f))), -1.0f);
                        }
                        const float mz = block128_max_abs(z);
#pragma unroll
                        for (int i = 0; i < kChunk; i++) z[i] = dv(clipf(mul(z[i], mx)), mz);
                        const float my = block128_max_abs(z);
#pragma unroll
                        for (int i = 0; i < kChunk; i++) acc[i] = dv(mul(z[i], mx), my);
                    } else {
#pragma unroll
                        for (int i = 0; i < kChunk; i++) acc[i] = shape(mode, acc[i], P[i]);
                    }
                } break;
                case OP_OVERDRIVE: {
                    float B[kChunk], D[kChunk], L[kChunk];
                    load_param(0, B);
                    load_param(1, D);
                    load_param(2, L);
                    const float FRAC_PI_4 = 0.785398163397448309615660845819875721f;
                    const float FRAC_2_PI = 0.636619772367581343075535053490057448f;
#pragma unroll
                    for (int i = 0; i < kChunk; i++) {
                        const float x = acc[i];
                        if (!(L[i] < 0.001f)) {
                            float c = atanf(mul(FRAC_PI_4, mul(x, B[i])));
                            float d = mul(FRAC_2_PI, c);
                            float mix = add(mul(D[i], d), mul(sub(1.0f, D[i]), x));
                            acc[i] = mul(mix, L[i]);
                        }
                    }
                } break;
                case OP_CHEBY: {
                    const float lp = op.p[0], ln = op.p[1], tp = op.p[2], tn = op.p[3];
#pragma unroll
                    for (int i = 0; i < kChunk; i++) {
                        const float x = acc[i];
                        if (x >= 0.0f) { if (!(lp < 0.001f)) acc[i] = dv(tanhf(mul(x, lp)), tp); }
                        else { if (!(ln < 0.001f)) acc[i] = dv(tanhf(mul(x, ln)), tn); }
                    }
                } break;
                case OP_MIX: {
                    float R[kChunk];
                    load_param(0, R);
                    const float4* s = vregs + (op.vreg * 4) * kThreads + t;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        float4 q = s[k * kThreads];
                        const float b[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const int i = 4 * k + e;
                            acc[i] = add(mul(b[e], R[i]), mul(acc[i], sub(1.0f, R[i])));
                        }
                    }
                } break;
                case OP_COMB: {
                    const RingDesc& r = prog.rings[op.aux & 0xff];
                    const float decay = op.p[0];
                    const int pslot = op.aux >> 8;  // 0 = not staged
                    float* rrow = r.base + (long long)ch * r.D;
                    if ((r.D & 15) == 0 && (r.pos & 15) == 0) {
                        const long long slot = (r.pos + n0) % r.D;
                        float old[kChunk];
                        if (pslot) {
                            const float4* s = stage + ((parity * prog.n_prefetch + (pslot - 1)) * 4) * kThreads + t;
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                float4 q = s[k * kThreads];
                                old[4 * k] = q.x; old[4 * k + 1] = q.y; old[4 * k + 2] = q.z; old[4 * k + 3] = q.w;
                            }
                        } else if (active) {
                            const float4* p = reinterpret_cast<const float4*>(rrow + slot);
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                float4 q = ldg_stream(p + k);
                                old[4 * k] = q.x; old[4 * k + 1] = q.y; old[4 * k + 2] = q.z; old[4 * k + 3] = q.w;
                            }
                        }
                        if (active) {
#pragma unroll
                            for (int i = 0; i < kChunk; i++) acc[i] = add(acc[i], mul(old[i], decay));
                            float4* p = reinterpret_cast<float4*>(rrow + slot);
#pragma unroll
                            for (int k = 0; k < 4; k++) stg_stream(p + k, make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]));
                        }
                    } else if (active) {  // ring length not a multiple of 16: element-wise wrap
                        long long slot = (r.pos + n0) % r.D;
#pragma unroll
                        for (int i = 0; i < kChunk; i++) {
                            float o = __ldcg(rrow + slot);
                            acc[i] = add(acc[i], mul(o, decay));
                            __stcg(rrow + slot, acc[i]);
                            if (++slot == r.D) slot = 0;
                        }
                    }
                } break;
                case OP_BIQUAD: {
                    DF1 f;
                    f.b0 = op.p[0]; f.b1 = op.p[1]; f.b2 = op.p[2]; f.a1 = op.p[3]; f.a2 = op.p[4];
                    run_recurrence<G>(f, acc, tile, sm_state + op.aux * G, g, j, ch_ok || true ? valid_f4 : 0);
                } break;
                case OP_LP1: {
                    LP1 f; f.r = op.p[0]; f.omr = op.p[1];
                    run_recurrence<G>(f, acc, tile, sm_state + op.aux * G, g, j, valid_f4);
                } break;
                case OP_HP1: {
                    HP1 f; f.r = op.p[0]; f.omr = op.p[1];
                    run_recurrence<G>(f, acc, tile, sm_state + op.aux * G, g, j, valid_f4);
                } break;
                case OP_ENVELOPE: {
                    Env f; f.ga = op.p[0]; f.gr = op.p[1];
                    run_recurrence<G>(f, acc, tile, sm_state + op.aux * G, g, j, valid_f4);
                } break;
                default: break;
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    for (int i = t; i < n_states * G; i += kThreads) {
        int s = i / G, c = c_begin + blockIdx.x * G + (i % G);
        if (c < c_end) reinterpret_cast<float4*>(prog.states[s])[c] = sm_state[s * G + (i % G)];
    }
}

template <int G>
int launch_g(const Program& prog, int c_begin, int c_end, int64_t T, int n_states, cudaStream_t st) {
    const int smem = fused_smem_bytes(prog, G);
    static int configured = -1;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(fused_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    const int n_cta = (c_end - c_begin + G - 1) / G;
    fused_kernel<G><<<n_cta, kThreads, smem, st>>>(prog, c_begin, c_end, (long long)T, n_states);
    return (int)cudaGetLastError();
}

}  // namespace

int fused_smem_bytes(const Program& prog, int G) {
    const int S = kTile / G;
    size_t b = (size_t)kMaxStates * G * 16;
    if (prog.needs_tile) b += (size_t)G * (S + 4) * 4;
    b += (size_t)2 * prog.n_prefetch * kTile * 4;
    b += (size_t)prog.n_vregs * kTile * 4;
    return (int)b;
}

int launch_fused(const Program& prog, int G, int c_begin, int c_end, int64_t T, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int n_states = 0;
    for (int i = 0; i < kMaxStates; i++)
        if (prog.states[i]) n_states = i + 1;
    switch (G) {
        case 1: return launch_g<1>(prog, c_begin, c_end, T, n_states, st);
        case 2: return launch_g<2>(prog, c_begin, c_end, T, n_states, st);
        case 4: return launch_g<4>(prog, c_begin, c_end, T, n_states, st);
        case 8: return launch_g<8>(prog, c_begin, c_end, T, n_states, st);
        case 16: return launch_g<16>(prog, c_begin, c_end, T, n_states, st);
        case 32: return launch_g<32>(prog, c_begin, c_end, T, n_states, st);
    }
    return (int)cudaErrorInvalidValue;
}

}  // namespace dspb
