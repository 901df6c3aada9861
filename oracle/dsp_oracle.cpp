// dsp_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain restatement, on the host CPU, of the arithmetic of simmsb/dsp-stuff's effect-node path
// (SimpleNode::process bodies + the per-block wrapper that feeds them).  It exists so that tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg can CHECK (and time
// next to) the CUDA engine.  Nothing in dsp_stuff_b200/ may include, link, import or call it.
//
// PARITY STATUS: "parity unpinned".  The reference ships no tests, golden vectors or fixtures
//  and cannot be built here (nightly Rust, 519 un-vendored crates, no cargo), so
// the authority of this file is the reference source text cited per function, plus the published
// algorithms of three crates whose sources are not under /root/reference:
//   biquad 0.4.2 (DirectForm1<f32>::run), dasp_envelope 0.11.0 (Detector::next, peak/full-wave),
//   rivulet@b2416e5 (ring capacity rounding -> Reverb delay length; exposed as `ring_granule`).
// Its own checks are the hand-derivable known answers in tests/test_oracle_kat.py and the
// independent scipy cross-checks in tests/test_oracle_scipy.py.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math (rustc never contracts or reassociates
// f32), glibc libm for tanhf/sinf/atanf/expf (what Rust std calls on linux-gnu).
//
// Citations are relative to the reference tree (dsp-stuff/src/...).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr int REF_BLOCK = 128;  // node.rs:257  BUF_SIZE

thread_local std::string g_err;
int fail(int code, const std::string& m) {
    g_err = m;
    return code;
}
enum { OK = 0, E_INVALID = -1, E_NODE = -2, E_PORT = -3, E_GRAPH = -4 };

// ---- Rust f32 helpers, evaluated the way rustc/LLVM evaluates them -------------------------------
inline float rs_clamp(float x, float lo, float hi) {  // f32::clamp: NaN passes through
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}
inline float rs_signum(float x) { return std::isnan(x) ? NAN : std::copysign(1.0f, x); }
inline float powi2(float x) { return x * x; }
inline float powi3(float x) { return (x * x) * x; }           // llvm.powi.f32(x,3)
inline float powi4(float x) { float s = x * x; return s * s; }  // llvm.powi.f32(x,4)
// Iterator::max_by(f32::total_cmp) over |x|: all candidates are >= +0 or a NaN; after abs() a NaN
// has its sign bit cleared and orders above +inf, so "largest bit pattern" is the answer.
inline float max_abs_total(const float* v, int n) {
    uint32_t best = 0;
    for (int i = 0; i < n; i++) {
        float a = std::fabs(v[i]);
        uint32_t b;
        std::memcpy(&b, &a, 4);
        if (b >= best) best = b;
    }
    float r;
    std::memcpy(&r, &best, 4);
    return r;
}

// ---- node base -----------------------------------------------------------------------------------
struct PortIO {
    const float* const* in;  // [n_in][128] averaged port buffers (node.rs:217-222)
    const bool* present;     // node.rs:221
    float* const* out;       // [n_out][128], zero-initialised (node.rs:271-275)
};

struct Node {
    std::string type;
    std::vector<std::string> in_ports, out_ports;
    int channels = 1;
    int sample_rate = 48000;
    virtual ~Node() {}
    virtual void process(int ch, const PortIO& io) = 0;  // SimpleNode::process, node.rs:135-146
    virtual int set_f32(const std::string&, float) { return E_PORT; }
    virtual int set_enum(const std::string&, const std::string&) { return E_PORT; }
    virtual int get_i64(const std::string&, int64_t*) { return E_PORT; }
    virtual void reset() {}
    virtual void init_channels() { reset(); }
    int in_idx(const std::string& n) const {
        for (size_t i = 0; i < in_ports.size(); i++)
            if (in_ports[i] == n) return (int)i;
        return -1;
    }
    int out_idx(const std::string& n) const {
        for (size_t i = 0; i < out_ports.size(); i++)
            if (out_ports[i] == n) return (int)i;
        return -1;
    }
};

// derive helper `<field>_input` — dsp-stuff-derive/src/lib.rs:122-161.  The store of p[0] back into
// the atomic (lib.rs:148) is not modelled: it is unobservable while the control port stays linked.
inline void param_or_mod(const PortIO& io, int port, float lo, float hi, float value, float* out) {
    if (port >= 0 && io.present[port]) {
        const float* c = io.in[port];
        for (int i = 0; i < REF_BLOCK; i++) {
            float y = (c[i] + 1.0f) / 2.0f;
            float z = rs_clamp(y, 0.0f, 1.0f);
            out[i] = lo + (hi - lo) * z;
        }
    } else {
        for (int i = 0; i < REF_BLOCK; i++) out[i] = value;
    }
}

// ---- nodes/gain.rs:25-38 -------------------------------------------------------------------------
struct Gain : Node {
    float level = 1.0f;  // gain.rs:21
    Gain() { type = "gain"; in_ports = {"in", "level"}; out_ports = {"out"}; }
    int set_f32(const std::string& f, float v) override {
        if (f == "level") { level = v; return OK; }
        return E_PORT;
    }
    void process(int, const PortIO& io) override {
        float lv[REF_BLOCK];
        param_or_mod(io, 1, 0.0f, 10.0f, level, lv);
        for (int i = 0; i < REF_BLOCK; i++) io.out[0][i] = io.in[0][i] * lv[i];
    }
};

// ---- nodes/distort.rs ----------------------------------------------------------------------------
inline float clip(float s) { return s < -1.0f ? -1.0f : (s > 1.0f ? 1.0f : s); }  // distort.rs:53-61
struct Distort : Node {
    enum Mode { HardClip, SoftClip, Tanh, RecipSoftClip, Fuzz, Sin, Atan, Square, Chebyshev4 };
    float level = 0.0f;    // distort.rs:46-47: no default => 0.0
    int mode = SoftClip;   // distort.rs:49
    Distort() { type = "distort"; in_ports = {"in", "level"}; out_ports = {"out"}; }
    int set_f32(const std::string& f, float v) override {
        if (f == "level") { level = v; return OK; }
        return E_PORT;
    }
    int set_enum(const std::string& f, const std::string& v) override {
        static const char* names[] = {"HardClip", "SoftClip", "Tanh", "RecipSoftClip", "Fuzz",
                                      "Sin", "Atan", "Square", "Chebyshev4"};
        if (f != "mode") return E_PORT;
        for (int i = 0; i < 9; i++)
            if (v == names[i]) { mode = i; return OK; }
        return E_PORT;
    }
    static float shape(int mode, float x, float l) {
        if (l < 0.001f) return x;  // every non-fuzz shaper: distort.rs:64,72,97,105,113,121,129,137
        switch (mode) {
            case HardClip: return clip(x * l) / l;  // 63-69
            case SoftClip: {                        // 71-86
                float s = x * l;
                if (s > 1.0f) s = 2.0f / 3.0f;
                else if (s >= -1.0f && s <= 1.0f) s = s - (powi3(s) / 3.0f);
                else s = -2.0f / 3.0f;
                return clip(s) / l;
            }
            case Tanh: return std::tanh(x * l);                                                  // 104-110
            case RecipSoftClip: return rs_signum(x) * (1.0f - 1.0f / (std::fabs(x) * l + 1.0f));  // 96-102
            case Sin: return std::sin(x * l);                                                    // 112-118
            case Atan: return std::atan(x * l);                                                  // 120-126
            case Square: return powi2(x * l) * rs_signum(x * l);                                 // 128-134
            case Chebyshev4: {                                                                   // 136-144
                float v = x * l;
                return 8.0f * powi4(v) - 8.0f * powi2(v) + 1.0f;
            }
        }
        return x;
    }
    static void fuzz(const float* in, float* out, const float* level) {  // distort.rs:146-172
        float mx = max_abs_total(in, REF_BLOCK);
        float z[REF_BLOCK], y[REF_BLOCK];
        for (int i = 0; i < REF_BLOCK; i++) {
            float q = clip(in[i] * level[i]) / mx;
            z[i] = std::copysign(1.0f - std::exp(std::copysign(q, -1.0f)), -1.0f);
        }
        float mz = max_abs_total(z, REF_BLOCK);
        for (int i = 0; i < REF_BLOCK; i++) y[i] = clip(z[i] * mx) / mz;
        float my = max_abs_total(y, REF_BLOCK);
        for (int i = 0; i < REF_BLOCK; i++) out[i] = y[i] * mx / my;
    }
    void process(int, const PortIO& io) override {  // distort.rs:174-196
        float lv[REF_BLOCK];
        param_or_mod(io, 1, 0.0f, 30.0f, level, lv);
        if (mode == Fuzz) { fuzz(io.in[0], io.out[0], lv); return; }
        for (int i = 0; i < REF_BLOCK; i++) io.out[0][i] = shape(mode, io.in[0][i], lv[i]);
    }
};

// ---- nodes/overdrive.rs:31-73 --------------------------------------------------------------------
struct Overdrive : Node {
    float boost = 0.0f, drive = 0.0f, level = 0.0f;
    Overdrive() { type = "overdrive"; in_ports = {"in", "boost", "drive", "level"}; out_ports = {"out"}; }
    int set_f32(const std::string& f, float v) override {
        if (f == "boost") boost = v; else if (f == "drive") drive = v; else if (f == "level") level = v;
        else return E_PORT;
        return OK;
    }
    void process(int, const PortIO& io) override {
        float b[REF_BLOCK], d[REF_BLOCK], l[REF_BLOCK];
        param_or_mod(io, 1, 0.0f, 30.0f, boost, b);
        param_or_mod(io, 3, 0.0f, 1.0f, level, l);
        param_or_mod(io, 2, 0.0f, 1.0f, drive, d);
        const float FRAC_PI_4 = 0.785398163397448309615660845819875721f;
        const float FRAC_2_PI = 0.636619772367581343075535053490057448f;
        for (int i = 0; i < REF_BLOCK; i++) {
            float x = io.in[0][i];
            if (l[i] < 0.001f) { io.out[0][i] = x; continue; }
            float a = x * b[i];
            float bb = FRAC_PI_4 * a;
            float c = std::atan(bb);
            float dd = FRAC_2_PI * c;
            float mix = d[i] * dd + (1.0f - d[i]) * x;
            io.out[0][i] = mix * l[i];
        }
    }
};

// ---- nodes/chebyshev.rs:28-63 --------------------------------------------------------------------
struct Chebyshev : Node {
    float level_pos = 0.0f, level_neg = 0.0f;
    Chebyshev() { type = "chebyshev"; in_ports = {"in"}; out_ports = {"out"}; }
    int set_f32(const std::string& f, float v) override {
        if (f == "level_pos") level_pos = v; else if (f == "level_neg") level_neg = v; else return E_PORT;
        return OK;
    }
    void process(int, const PortIO& io) override {
        for (int i = 0; i < REF_BLOCK; i++) {
            float x = io.in[0][i], r;
            if (x >= 0.0f) r = level_pos < 0.001f ? x : std::tanh(x * level_pos) / std::tanh(level_pos);
            else r = level_neg < 0.001f ? x : std::tanh(x * level_neg) / std::tanh(level_neg);
            io.out[0][i] = r;
        }
    }
};

// ---- nodes/biquad.rs + crate biquad 0.4.2 DirectForm1<f32> (published formula) --------------------
struct BiQuad : Node {
    float a0 = 1.0f, a1 = -0.24f, a2 = 0.0f, b0 = 0.758f, b1 = 0.0f, b2 = 0.0f;  // biquad.rs:25-41
    struct Coef { float a1, a2, b0, b1, b2; } c{-0.24f, 0.0f, 0.758f, 0.0f, 0.0f};  // biquad.rs:48-55
    struct St { float x1 = 0, x2 = 0, y1 = 0, y2 = 0; };
    std::vector<St> st;
    BiQuad() { type = "biquad"; in_ports = {"in"}; out_ports = {"out"}; }
    void reset() override { st.assign(channels, St{}); }
    int set_f32(const std::string& f, float v) override {
        if (f == "a0") a0 = v; else if (f == "a1") a1 = v; else if (f == "a2") a2 = v;
        else if (f == "b0") b0 = v; else if (f == "b1") b1 = v; else if (f == "b2") b2 = v;
        else return E_PORT;
        // regenerate_filter, biquad.rs:62-76: coefficients / a0 in f32, state reset.
        c = Coef{a1 / a0, a2 / a0, b0 / a0, b1 / a0, b2 / a0};
        reset();
        return OK;
    }
    void process(int ch, const PortIO& io) override {  // biquad.rs:79-88
        St& s = st[ch];
        for (int i = 0; i < REF_BLOCK; i++) {
            float x = io.in[0][i];
            // DirectForm1::run: b0*x + b1*x1 + b2*x2 - a1*y1 - a2*y2, left to right, no FMA
            float out = c.b0 * x + c.b1 * s.x1 + c.b2 * s.x2 - c.a1 * s.y1 - c.a2 * s.y2;
            s.x2 = s.x1; s.x1 = x; s.y2 = s.y1; s.y1 = out;
            io.out[0][i] = out;
        }
    }
};

// ---- nodes/low_pass.rs:26-42, nodes/high_pass.rs:26-42 -------------------------------------------
struct OnePole : Node {
    bool high;
    float ratio = 0.5f;
    std::vector<float> z;
    explicit OnePole(bool hp) : high(hp) { type = hp ? "high_pass" : "low_pass"; in_ports = {"in"}; out_ports = {"out"}; }
    void reset() override { z.assign(channels, 0.0f); }
    int set_f32(const std::string& f, float v) override {
        if (f == "ratio") { ratio = v; return OK; }
        return E_PORT;
    }
    void process(int ch, const PortIO& io) override {
        float zz = z[ch];
        for (int i = 0; i < REF_BLOCK; i++) {
            float x = io.in[0][i];
            if (!high) { float o = x * (1.0f - ratio) + ratio * zz; zz = o; io.out[0][i] = o; }
            else { zz = x * (1.0f - ratio) + ratio * zz; io.out[0][i] = x - zz; }
        }
        z[ch] = zz;
    }
};

// ---- nodes/reverb.rs (the "delay": feedback comb through a zero-prefilled ring) -------------------
struct Reverb : Node {
    float seconds = 0.5f, decay = 0.5f;  // reverb.rs:26-39
    int granule = 1024;
    int64_t D = 0;
    std::vector<std::vector<float>> ring;
    std::vector<int64_t> pos;
    Reverb() { type = "reverb"; in_ports = {"in"}; out_ports = {"out"}; }
    static int64_t round_up(int64_t n, int64_t g) { return g <= 1 ? n : (n + g - 1) / g * g; }
    void init_channels() override {
        // make_buffer(), reverb.rs:44-52: circular_buffer(128), zero-fill view().len()
        D = round_up(128, granule);
        reset();
    }
    void refresh() {  // refresh_seconds, reverb.rs:55-71
        float prod = seconds * (float)sample_rate;
        int64_t num = 0;  // `as usize`: NaN/negative -> 0, saturating, truncating
        if (prod > 0.0f) num = prod >= 9.0e18f ? INT64_MAX / 2 : (int64_t)prod;
        num = std::max<int64_t>(num, 128);
        D = round_up(num, granule);
        reset();
    }
    void reset() override {
        ring.assign(channels, std::vector<float>((size_t)D, 0.0f));
        pos.assign(channels, 0);
    }
    int set_f32(const std::string& f, float v) override {
        if (f == "seconds") { seconds = v; refresh(); return OK; }
        if (f == "decay") { decay = v; refresh(); return OK; }  // any slider change runs after_settings_change (lib.rs:560-568)
        return E_PORT;
    }
    int get_i64(const std::string& k, int64_t* o) override {
        if (k == "delay_samples") { *o = D; return OK; }
        return E_PORT;
    }
    void process(int ch, const PortIO& io) override {  // reverb.rs:74-111
        std::vector<float>& r = ring[ch];
        int64_t p = pos[ch];
        for (int i = 0; i < REF_BLOCK; i++) {
            float o = io.in[0][i] + r[(size_t)p] * decay;  // a + b * decay, reverb.rs:90
            io.out[0][i] = o;
            r[(size_t)p] = o;                              // reverb.rs:99-103
            if (++p == D) p = 0;
        }
        pos[ch] = p;
    }
};

// ---- nodes/fir.rs:179-225 ------------------------------------------------------------------------
struct Fir : Node {
    enum Mode { Average, Balanced };
    int mode = Balanced;              // fir.rs:55
    std::vector<double> taps{1.0};    // fir.rs:61, stored REVERSED (fir.rs:153-171)
    struct Hist { std::vector<double> buf; size_t start = 0, len = 0; };
    std::vector<Hist> hist;
    Fir() { type = "fir"; in_ports = {"in"}; out_ports = {"out"}; }
    void reset() override {
        hist.assign(channels, Hist{});
        for (auto& h : hist) h.buf.assign(2 * taps.size() + 2, 0.0);
    }
    int set_enum(const std::string& f, const std::string& v) override {
        if (f != "mode") return E_PORT;
        if (v == "Average") mode = Average; else if (v == "Balanced") mode = Balanced; else return E_PORT;
        return OK;
    }
    int get_i64(const std::string& k, int64_t* o) override {
        if (k == "n_taps") { *o = (int64_t)taps.size(); return OK; }
        return E_PORT;
    }
    void process(int ch, const PortIO& io) override {
        Hist& h = hist[ch];
        const size_t N = taps.size();
        const float divisor = mode == Average ? 1.0f / (float)N : 1.0f;  // fir.rs:187-190
        const double* t = taps.data();
        for (int i = 0; i < REF_BLOCK; i++) {
            // state.push_back(x as f64); if len > N pop_front  (fir.rs:193-197), kept contiguous
            if (h.start + h.len == h.buf.size()) {
                std::memmove(h.buf.data(), h.buf.data() + h.start, h.len * sizeof(double));
                h.start = 0;
            }
            h.buf[h.start + h.len] = (double)io.in[0][i];
            h.len++;
            if (h.len > N) { h.start++; h.len--; }
            // zip(state, taps).map(x*c).sum::<f64>() as f32: sequential f64, oldest sample first.
            // (The reference splits the sum at the VecDeque wrap point and adds the two f32-cast
            //  halves, fir.rs:201-216; the split position depends on std's VecDeque internals and
            //  moves the result by <= 1 f32 ulp.  Single sum here; documented in DESIGN.md.)
            const double* s = h.buf.data() + h.start;
            double acc = 0.0;
            for (size_t k = 0; k < h.len; k++) acc += s[k] * t[k];
            float val = (float)acc;
            io.out[0][i] = val * divisor;  // fir.rs:222
        }
    }
};

// ---- nodes/add.rs, mix.rs, mux.rs, demux.rs ------------------------------------------------------
struct Add : Node {
    Add() { type = "add"; in_ports = {"a", "b"}; out_ports = {"out"}; }
    void process(int, const PortIO& io) override {
        for (int i = 0; i < REF_BLOCK; i++) io.out[0][i] = io.in[0][i] + io.in[1][i];
    }
};
struct Mix : Node {
    float ratio = 0.5f;
    Mix() { type = "mix"; in_ports = {"a", "b", "ratio"}; out_ports = {"out"}; }
    int set_f32(const std::string& f, float v) override {
        if (f == "ratio") { ratio = v; return OK; }
        return E_PORT;
    }
    void process(int, const PortIO& io) override {  // mix.rs:33-47
        float r[REF_BLOCK];
        param_or_mod(io, 2, 0.0f, 1.0f, ratio, r);
        for (int i = 0; i < REF_BLOCK; i++) io.out[0][i] = (io.in[1][i] * r[i]) + (io.in[0][i] * (1.0f - r[i]));
    }
};
struct Mux : Node {
    int port = 0;
    Mux() { type = "mux"; in_ports = {"a", "b"}; out_ports = {"out"}; }
    int set_enum(const std::string& f, const std::string& v) override {
        if (f != "in_port") return E_PORT;
        if (v == "A") port = 0; else if (v == "B") port = 1; else return E_PORT;
        return OK;
    }
    void process(int, const PortIO& io) override { std::memcpy(io.out[0], io.in[port], REF_BLOCK * 4); }
};
struct Demux : Node {
    int port = 0;
    Demux() { type = "demux"; in_ports = {"in"}; out_ports = {"a", "b"}; }
    int set_enum(const std::string& f, const std::string& v) override {
        if (f != "out_port") return E_PORT;
        if (v == "A") port = 0; else if (v == "B") port = 1; else return E_PORT;
        return OK;
    }
    void process(int, const PortIO& io) override { std::memcpy(io.out[port], io.in[0], REF_BLOCK * 4); }
};

// ---- nodes/envelope.rs:34-52 + dasp_envelope 0.11.0 Detector<f32, Peak<FullWave>> ------------------
struct Envelope : Node {
    float attack = 0.0f, release = 0.0f;
    std::vector<float> last;
    Envelope() { type = "envelope"; in_ports = {"in"}; out_ports = {"out"}; }
    void reset() override { last.assign(channels, 0.0f); }
    int set_f32(const std::string& f, float v) override {
        if (f == "attack") attack = v; else if (f == "release") release = v; else return E_PORT;
        return OK;
    }
    static float calc_gain(float frames) { return frames == 0.0f ? 0.0f : std::exp(-1.0f / frames); }
    void process(int ch, const PortIO& io) override {
        float ga = calc_gain(attack), gr = calc_gain(release);  // set_*_frames every block, envelope.rs:45-46
        float prev = last[ch];
        for (int i = 0; i < REF_BLOCK; i++) {
            float d = std::fabs(io.in[0][i]);               // Peak<FullWave>
            float g = prev < d ? ga : gr;                   // Detector::next
            prev = d + (prev - d) * g;
            io.out[0][i] = prev;
        }
        last[ch] = prev;
    }
};

// ---- EXTENSION (no reference node; BASELINE north_star names a noise gate) ----------
// Hard gate keyed by the Envelope node's detector: out = env(in) >= threshold ? in : 0.
struct Gate : Node {
    float threshold = 0.0f, attack = 0.0f, release = 0.0f;
    std::vector<float> last;
    Gate() { type = "gate"; in_ports = {"in"}; out_ports = {"out"}; }
    void reset() override { last.assign(channels, 0.0f); }
    int set_f32(const std::string& f, float v) override {
        if (f == "threshold") threshold = v; else if (f == "attack") attack = v; else if (f == "release") release = v; else return E_PORT;
        return OK;
    }
    void process(int ch, const PortIO& io) override {
        float ga = Envelope::calc_gain(attack), gr = Envelope::calc_gain(release);
        float prev = last[ch];
        for (int i = 0; i < REF_BLOCK; i++) {
            float d = std::fabs(io.in[0][i]);
            float g = prev < d ? ga : gr;
            prev = d + (prev - d) * g;
            io.out[0][i] = prev >= threshold ? io.in[0][i] : 0.0f;
        }
        last[ch] = prev;
    }
};

// ---- nodes/signal_gen.rs:55-130 ------------------------------------------------------------------
struct SignalGen : Node {
    enum Mode { Sine, Triangle, Square, Constant };
    float amplitude = 0.5f, frequency = 100.0f;
    int mode = Sine;
    std::vector<float> clock;
    SignalGen() { type = "signal_gen"; in_ports = {"amplitude", "frequency"}; out_ports = {"out"}; }
    void reset() override { clock.assign(channels, 0.0f); }
    int set_f32(const std::string& f, float v) override {
        if (f == "amplitude") amplitude = v; else if (f == "frequency") frequency = v; else return E_PORT;
        return OK;
    }
    int set_enum(const std::string& f, const std::string& v) override {
        static const char* names[] = {"Sine", "Triangle", "Square", "Constant"};
        if (f != "mode") return E_PORT;
        for (int i = 0; i < 4; i++)
            if (v == names[i]) { mode = i; return OK; }
        return E_PORT;
    }
    void process(int ch, const PortIO& io) override {
        float amp[REF_BLOCK], frq[REF_BLOCK];
        param_or_mod(io, 0, -1.0f, 1.0f, amplitude, amp);
        param_or_mod(io, 1, 0.1f, 20000.0f, frequency, frq);
        float* out = io.out[0];
        if (mode == Constant) { std::memcpy(out, amp, sizeof(amp)); return; }  // do_const: clock untouched
        const float sr = (float)sample_rate;
        const float TAU = 6.28318530717958647692528676655900577f;
        float clk = clock[ch], total = 0.0f;
        for (int i = 0; i < REF_BLOCK; i++) {
            float step = frq[i] / sr;
            total += step;
            if (mode == Sine) out[i] = std::sin((clk + total) * TAU) * amp[i];
            else if (mode == Triangle) out[i] = (2.0f * std::fmod(clk + total, 1.0f) - 1.0f) * amp[i];
            else out[i] = (total > 0.5f ? 1.0f : -1.0f) * amp[i];  // do_square ignores clock (sic)
        }
        clock[ch] = std::fmod(clk + total, 1.0f);
    }
};

// graph terminals: nodes/input.rs:213-241 (raw copy to every link), nodes/output.rs:215-250
struct Terminal : Node {
    bool is_input;
    explicit Terminal(bool in) : is_input(in) {
        type = in ? "input" : "output";
        if (in) out_ports = {"out"}; else in_ports = {"in"};
    }
    void process(int, const PortIO&) override {}
};

std::unique_ptr<Node> make_node(const std::string& t) {  // nodes/mod.rs:92-123 (RESTORE keys)
    if (t == "gain") return std::make_unique<Gain>();
    if (t == "distort") return std::make_unique<Distort>();
    if (t == "overdrive") return std::make_unique<Overdrive>();
    if (t == "chebyshev") return std::make_unique<Chebyshev>();
    if (t == "biquad") return std::make_unique<BiQuad>();
    if (t == "low_pass") return std::make_unique<OnePole>(false);
    if (t == "high_pass") return std::make_unique<OnePole>(true);
    if (t == "reverb") return std::make_unique<Reverb>();
    if (t == "fir") return std::make_unique<Fir>();
    if (t == "add") return std::make_unique<Add>();
    if (t == "mix") return std::make_unique<Mix>();
    if (t == "mux") return std::make_unique<Mux>();
    if (t == "demux") return std::make_unique<Demux>();
    if (t == "envelope") return std::make_unique<Envelope>();
    if (t == "signal_gen") return std::make_unique<SignalGen>();
    if (t == "gate") return std::make_unique<Gate>();
    if (t == "input") return std::make_unique<Terminal>(true);
    if (t == "output") return std::make_unique<Terminal>(false);
    return nullptr;
}

struct Link { int src, sport, dst, dport; };

struct Engine {
    int channels = 1, sample_rate = 48000, granule = 1024, threads = 0;
    std::vector<std::unique_ptr<Node>> nodes;
    std::vector<int64_t> ids;
    std::vector<Link> links;
    std::vector<int> order;           // topological order of running nodes
    std::vector<int> in_terms, out_terms;
    // per node, per input port: link indices (creation order)
    std::vector<std::vector<std::vector<int>>> in_links;
    std::vector<std::vector<std::vector<int>>> out_links;
    bool compiled = false;
    int find(int64_t id) const {
        for (size_t i = 0; i < ids.size(); i++)
            if (ids[i] == id) return (int)i;
        return -1;
    }
};

// One channel, one 128-sample reference block through the whole graph: the block-synchronous
// equivalent of the per-node task loops (runtime.rs:646-732) on a DAG.
struct Scratch {
    std::vector<float> link_buf;   // [n_links][128]  one ring per link (runtime.rs:566-578)
    std::vector<float> in_buf;     // [max_in][128]
    std::vector<float> out_buf;    // [max_out][128]
};

void run_block(Engine& e, Scratch& s, int ch, const float* const* ins, float* const* outs, int64_t n,
               int64_t off) {
    const float* inp[8];
    float* outp[8];
    bool present[8];
    for (int ni : e.order) {
        Node& nd = *e.nodes[ni];
        if (nd.type == "input") {  // nodes/input.rs:221-229: raw samples copied into every out link
            int t = (int)(std::find(e.in_terms.begin(), e.in_terms.end(), ni) - e.in_terms.begin());
            const float* src = ins[t] + (size_t)ch * n + off;
            for (int l : e.out_links[ni][0]) std::memcpy(&s.link_buf[(size_t)l * REF_BLOCK], src, REF_BLOCK * 4);
            continue;
        }
        const int n_in = (int)nd.in_ports.size(), n_out = (int)nd.out_ports.size();
        for (int p = 0; p < n_in; p++) {
            // collect_and_average, node.rs:162-194
            float* buf = &s.in_buf[(size_t)p * REF_BLOCK];
            for (int i = 0; i < REF_BLOCK; i++) buf[i] = 0.0f;  // input_buf.resize(.., 0.0), node.rs:288
            float num_frames = 0.0001f;
            bool r = false;
            for (int l : e.in_links[ni][p]) {
                r = true;
                num_frames += 1.0f;
                const float* v = &s.link_buf[(size_t)l * REF_BLOCK];
                for (int i = 0; i < REF_BLOCK; i++) buf[i] += v[i];
            }
            for (int i = 0; i < REF_BLOCK; i++) buf[i] /= num_frames;
            inp[p] = buf;
            present[p] = r;
        }
        if (nd.type == "output") {  // nodes/output.rs:223-233: averaged block goes to the device sink
            int t = (int)(std::find(e.out_terms.begin(), e.out_terms.end(), ni) - e.out_terms.begin());
            std::memcpy(outs[t] + (size_t)ch * n + off, inp[0], REF_BLOCK * 4);
            continue;
        }
        for (int q = 0; q < n_out; q++) {
            outp[q] = &s.out_buf[(size_t)q * REF_BLOCK];
            for (int i = 0; i < REF_BLOCK; i++) outp[q][i] = 0.0f;  // node.rs:271-275
        }
        PortIO io{inp, present, outp};
        nd.process(ch, io);  // node.rs:317
        for (int q = 0; q < n_out; q++)  // node.rs:321-325: copy to every link of the port
            for (int l : e.out_links[ni][q]) std::memcpy(&s.link_buf[(size_t)l * REF_BLOCK], outp[q], REF_BLOCK * 4);
    }
}

}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

int orc_engine_create(int channels, int sample_rate, int ring_granule, int threads, void** out) {
    if (!out || channels <= 0) return fail(E_INVALID, "bad arguments");
    auto* e = new Engine();
    e->channels = channels;
    e->sample_rate = sample_rate > 0 ? sample_rate : 48000;
    e->granule = ring_granule > 0 ? ring_granule : 1024;
    e->threads = threads;
    *out = e;
    return OK;
}
void orc_engine_destroy(void* h) { delete (Engine*)h; }

int orc_node_add(void* h, const char* cfg_name, int64_t id) {
    Engine& e = *(Engine*)h;
    if (e.find(id) >= 0) return fail(E_INVALID, "duplicate node id");
    auto n = make_node(cfg_name ? cfg_name : "");
    if (!n) return fail(E_NODE, std::string("unknown typename ") + (cfg_name ? cfg_name : "(null)"));
    n->channels = e.channels;
    n->sample_rate = e.sample_rate;
    if (auto* r = dynamic_cast<Reverb*>(n.get())) r->granule = e.granule;
    n->init_channels();
    e.nodes.push_back(std::move(n));
    e.ids.push_back(id);
    e.compiled = false;
    return OK;
}
int orc_node_set_f32(void* h, int64_t id, const char* field, float v) {
    Engine& e = *(Engine*)h;
    int i = e.find(id);
    if (i < 0) return fail(E_NODE, "unknown node id");
    int r = e.nodes[i]->set_f32(field, v);
    return r == OK ? OK : fail(r, std::string("unknown f32 field ") + field);
}
int orc_node_set_enum(void* h, int64_t id, const char* field, const char* variant) {
    Engine& e = *(Engine*)h;
    int i = e.find(id);
    if (i < 0) return fail(E_NODE, "unknown node id");
    int r = e.nodes[i]->set_enum(field, variant);
    return r == OK ? OK : fail(r, std::string("unknown enum field/variant ") + field + "=" + variant);
}
int orc_node_set_taps(void* h, int64_t id, const double* taps, int64_t n) {
    Engine& e = *(Engine*)h;
    int i = e.find(id);
    if (i < 0) return fail(E_NODE, "unknown node id");
    auto* f = dynamic_cast<Fir*>(e.nodes[i].get());
    if (!f || n <= 0) return fail(E_INVALID, "not a fir node / empty taps");
    f->taps.assign(taps, taps + n);
    f->reset();
    return OK;
}
int orc_node_get_i64(void* h, int64_t id, const char* key, int64_t* out) {
    Engine& e = *(Engine*)h;
    int i = e.find(id);
    if (i < 0) return fail(E_NODE, "unknown node id");
    std::string k = key;
    if (k == "n_inputs") { *out = (int64_t)e.nodes[i]->in_ports.size(); return OK; }
    if (k == "n_outputs") { *out = (int64_t)e.nodes[i]->out_ports.size(); return OK; }
    int r = e.nodes[i]->get_i64(k, out);
    return r == OK ? OK : fail(r, "unknown key " + k);
}
int orc_node_port_index(void* h, int64_t id, const char* port, int is_output, int32_t* out) {
    Engine& e = *(Engine*)h;
    int i = e.find(id);
    if (i < 0) return fail(E_NODE, "unknown node id");
    int p = is_output ? e.nodes[i]->out_idx(port) : e.nodes[i]->in_idx(port);
    if (p < 0) return fail(E_PORT, std::string("unknown port ") + port);
    *out = p;
    return OK;
}
int orc_link(void* h, int64_t src, const char* oport, int64_t dst, const char* iport) {
    Engine& e = *(Engine*)h;
    int s = e.find(src), d = e.find(dst);
    if (s < 0 || d < 0) return fail(E_NODE, "unknown node id in link");
    int sp = e.nodes[s]->out_idx(oport), dp = e.nodes[d]->in_idx(iport);
    if (sp < 0 || dp < 0) return fail(E_PORT, std::string("unknown port in link ") + oport + " -> " + iport);
    e.links.push_back(Link{s, sp, d, dp});
    e.compiled = false;
    return OK;
}
int orc_compile(void* h) {
    Engine& e = *(Engine*)h;
    const int N = (int)e.nodes.size();
    e.in_links.assign(N, {});
    e.out_links.assign(N, {});
    for (int i = 0; i < N; i++) {
        e.in_links[i].assign(e.nodes[i]->in_ports.size(), {});
        e.out_links[i].assign(e.nodes[i]->out_ports.size(), {});
    }
    std::vector<int> indeg(N, 0), nlinks(N, 0);
    for (size_t l = 0; l < e.links.size(); l++) {
        const Link& k = e.links[l];
        e.in_links[k.dst][k.dport].push_back((int)l);
        e.out_links[k.src][k.sport].push_back((int)l);
        indeg[k.dst]++;
        nlinks[k.src]++;
        nlinks[k.dst]++;
    }
    e.in_terms.clear();
    e.out_terms.clear();
    for (int i = 0; i < N; i++) {
        if (e.nodes[i]->type == "input") e.in_terms.push_back(i);
        if (e.nodes[i]->type == "output") e.out_terms.push_back(i);
    }
    // Kahn; nodes without any link do not run (runtime.rs:661-668).
    e.order.clear();
    std::vector<int> ready;
    for (int i = 0; i < N; i++)
        if (indeg[i] == 0) ready.push_back(i);
    size_t head = 0, seen = 0;
    while (head < ready.size()) {
        int n = ready[head++];
        seen++;
        if (nlinks[n] > 0) e.order.push_back(n);
        for (auto& port : e.out_links[n])
            for (int l : port)
                if (--indeg[e.links[l].dst] == 0) ready.push_back(e.links[l].dst);
    }
    if ((int)seen != N) return fail(E_GRAPH, "graph has a cycle (the reference would deadlock, runtime.rs:568)");
    e.compiled = true;
    return OK;
}
int orc_reset_state(void* h) {
    Engine& e = *(Engine*)h;
    for (auto& n : e.nodes) n->reset();
    return OK;
}

// inputs[t]: [C x n] for the t-th `input` terminal; outputs[t]: [C x n] for the t-th `output`.
int orc_process(void* h, const float* const* inputs, float* const* outputs, int64_t n) {
    Engine& e = *(Engine*)h;
    if (!e.compiled) return fail(E_GRAPH, "graph not compiled");
    if (n <= 0 || n % REF_BLOCK) return fail(E_INVALID, "n_samples must be a positive multiple of 128");
    for (size_t t = 0; t < e.out_terms.size(); t++) std::memset(outputs[t], 0, (size_t)e.channels * n * 4);
    int nt = e.threads > 0 ? e.threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, e.channels));
    std::atomic<int> next{0};
    auto worker = [&]() {
        Scratch s;
        s.link_buf.assign(e.links.size() * REF_BLOCK + 1, 0.0f);
        s.in_buf.assign(8 * REF_BLOCK, 0.0f);
        s.out_buf.assign(8 * REF_BLOCK, 0.0f);
        for (;;) {
            int ch = next.fetch_add(1);
            if (ch >= e.channels) break;
            for (int64_t off = 0; off < n; off += REF_BLOCK) run_block(e, s, ch, inputs, outputs, n, off);
        }
    };
    if (nt == 1) worker();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < nt; i++) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    return OK;
}

// One SimpleNode::process call per 128-sample block on pre-averaged port buffers (node.rs:135-146).
int orc_node_process(void* h, int64_t id, const float* const* port_in, const uint8_t* present_in,
                     float* const* port_out, int64_t n) {
    Engine& e = *(Engine*)h;
    int ni = e.find(id);
    if (ni < 0) return fail(E_NODE, "unknown node id");
    if (n <= 0 || n % REF_BLOCK) return fail(E_INVALID, "n_samples must be a positive multiple of 128");
    Node& nd = *e.nodes[ni];
    const int n_in = (int)nd.in_ports.size(), n_out = (int)nd.out_ports.size();
    std::vector<float> zeros(REF_BLOCK, 0.0f), ob((size_t)n_out * REF_BLOCK);
    for (int ch = 0; ch < e.channels; ch++)
        for (int64_t off = 0; off < n; off += REF_BLOCK) {
            const float* inp[8];
            bool present[8];
            float* outp[8];
            for (int p = 0; p < n_in; p++) {
                bool has = port_in[p] != nullptr && (!present_in || present_in[p]);
                inp[p] = has ? port_in[p] + (size_t)ch * n + off : zeros.data();
                present[p] = has;
            }
            for (int q = 0; q < n_out; q++) {
                outp[q] = &ob[(size_t)q * REF_BLOCK];
                for (int i = 0; i < REF_BLOCK; i++) outp[q][i] = 0.0f;
            }
            PortIO io{inp, present, outp};
            nd.process(ch, io);
            for (int q = 0; q < n_out; q++)
                if (port_out[q]) std::memcpy(port_out[q] + (size_t)ch * n + off, outp[q], REF_BLOCK * 4);
        }
    return OK;
}


// ---- playback-side sample-rate converter -----------------------------------------------------
// devices.rs:550-556 builds `Converter::from_hz_to_hz(CountingSignal::new(), Sinc::new(Fixed::from([0.0; 16])), 48_000.0,
// target)` and devices.rs:443-500 (`do_write_2`) pulls one frame per stereo output frame from it and writes it to both slots.
// The arithmetic lives in two crates that are NOT under /root/reference (Cargo.lock:1229-1236, 1273-1284): dasp_signal
// 0.11.0 `interpolate::Converter` and dasp_interpolate 0.11.0 `sinc::Sinc`; restated here from their published sources --
// PARITY UNPINNED, like biquad / dasp_envelope / rivulet:
//   Converter::next:   while interpolation_value >= 1.0 { interpolator.next_source_frame(source.next()); value -= 1.0 }
//                      out = interpolator.interpolate(value); value += source_hz / target_hz
//   Sinc::next_source_frame: frames.push(frame) (a 16-frame ring, oldest dropped); if idx < depth { idx += 1 }   (depth = 8)
//   Sinc::interpolate(x):    nl = idx, nr = idx + 1; max_depth = nl + depth >= 16 ? 16 - depth : nr - depth < 0 ? nr : depth;
//                            v = 0; for n in 0..max_depth { a = PI (x + n);       v += f32(sinc(a) hann(a) f64(frames[nl - n]))
//                                                           a = PI (1 - x + n);   v += f32(sinc(a) hann(a) f64(frames[nr + n])) }
//                            sinc(a) = a == 0 ? 1 : sin(a) / a, hann(a) = 0.5 + 0.5 cos(a / depth); ring_buffer::Fixed indexes
//                            modulo its length, so frames[16] (n = 7 on the right in the steady state) is the OLDEST frame.
//   CountingSignal::next (devices.rs:377-392): inner[index++] while inside the prepared buffer, 0.0 (index unchanged) beyond.
struct SincConverter {
    double value = 0.0, ratio = 1.0;
    float ring[16];
    int first = 0, idx = 0;
    SincConverter() { for (auto& r : ring) r = 0.0f; }
    float frame(int i) const { return ring[(first + i) % 16]; }
    void push(float f) { ring[first] = f; first = (first + 1) % 16; if (idx < 8) idx++; }
    float interpolate(double x) const {
        const int depth = 8, nl = idx, nr = idx + 1;
        const int rightmost = nl + depth, leftmost = nr - depth;
        const int max_depth = rightmost >= 16 ? 16 - depth : (leftmost < 0 ? depth + leftmost : depth);
        const double phil = x, phir = 1.0 - x;
        float v = 0.0f;
        for (int n = 0; n < max_depth; n++) {
            double a = M_PI * (phil + (double)n);
            double first_ = a == 0.0 ? 1.0 : std::sin(a) / a;
            double second = 0.5 + 0.5 * std::cos(a / (double)depth);
            v = v + (float)(first_ * second * (double)frame(nl - n));
            a = M_PI * (phir + (double)n);
            first_ = a == 0.0 ? 1.0 : std::sin(a) / a;
            second = 0.5 + 0.5 * std::cos(a / (double)depth);
            v = v + (float)(first_ * second * (double)frame(nr + n));
        }
        return v;
    }
};
struct Resampler {
    int channels;
    std::vector<SincConverter> conv;
};

int orc_resampler_create(int channels, double source_hz, double target_hz, void** out) {
    if (channels <= 0 || !(source_hz > 0.0) || !(target_hz > 0.0) || !out) return fail(E_INVALID, "bad resampler arguments");
    auto* r = new Resampler{channels, std::vector<SincConverter>((size_t)channels)};
    for (auto& c : r->conv) c.ratio = source_hz / target_hz;  // Converter::from_hz_to_hz -> scale_playback_hz(source / target)
    *out = r;
    return OK;
}
void orc_resampler_destroy(void* h) { delete (Resampler*)h; }

// mono [C x n_in] at 48 kHz -> interleaved stereo [C x n_out x 2] at the target rate; *consumed = CountingSignal::index after
// the call (what do_write_2 releases from the link ring), identical for every channel.
int orc_resampler_process(void* h, const float* mono, int64_t n_in, float* interleaved, int64_t n_out, int64_t* consumed) {
    Resampler& r = *(Resampler*)h;
    int64_t used = 0;
    for (int ch = 0; ch < r.channels; ch++) {
        SincConverter& c = r.conv[(size_t)ch];
        const float* src = mono + (size_t)ch * n_in;
        float* dst = interleaved + (size_t)ch * n_out * 2;
        int64_t index = 0;  // CountingSignal::prep resets it
        for (int64_t m = 0; m < n_out; m++) {
            while (c.value >= 1.0) {
                float f = 0.0f;
                if (index < n_in) f = src[index++];
                c.push(f);
                c.value -= 1.0;
            }
            const float x = c.interpolate(c.value);
            c.value += c.ratio;
            dst[2 * m] = x;      // o.fill(x), devices.rs:487-491
            dst[2 * m + 1] = x;
        }
        used = index;
    }
    if (consumed) *consumed = used;
    return OK;
}

}  // extern "C"
