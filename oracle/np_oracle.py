"""Second, independently written restatement of the reference's node arithmetic -- pure numpy, f32.  TEST INFRASTRUCTURE ONLY.

Why it exists: the reference ships no tests or golden vectors and cannot be built in this image, so
parity is UNPINNED by the reference itself.  The C++ oracle (oracle/dsp_oracle.cpp) and the CUDA ops were written from
one reading of the Rust sources; a transcription slip there would pass every GPU-vs-oracle test silently.  This module
was written separately, straight from the Rust text cited per function (paths relative to the reference tree), shares no
code with either, and tests/test_oracle_cross.py requires the two CPU restatements to agree BIT FOR BIT on every
FMA-free node and graph (and within the float-audio tolerance where libm transcendentals are involved: numpy's
tanh/sin/atan/exp are not glibc's).  Slow by design (Python loop per sample for the recurrences): small cases only.

Not restated here: the `gate` extension (no reference node exists; it is defined by dsp_oracle.cpp alone) and `muff`
(source unavailable).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

F = np.float32
BUF_SIZE = 128  # dsp-stuff/src/node.rs:257

# port order: declared input= / output= attributes, then slider(as_input) fields in struct order
# (dsp-stuff-derive/src/lib.rs:191-219); read off the #[dsp(...)] blocks of nodes/*.rs
PORTS: Dict[str, Tuple[Tuple[str, ...], Tuple[str, ...]]] = {
    "gain": (("in", "level"), ("out",)),                       # nodes/gain.rs:5-23
    "distort": (("in", "level"), ("out",)),                    # nodes/distort.rs:30-51
    "overdrive": (("in", "boost", "drive", "level"), ("out",)),  # nodes/overdrive.rs:5-29 (fields: boost, drive, level)
    "chebyshev": (("in",), ("out",)),
    "biquad": (("in",), ("out",)),
    "low_pass": (("in",), ("out",)),
    "high_pass": (("in",), ("out",)),
    "reverb": (("in",), ("out",)),
    "fir": (("in",), ("out",)),
    "add": (("a", "b"), ("out",)),
    "mix": (("a", "b", "ratio"), ("out",)),
    "mux": (("a", "b"), ("out",)),
    "demux": (("in",), ("a", "b")),
    "envelope": (("in",), ("out",)),
    "signal_gen": (("amplitude", "frequency"), ("out",)),
    "input": ((), ("out",)),
    "output": (("in",), ()),
}
# slider defaults and ranges, #[dsp(slider(range = ...), default = ...)]; a field without default = Default::default() = 0.0
DEFAULTS: Dict[str, Dict[str, float]] = {
    "gain": {"level": 1.0},
    "distort": {"level": 0.0},
    "overdrive": {"boost": 0.0, "drive": 0.0, "level": 0.0},
    "chebyshev": {"level_pos": 0.0, "level_neg": 0.0},
    "biquad": {"a0": 1.0, "a1": -0.24, "a2": 0.0, "b0": 0.758, "b1": 0.0, "b2": 0.0},
    "low_pass": {"ratio": 0.5},
    "high_pass": {"ratio": 0.5},
    "reverb": {"seconds": 0.5, "decay": 0.5},
    "mix": {"ratio": 0.5},
    "envelope": {"attack": 0.0, "release": 0.0},
    "signal_gen": {"amplitude": 0.5, "frequency": 100.0},
}
RANGES = {("gain", "level"): (0.0, 10.0), ("distort", "level"): (0.0, 30.0), ("overdrive", "boost"): (0.0, 30.0),
          ("overdrive", "drive"): (0.0, 1.0), ("overdrive", "level"): (0.0, 1.0), ("mix", "ratio"): (0.0, 1.0),
          ("signal_gen", "amplitude"): (-1.0, 1.0), ("signal_gen", "frequency"): (0.1, 20000.0)}
ENUM_DEFAULTS = {"distort": {"mode": "SoftClip"}, "fir": {"mode": "Balanced"}, "mux": {"in_port": "A"},
                 "demux": {"out_port": "A"}, "signal_gen": {"mode": "Sine"}}


def powi(a: np.ndarray, n: int) -> np.ndarray:
    """f32::powi as compiler-rt's __powisf2 evaluates it (square-and-multiply): powi(3) = a * (a*a), powi(4) = (a*a)*(a*a)."""
    r = np.ones_like(a)
    b = n
    while True:
        if b & 1:
            r = r * a
        b >>= 1
        if b == 0:
            break
        a = a * a
    return r


def clip(s):  # nodes/distort.rs:53-61
    return np.where(s < F(-1.0), F(-1.0), np.where(s > F(1.0), F(1.0), s)).astype(F)


def signum(x):  # f32::signum: NaN -> NaN, else copysign(1, x)
    return np.where(np.isnan(x), x, np.copysign(F(1.0), x)).astype(F)


def max_total_cmp_abs(v):
    """max_by(f32::total_cmp) over |v| along the last axis: for non-negative floats total order = bit-pattern order (NaN on top)."""
    bits = np.abs(v).astype(F).view(np.uint32)
    return bits.max(axis=-1, keepdims=True).view(F)


class _Node:
    def __init__(self, typename: str, C: int, granule: int, sample_rate: int):
        self.t = typename
        self.p = {k: F(v) for k, v in DEFAULTS.get(typename, {}).items()}
        self.e = dict(ENUM_DEFAULTS.get(typename, {}))
        self.C, self.granule, self.sr = C, granule, sample_rate
        self.taps = np.array([1.0], dtype=np.float64)       # nodes/fir.rs:61
        self.bq = (F(0.758), F(0.0), F(0.0), F(-0.24), F(0.0))  # b0 b1 b2 a1 a2: BiQuad::initial_filter, biquad.rs:47-60
        self.D = self._ring_len(None)
        self.reset()

    def _ring_len(self, seconds) -> int:
        if seconds is None:       # make_buffer(): circular_buffer::<f32>(128), reverb.rs:44-52
            num = 128
        else:                     # refresh_seconds(): ((seconds * 48000.0) as usize).max(128), reverb.rs:55-58
            prod = F(seconds) * F(self.sr)
            num = max(int(prod) if prod > 0 else 0, 128)   # `as usize` truncates toward zero, saturates negatives to 0
        g = self.granule   # rivulet rounds the capacity up to its page granule (UNPINNED: SURVEY.md section 8c), 1 = nominal
        return (num + g - 1) // g * g if g > 1 else num

    def reset(self):
        C = self.C
        self.x1 = np.zeros(C, F); self.x2 = np.zeros(C, F); self.y1 = np.zeros(C, F); self.y2 = np.zeros(C, F)
        self.z = np.zeros(C, F)
        self.ring = np.zeros((C, self.D), F)      # "filled with zeros": reverb.rs:47-49, 63-66
        self.ring_pos = 0
        self.hist = np.zeros((C, 0), np.float64)  # Fir.state: VecDeque<f64>, starts EMPTY (fir.rs:63-64)
        self.env = np.zeros(C, F)
        self.clock = np.zeros(C, F)

    def set_f32(self, field: str, v: float):
        if field not in self.p:
            raise KeyError(field)
        self.p[field] = F(v)
        if self.t == "biquad":   # after_settings_change = regenerate_filter (biquad.rs:62-76): coeffs / a0 in f32, state reset
            a0 = self.p["a0"]
            self.bq = (self.p["b0"] / a0, self.p["b1"] / a0, self.p["b2"] / a0, self.p["a1"] / a0, self.p["a2"] / a0)
            self.x1[:] = 0; self.x2[:] = 0; self.y1[:] = 0; self.y2[:] = 0
        if self.t == "reverb":   # after_settings_change = refresh_seconds on any slider of the node: new zero-filled ring
            self.D = self._ring_len(self.p["seconds"])
            self.ring = np.zeros((self.C, self.D), F)
            self.ring_pos = 0

    # derive helper <field>_input (dsp-stuff-derive/src/lib.rs:122-161)
    def param(self, field: str, ins: Dict[str, np.ndarray], present: Dict[str, bool]) -> np.ndarray:
        if present.get(field, False):
            lo, hi = RANGES[(self.t, field)]
            x = ins[field]
            y = (x + F(1.0)) / F(2.0)
            z = np.where(np.isnan(y), y, np.minimum(np.maximum(y, F(0.0)), F(1.0))).astype(F)  # f32::clamp keeps NaN
            return (F(lo) + (F(hi) - F(lo)) * z).astype(F)
        return np.full((self.C, BUF_SIZE), self.p[field], F)

    def process(self, ins: Dict[str, np.ndarray], present: Dict[str, bool]) -> Dict[str, np.ndarray]:
        t = self.t
        C = self.C
        if t == "gain":  # nodes/gain.rs:25-38
            return {"out": (ins["in"] * self.param("level", ins, present)).astype(F)}
        if t == "distort":
            return {"out": self._distort(ins["in"], self.param("level", ins, present))}
        if t == "overdrive":  # nodes/overdrive.rs:31-43; the level < 0.001 test is on `level`, mix = drive*d + (1-drive)*x
            x = ins["in"]
            boost, drive, level = (self.param(k, ins, present) for k in ("boost", "drive", "level"))
            a = x * boost
            b = F(0.785398163397448309615660845819875721) * a
            c = np.arctan(b).astype(F)
            d = F(0.636619772367581343075535053490057448) * c
            mix = drive * d + (F(1.0) - drive) * x
            return {"out": np.where(level < F(0.001), x, mix * level).astype(F)}
        if t == "chebyshev":  # nodes/chebyshev.rs:28-42
            x = ins["in"]
            lp, ln = self.p["level_pos"], self.p["level_neg"]
            pos = x if lp < F(0.001) else (np.tanh(x * lp).astype(F) / F(np.tanh(lp)))
            neg = x if ln < F(0.001) else (np.tanh(x * ln).astype(F) / F(np.tanh(ln)))
            return {"out": np.where(x >= F(0.0), pos, neg).astype(F)}
        if t == "biquad":  # biquad 0.4.2 DirectForm1::run: b0*x + b1*x1 + b2*x2 - a1*y1 - a2*y2, left to right, then shift
            b0, b1, b2, a1, a2 = self.bq
            x = ins["in"]
            out = np.empty((C, BUF_SIZE), F)
            x1, x2, y1, y2 = self.x1, self.x2, self.y1, self.y2
            for i in range(BUF_SIZE):
                xi = x[:, i]
                o = b0 * xi + b1 * x1 + b2 * x2 - a1 * y1 - a2 * y2
                x2, x1, y2, y1 = x1, xi, y1, o
                out[:, i] = o
            self.x1, self.x2, self.y1, self.y2 = x1.copy(), x2.copy(), y1.copy(), y2.copy()
            return {"out": out}
        if t in ("low_pass", "high_pass"):  # nodes/low_pass.rs:36-39, nodes/high_pass.rs:36-39
            r = self.p["ratio"]
            omr = F(1.0) - r
            x = ins["in"]
            out = np.empty((C, BUF_SIZE), F)
            z = self.z
            for i in range(BUF_SIZE):
                z = x[:, i] * omr + r * z
                out[:, i] = z if t == "low_pass" else x[:, i] - z
            self.z = z.copy()
            return {"out": out}
        if t == "reverb":  # nodes/reverb.rs:74-111: out = in + ring_front * decay; pop 128; push out  =>  y[n] = x[n] + decay*y[n-D]
            x = ins["in"]
            idx = (self.ring_pos + np.arange(BUF_SIZE)) % self.D
            if BUF_SIZE <= self.D:
                out = (x + self.ring[:, idx] * self.p["decay"]).astype(F)
                self.ring[:, idx] = out
            else:  # cannot happen: D >= 128
                raise AssertionError
            self.ring_pos = (self.ring_pos + BUF_SIZE) % self.D
            return {"out": out}
        if t == "fir":  # nodes/fir.rs:179-225: push, pop if longer than taps, f64 dot of (state, taps) in order, cast, * divisor
            taps = self.taps
            N = len(taps)
            divisor = F(1.0) / F(N) if self.e["mode"] == "Average" else F(1.0)
            x = ins["in"]
            out = np.empty((C, BUF_SIZE), F)
            hist = self.hist
            for i in range(BUF_SIZE):
                hist = np.concatenate([hist, x[:, i:i + 1].astype(np.float64)], axis=1)
                if hist.shape[1] > N:
                    hist = hist[:, 1:]
                prod = hist * taps[None, :hist.shape[1]]
                val = np.cumsum(prod, axis=1)[:, -1].astype(F)   # cumsum is strictly sequential: Iterator::sum order
                out[:, i] = val * divisor
            self.hist = hist
            return {"out": out}
        if t == "add":
            return {"out": (ins["a"] + ins["b"]).astype(F)}
        if t == "mix":  # nodes/mix.rs:45: (b * ratio) + (a * (1.0 - ratio))
            r = self.param("ratio", ins, present)
            return {"out": ((ins["b"] * r) + (ins["a"] * (F(1.0) - r))).astype(F)}
        if t == "mux":
            return {"out": ins["a" if self.e["in_port"] == "A" else "b"].copy()}
        if t == "demux":  # the unselected output keeps the wrapper's zero fill (node.rs:271-275)
            sel = "a" if self.e["out_port"] == "A" else "b"
            other = "b" if sel == "a" else "a"
            return {sel: ins["in"].copy(), other: np.zeros((C, BUF_SIZE), F)}
        if t == "envelope":  # dasp_envelope 0.11.0 Detector<f32, Peak<FullWave>>: gains re-set every block (envelope.rs:45-46)
            def gain(frames):
                return F(0.0) if frames == F(0.0) else F(math.exp(float(F(-1.0) / frames)))
            ga, gr = gain(self.p["attack"]), gain(self.p["release"])
            x = ins["in"]
            out = np.empty((C, BUF_SIZE), F)
            prev = self.env
            for i in range(BUF_SIZE):
                d = np.abs(x[:, i])
                g = np.where(prev < d, ga, gr).astype(F)
                prev = (d + (prev - d) * g).astype(F)
                out[:, i] = prev
            self.env = prev.copy()
            return {"out": out}
        if t == "signal_gen":  # nodes/signal_gen.rs:55-130
            amp = self.param("amplitude", ins, present)
            freq = self.param("frequency", ins, present)
            mode = self.e["mode"]
            if mode == "Constant":
                return {"out": amp.copy()}
            out = np.empty((C, BUF_SIZE), F)
            total = np.zeros(C, F)
            clock = self.clock
            for i in range(BUF_SIZE):
                step = freq[:, i] / F(48000.0)
                total = (total + step).astype(F)
                if mode == "Sine":
                    v = np.sin((clock + total) * F(6.28318530717958647692528676655900577)).astype(F)
                elif mode == "Triangle":
                    v = F(2.0) * np.fmod(clock + total, F(1.0)) - F(1.0)
                else:  # Square: compares `total`, not clock + total (signal_gen.rs:93) -- reproduced as written
                    v = np.where(total > F(0.5), F(1.0), F(-1.0)).astype(F)
                out[:, i] = v * amp[:, i]
            self.clock = np.fmod(clock + total, F(1.0)).astype(F)
            return {"out": out}
        raise KeyError(t)

    def _distort(self, x, level):  # nodes/distort.rs:63-196
        mode = self.e["mode"]
        if mode == "Fuzz":  # no level < 0.001 guard; per 128-sample block
            mx = max_total_cmp_abs(x)
            with np.errstate(all="ignore"):
                q = clip(x * level) / mx
                z = np.copysign(F(1.0) - np.exp(np.copysign(q, F(-1.0))).astype(F), F(-1.0)).astype(F)
                mz = max_total_cmp_abs(z)
                y = clip(z * mx) / mz
                my = max_total_cmp_abs(y)
                return (y * mx / my).astype(F)
        s = x * level
        with np.errstate(all="ignore"):
            if mode == "HardClip":
                y = clip(s) / level
            elif mode == "SoftClip":
                inner = np.where(s > F(1.0), F(2.0) / F(3.0),
                                 np.where((s >= F(-1.0)) & (s <= F(1.0)), s - (powi(s, 3) / F(3.0)), F(-2.0) / F(3.0))).astype(F)
                y = clip(inner) / level
            elif mode == "Tanh":
                y = np.tanh(s)
            elif mode == "RecipSoftClip":
                y = signum(x) * (F(1.0) - F(1.0) / (np.abs(x) * level + F(1.0)))
            elif mode == "Sin":
                y = np.sin(s)
            elif mode == "Atan":
                y = np.arctan(s)
            elif mode == "Square":
                y = powi(s, 2) * signum(s)
            elif mode == "Chebyshev4":
                y = F(8.0) * powi(s, 4) - F(8.0) * powi(s, 2) + F(1.0)
            else:
                raise KeyError(mode)
        return np.where(level < F(0.001), x, y).astype(F)


class NpOracle:
    """Same graph-building surface as the C++ oracle / the Engine: add_node, set_f32, set_enum, set_taps, link, compile."""

    def __init__(self, channels: int, ring_granule: int = 1024, sample_rate: int = 48000):
        self.C, self.granule, self.sr = channels, ring_granule, sample_rate
        self.nodes: Dict[int, _Node] = {}
        self.links: List[Tuple[int, str, int, str]] = []
        self.order: List[int] = []

    def add_node(self, typename: str, node_id: int):
        if typename not in PORTS:
            raise KeyError(typename)
        self.nodes[node_id] = _Node(typename, self.C, self.granule, self.sr)

    def set_f32(self, node_id: int, field: str, value: float):
        self.nodes[node_id].set_f32(field, value)

    def set_enum(self, node_id: int, field: str, variant: str):
        self.nodes[node_id].e[field] = variant

    def set_taps(self, node_id: int, taps):
        nd = self.nodes[node_id]
        nd.taps = np.asarray(taps, dtype=np.float64).copy()
        nd.hist = np.zeros((self.C, 0), np.float64)

    def link(self, src: int, out_port: str, dst: int, in_port: str):
        assert out_port in PORTS[self.nodes[src].t][1] and in_port in PORTS[self.nodes[dst].t][0]
        self.links.append((src, out_port, dst, in_port))

    def compile(self):
        # a block-synchronous topological order is result-equivalent to the reference's dataflow execution on a DAG
        indeg = {i: 0 for i in self.nodes}
        for (_, _, d, _) in self.links:
            indeg[d] += 1
        ready = [i for i in self.nodes if indeg[i] == 0]
        order = []
        while ready:
            n = ready.pop(0)
            order.append(n)
            for (s, _, d, _) in self.links:
                if s == n:
                    indeg[d] -= 1
                    if indeg[d] == 0:
                        ready.append(d)
        assert len(order) == len(self.nodes), "cycle"
        linked = {s for (s, _, _, _) in self.links} | {d for (_, _, d, _) in self.links}
        self.order = [i for i in order if i in linked]   # a node with no links never runs (runtime.rs:661-668)
        self.in_terms = [i for i, n in self.nodes.items() if n.t == "input"]
        self.out_terms = [i for i, n in self.nodes.items() if n.t == "output"]

    def reset_state(self):
        for n in self.nodes.values():
            n.reset()

    def _collect_and_average(self, vals: Dict[Tuple[int, str], np.ndarray], node: int, port: str):
        """node.rs:162-194: buf (zeros) += every delivering link, in order; num_frames = 0.0001 (+ 1.0 per link); buf /= num_frames"""
        buf = np.zeros((self.C, BUF_SIZE), F)
        nf = F(0.0001)
        present = False
        for (s, op, d, ip) in self.links:
            if d == node and ip == port:
                buf = (buf + vals[(s, op)]).astype(F)
                nf = F(nf + F(1.0))
                present = True
        return (buf / nf).astype(F), present

    def process(self, inputs, n: Optional[int] = None) -> List[np.ndarray]:
        if isinstance(inputs, np.ndarray):
            inputs = [inputs]
        xs = [np.ascontiguousarray(x, dtype=F) for x in inputs]
        n = xs[0].shape[1] if xs else int(n)
        assert n % BUF_SIZE == 0
        outs = [np.zeros((self.C, n), F) for _ in self.out_terms]
        for b in range(n // BUF_SIZE):
            sl = slice(b * BUF_SIZE, (b + 1) * BUF_SIZE)
            vals: Dict[Tuple[int, str], np.ndarray] = {}
            for nid in self.order:
                nd = self.nodes[nid]
                if nd.t == "input":   # nodes/input.rs:226: raw samples to every link
                    vals[(nid, "out")] = xs[self.in_terms.index(nid)][:, sl]
                    continue
                ins, present = {}, {}
                for p in PORTS[nd.t][0]:
                    ins[p], present[p] = self._collect_and_average(vals, nid, p)
                if nd.t == "output":  # nodes/output.rs:223: the sink averages its links once more
                    outs[self.out_terms.index(nid)][:, sl] = ins["in"]
                    continue
                with np.errstate(all="ignore"):
                    for q, v in nd.process(ins, present).items():
                        vals[(nid, q)] = v
        return outs

    def process_n(self, n: int) -> List[np.ndarray]:
        return self.process([], n)


class NpResampler:
    """Independent restatement (pure Python / numpy scalars) of the playback-side converter: dasp_signal 0.11.0
    `Converter::from_hz_to_hz(CountingSignal, Sinc::new(Fixed::from([0.0; 16])), 48_000.0, target)` as built at devices.rs:550-556
    and driven by `do_write_2` (devices.rs:443-500).  Restated from the published crates (parity UNPINNED); cross-checked
    bit for bit against the C++ oracle's restatement in tests/test_oracle_cross.py.  One channel at a time, slow."""

    DEPTH = 8

    def __init__(self, channels: int, target_hz: float, source_hz: float = 48000.0):
        self.C = channels
        self.ratio = float(source_hz) / float(target_hz)
        self.value = [0.0] * channels                      # Converter::interpolation_value
        self.frames = [[F(0.0)] * 16 for _ in range(channels)]   # ring_buffer::Fixed: index 0 = oldest
        self.idx = [0] * channels                          # Sinc::idx

    def _interpolate(self, frames, idx, x):
        depth = self.DEPTH
        nl, nr = idx, idx + 1
        rightmost, leftmost = nl + depth, nr - depth
        max_depth = (16 - depth) if rightmost >= 16 else ((depth + leftmost) if leftmost < 0 else depth)
        v = F(0.0)
        for n in range(max_depth):
            for phi, i in ((x + n, nl - n), ((1.0 - x) + n, nr + n)):
                a = math.pi * phi
                first = 1.0 if a == 0.0 else math.sin(a) / a
                second = 0.5 + 0.5 * math.cos(a / depth)
                v = F(v + F(first * second * float(frames[i % 16])))   # Fixed indexes modulo its length
        return v

    def process(self, mono, n_out: int):
        x = np.ascontiguousarray(mono, dtype=F)
        out = np.zeros((self.C, n_out, 2), F)
        used = 0
        for c in range(self.C):
            index = 0
            fr, idx, val = self.frames[c], self.idx[c], self.value[c]
            for m in range(n_out):
                while val >= 1.0:
                    f = F(0.0)
                    if index < x.shape[1]:
                        f = x[c, index]
                        index += 1
                    fr = fr[1:] + [f]           # push: the oldest frame drops out
                    if idx < self.DEPTH:
                        idx += 1
                    val -= 1.0
                y = self._interpolate(fr, idx, val)
                val += self.ratio
                out[c, m, 0] = y
                out[c, m, 1] = y
            self.frames[c], self.idx[c], self.value[c] = fr, idx, val
            used = index
        return out, used
