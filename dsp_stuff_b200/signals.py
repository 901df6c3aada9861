"""Synthetic signals and the BASELINE.json workloads, bit-identical wherever generated.

noise : x[c][n] = float(int32(splitmix64(seed ^ (c<<32 | n)) >> 40) - 2^23) * 2^-23 * 0.5
        uniform in [-0.5, 0.5), every value exactly representable in f32.
sweep : log sine 20 Hz -> 20 kHz over 10 s, amplitude 0.5, phase offset 2*pi*c/C, f64 then f32.
"""
from __future__ import annotations

import math

import numpy as np

from .graph import GraphSpec

SAMPLE_RATE = 48000
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def noise(channels: int, n: int, seed: int = 42, channel_offset: int = 0, sample_offset: int = 0) -> np.ndarray:
    c = (np.arange(channels, dtype=np.uint64) + np.uint64(channel_offset))[:, None] << np.uint64(32)
    i = (np.arange(n, dtype=np.uint64) + np.uint64(sample_offset))[None, :]
    h = _splitmix64(np.uint64(seed) ^ (c | i))
    v = (h >> np.uint64(40)).astype(np.int64) - (1 << 23)
    return (v.astype(np.float32) * np.float32(2.0 ** -23) * np.float32(0.5)).astype(np.float32)


def sweep(channels: int, n: int, total_channels: int | None = None, channel_offset: int = 0,
          sample_offset: int = 0) -> np.ndarray:
    C = total_channels or channels
    t = (np.arange(n, dtype=np.float64) + sample_offset) / SAMPLE_RATE
    f0, f1, T = 20.0, 20000.0, 10.0
    k = math.log(f1 / f0)
    phase = 2.0 * math.pi * f0 * T / k * (np.exp(t / T * k) - 1.0)
    ch = (np.arange(channels, dtype=np.float64) + channel_offset)[:, None]
    return (0.5 * np.sin(phase[None, :] + 2.0 * math.pi * ch / C)).astype(np.float32)


def impulse(channels: int, n: int, at: int = 0, amplitude: float = 1.0) -> np.ndarray:
    x = np.zeros((channels, n), dtype=np.float32)
    x[:, at] = amplitude
    return x


def rbj_biquad(kind: str, fc: float, q: float = 0.7071, fs: float = SAMPLE_RATE):
    """RBJ cookbook LP/HP, computed in f64, rounded to f32, normalised so a0 = 1."""
    w0 = 2.0 * math.pi * fc / fs
    alpha = math.sin(w0) / (2.0 * q)
    cw = math.cos(w0)
    if kind == "lp":
        b0, b1, b2 = (1 - cw) / 2, 1 - cw, (1 - cw) / 2
    elif kind == "hp":
        b0, b1, b2 = (1 + cw) / 2, -(1 + cw), (1 + cw) / 2
    else:
        raise ValueError(kind)
    a0, a1, a2 = 1 + alpha, -2 * cw, 1 - alpha
    f = lambda v: float(np.float32(v / a0))
    return dict(a0=1.0, a1=f(a1), a2=f(a2), b0=f(b0), b1=f(b1), b2=f(b2))


def reverb_ir(n_taps: int = 4096, seed: int = 7) -> np.ndarray:
    """Config-4 impulse response h: noise(seed 7) * exp(-6.9 i / N), unit l2 energy, f64 (NOT reversed)."""
    x = noise(1, n_taps, seed=seed)[0].astype(np.float64) * 2.0
    h = x * np.exp(-6.9 * np.arange(n_taps, dtype=np.float64) / n_taps)
    return h / math.sqrt(float(np.sum(h * h)))


# ---- BASELINE.json configs as graphs -------------------------------------------------------------------
def _io(g: GraphSpec, first: int, last: int) -> GraphSpec:
    g.node(1000, "input").node(1001, "output")
    g.link(1000, "out", first, "in").link(last, "out", 1001, "in")
    return g


def random_graph(seed, n_nodes=None):
    """A random DAG over the FMA-free node types: every input port gets 1-3 links from earlier nodes (fan-in averaging with
    different divisors, fan-out by reuse), control ports are sometimes driven, two sinks."""
    rng = np.random.default_rng(seed)
    from .graph import NODE_PORTS as ports

    g = GraphSpec().node(100, "input").node(101, "input")
    outs = [(100, "out"), (101, "out")]
    kinds = ["gain", "distort", "biquad", "low_pass", "high_pass", "reverb", "add", "mix", "mux", "demux", "envelope", "fir"]
    for nid in range(int(rng.integers(4, 9)) if n_nodes is None else n_nodes):
        t = kinds[int(rng.integers(len(kinds)))]
        params = {}
        taps = None
        if t == "gain":
            params = dict(level=float(rng.uniform(0.2, 3.0)))
        elif t == "distort":
            params = dict(mode=["HardClip", "SoftClip", "RecipSoftClip", "Square", "Chebyshev4"][int(rng.integers(5))], level=float(rng.uniform(0.0, 6.0)))
        elif t == "biquad":
            r, th = float(rng.uniform(0.1, 0.95)), float(rng.uniform(0.1, 3.0))
            params = dict(a0=float(rng.uniform(0.5, 2.0)), a1=-2 * r * np.cos(th), a2=r * r, b0=float(rng.uniform(0.1, 1.0)), b1=float(rng.uniform(-0.5, 0.5)), b2=float(rng.uniform(-0.5, 0.5)))
        elif t in ("low_pass", "high_pass"):
            params = dict(ratio=float(rng.uniform(0.0, 0.99)))
        elif t == "reverb":
            params = dict(seconds=float(rng.uniform(0.003, 0.012)), decay=float(rng.uniform(0.1, 0.9)))
        elif t == "mix":
            params = dict(ratio=float(rng.uniform(0, 1)))
        elif t == "mux":
            params = dict(in_port=["A", "B"][int(rng.integers(2))])
        elif t == "demux":
            params = dict(out_port=["A", "B"][int(rng.integers(2))])
        elif t == "envelope":
            params = dict(attack=float(rng.integers(0, 50)), release=float(rng.integers(0, 400)))
        elif t == "fir":
            params = dict(mode=["Average", "Balanced"][int(rng.integers(2))])
            taps = rng.uniform(-1, 1, int(rng.integers(1, 40)))
        g.node(nid, t, taps=taps, **params)
        ins, node_outs = ports[t]
        for p in ins:
            control = p not in ("in", "a", "b")
            if control and rng.uniform() < 0.6:
                continue   # slider value
            for _ in range(int(rng.integers(1, 4)) if not control else 1):
                s, sp = outs[int(rng.integers(len(outs)))]
                g.link(s, sp, nid, p)
        outs += [(nid, q) for q in node_outs]
    g.node(200, "output").node(201, "output")
    for sink in (200, 201):
        for _ in range(int(rng.integers(1, 4))):
            s, sp = outs[int(rng.integers(2, len(outs)))] if len(outs) > 2 else outs[0]
            g.link(s, sp, sink, "in")
    return g


def config1() -> GraphSpec:
    g = GraphSpec()
    g.node(0, "gain", level=2.0).node(1, "distort", mode="SoftClip", level=4.0)
    g.node(2, "reverb", seconds=0.25, decay=0.5)
    return _io(g.chain([0, 1, 2]), 0, 2)


def config2(one_pole: bool = False) -> GraphSpec:
    g = GraphSpec()
    if one_pole:
        g.node(0, "low_pass", ratio=0.9).node(1, "high_pass", ratio=0.99)
    else:
        g.node(0, "biquad", **rbj_biquad("lp", 1000.0)).node(1, "biquad", **rbj_biquad("hp", 200.0))
    return _io(g.chain([0, 1]), 0, 1)


def config3() -> GraphSpec:
    g = GraphSpec()
    g.node(0, "gain", level=2.0).node(1, "distort", mode="SoftClip", level=4.0)
    g.node(2, "biquad", **rbj_biquad("lp", 1000.0)).node(3, "reverb", seconds=0.25, decay=0.5)
    return _io(g.chain([0, 1, 2, 3]), 0, 3)


def config4(n_taps: int = 4096) -> GraphSpec:
    g = GraphSpec()
    g.node(0, "fir", mode="Balanced", taps=reverb_ir(n_taps)[::-1].copy())
    return _io(g, 0, 0)


def target_chain(n_taps: int = 4096) -> GraphSpec:
    """north_star target: gain -> distortion -> biquad -> delay(250 ms) -> FIR(4096)."""
    g = GraphSpec()
    g.node(0, "gain", level=2.0).node(1, "distort", mode="SoftClip", level=4.0)
    g.node(2, "biquad", **rbj_biquad("lp", 1000.0)).node(3, "reverb", seconds=0.25, decay=0.5)
    g.node(4, "fir", mode="Balanced", taps=reverb_ir(n_taps)[::-1].copy())
    return _io(g.chain([0, 1, 2, 3, 4]), 0, 4)


def config5(n_taps: int = 4096) -> GraphSpec:
    """Full effect graph: fan-out to path A {gain, distort(Tanh), fir}, path B {biquad, demux/mux,
    reverb .25, reverb .125} and dry; mix(A,B) -> add(mix, dry) -> sink fed by add.out AND path A."""
    g = GraphSpec()
    g.node(100, "input").node(101, "output")
    g.node(0, "gain", level=2.0).node(1, "distort", mode="Tanh", level=4.0)
    g.node(2, "fir", mode="Balanced", taps=reverb_ir(n_taps)[::-1].copy())
    g.node(3, "biquad", **rbj_biquad("lp", 1000.0))
    g.node(4, "demux", out_port="B").node(5, "mux", in_port="B")
    g.node(6, "reverb", seconds=0.25, decay=0.5).node(7, "reverb", seconds=0.125, decay=0.5)
    g.node(8, "mix", ratio=0.5).node(9, "add")
    g.link(100, "out", 0, "in").link(100, "out", 3, "in").link(100, "out", 9, "b")
    g.link(0, "out", 1, "in").link(1, "out", 2, "in")
    g.link(3, "out", 4, "in").link(4, "a", 5, "a").link(4, "b", 5, "b").link(5, "out", 6, "in").link(6, "out", 7, "in")
    g.link(2, "out", 8, "a").link(7, "out", 8, "b").link(8, "out", 9, "a")
    g.link(9, "out", 101, "in").link(2, "out", 101, "in")
    return g


WORKLOADS = {
    "config1": (config1, 16),
    "config2": (config2, 8),
    "config2_one_pole": (lambda: config2(one_pole=True), 8),   # LowPass(0.9) -> HighPass(0.99), SURVEY §8d config 2 variant
    "config3": (config3, 16),
    "config4": (config4, 8),
    "config5": (config5, 24),
    "target": (target_chain, 16),
}
"""name -> (graph factory, ALGORITHMIC bytes per channel-sample: 4 B per external input + 4 B per
external output + 8 B per Reverb node; SURVEY.md §8d)."""
