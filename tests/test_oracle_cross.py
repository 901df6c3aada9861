"""Two independently written CPU restatements of the reference must agree (CPU only).

The reference has no tests, fixtures or golden vectors, and cannot be built here: parity is unpinned
by the reference.  What CAN be pinned is the reading of its source: oracle/np_oracle.py (pure numpy f32, written straight
from the Rust text) and oracle/dsp_oracle.cpp (C++, the oracle the GPU tests use) share no code.  Every FMA-free node
and every BASELINE graph has to come out bit-identical from both; nodes that call libm transcendentals (numpy's
implementations are not glibc's) within the float-audio tolerance."""
import numpy as np
import pytest

from dsp_stuff_b200 import GraphSpec
from dsp_stuff_b200 import signals as S
from oracle.np_oracle import NpOracle
from tests.test_oracle_kat import single
from tests.util import assert_audio_close, assert_bit_exact, make_oracle


def both(oracle_mod, spec, xs, C, n=None, granule=1024, chunks=1):
    a = make_oracle(oracle_mod, spec, C, ring_granule=granule, threads=1)
    b = NpOracle(C, ring_granule=granule)
    spec.apply(b)
    xs = [xs] if isinstance(xs, np.ndarray) else list(xs)
    if xs:
        n = xs[0].shape[1]
    step = n // chunks
    ya, yb = [], []
    for k in range(chunks):   # state carried across calls in both
        part = [x[:, k * step:(k + 1) * step] for x in xs]
        ya.append(a.process(part) if xs else a.process_n(step))
        yb.append(b.process(part) if xs else b.process_n(step))
    cat = lambda ys: [np.concatenate([y[i] for y in ys], axis=1) for i in range(len(ys[0]))]
    return cat(ya), cat(yb)


EXACT = [
    ("gain", dict(level=2.5)),
    ("distort", dict(mode="HardClip", level=4.0)),
    ("distort", dict(mode="SoftClip", level=4.0)),
    ("distort", dict(mode="SoftClip", level=0.0005)),       # level < 0.001: bypass
    ("distort", dict(mode="RecipSoftClip", level=7.0)),
    ("distort", dict(mode="Square", level=3.0)),
    ("distort", dict(mode="Chebyshev4", level=1.5)),
    ("biquad", dict()),
    ("biquad", S.rbj_biquad("lp", 1000.0)),
    ("biquad", S.rbj_biquad("hp", 200.0)),
    ("biquad", dict(a0=2.0, a1=-0.5, a2=0.25, b0=1.0, b1=0.5, b2=0.25)),
    ("low_pass", dict(ratio=0.9)),
    ("high_pass", dict(ratio=0.99)),
    ("reverb", dict(seconds=0.01, decay=0.7)),
    ("reverb", dict()),                                      # fresh node: make_buffer() ring
    ("envelope", dict(attack=20.0, release=400.0)),
    ("envelope", dict()),
    ("fir", dict()),
]
APPROX = [
    ("distort", dict(mode="Tanh", level=6.0)),
    ("distort", dict(mode="Sin", level=9.0)),
    ("distort", dict(mode="Atan", level=30.0)),
    ("distort", dict(mode="Fuzz", level=4.0)),
    ("overdrive", dict(boost=12.0, drive=0.7, level=0.8)),
    ("chebyshev", dict(level_pos=5.0, level_neg=2.0)),
]
_ids = lambda v: str(v) if isinstance(v, str) else "-".join(f"{k}{v}" for k, v in v.items())[:40]


@pytest.mark.parametrize("typename,params", EXACT, ids=_ids)
def test_single_node_bit_exact(oracle_mod, typename, params):
    x = S.noise(3, 128 * 12)
    ya, yb = both(oracle_mod, single(typename, **params), x, 3, chunks=3)
    assert_bit_exact(yb[0], ya[0], f"{typename} {params}: numpy restatement vs C++ oracle")


@pytest.mark.parametrize("typename,params", APPROX, ids=_ids)
def test_single_node_with_transcendentals(oracle_mod, typename, params):
    x = S.noise(3, 128 * 8)
    ya, yb = both(oracle_mod, single(typename, **params), x, 3)
    assert_audio_close(yb[0], ya[0], what=f"{typename} {params}")


def test_fir_taps_modes_and_warm_up(oracle_mod):
    x = S.noise(2, 128 * 6)
    for mode in ("Balanced", "Average"):
        g = single("fir", mode=mode)
        g.nodes[0].taps = S.reverb_ir(200)[::-1].copy()        # warm-up (n < N - 1) spans two blocks
        ya, yb = both(oracle_mod, g, x, 2, chunks=2)
        assert_bit_exact(yb[0], ya[0], f"fir {mode}")


@pytest.mark.parametrize("seconds,granule", [(0.25, 1024), (0.25, 1), (0.004, 1), (0.0301, 1)])
def test_reverb_ring_length_and_feedback(oracle_mod, seconds, granule):
    spec = single("reverb", seconds=seconds, decay=0.5)
    a = make_oracle(oracle_mod, spec, 1, ring_granule=granule)
    b = NpOracle(1, ring_granule=granule)
    spec.apply(b)
    D = b.nodes[0].D
    assert a.get_i64(0, "delay_samples") == D                 # integer index work: both readings of reverb.rs:55-68
    n = 128 * ((2 * D) // 128 + 2) if D <= 2048 else 128 * 4
    x = S.noise(1, n)
    assert_bit_exact(b.process(x)[0], a.process(x)[0], f"reverb D={D}")


@pytest.mark.parametrize("name,C,n", [("config1", 2, 128 * 110), ("config2", 2, 128 * 12), ("config2_one_pole", 2, 128 * 12),
                                      ("config3", 2, 128 * 110)])
def test_baseline_chains_bit_exact(oracle_mod, name, C, n):
    spec = S.WORKLOADS[name][0]()
    x = S.noise(C, n)
    ya, yb = both(oracle_mod, spec, x, C)
    assert_bit_exact(yb[0], ya[0], name)


def test_target_chain_short_taps_bit_exact(oracle_mod):
    x = S.noise(2, 128 * 104)   # longer than the 12288-sample comb: the feedback path is exercised
    ya, yb = both(oracle_mod, S.target_chain(n_taps=48), x, 2)
    assert_bit_exact(yb[0], ya[0], "target chain")


def test_fan_in_fan_out_mix_add_mux_demux_bit_exact(oracle_mod):
    g = GraphSpec().node(10, "input").node(12, "input").node(11, "output").node(13, "output")
    g.node(0, "gain", level=2.0).node(1, "mix", ratio=0.25).node(2, "add").node(3, "demux", out_port="B").node(4, "mux", in_port="B")
    g.link(10, "out", 0, "in").link(10, "out", 1, "a").link(12, "out", 1, "b").link(0, "out", 2, "a").link(1, "out", 2, "b")
    g.link(2, "out", 3, "in").link(3, "a", 4, "a").link(3, "b", 4, "b").link(4, "out", 11, "in").link(0, "out", 11, "in")
    g.link(12, "out", 11, "in").link(3, "a", 13, "in")
    xs = [S.noise(3, 512), S.noise(3, 512, seed=5)]
    ya, yb = both(oracle_mod, g, xs, 3)
    assert_bit_exact(yb[0], ya[0], "three-link fan-in sink")
    assert_bit_exact(yb[1], ya[1], "demux unselected port")


def test_modulated_parameters_bit_exact(oracle_mod):
    g = GraphSpec().node(10, "input").node(12, "input").node(11, "output")
    g.node(0, "gain").node(1, "distort", mode="HardClip").node(2, "mix")
    g.link(10, "out", 0, "in").link(12, "out", 0, "level").link(0, "out", 1, "in").link(12, "out", 1, "level")
    g.link(1, "out", 2, "a").link(10, "out", 2, "b").link(12, "out", 2, "ratio").link(2, "out", 11, "in")
    xs = [S.noise(3, 512), S.sweep(3, 512) * 2.5]
    ya, yb = both(oracle_mod, g, xs, 3)
    assert_bit_exact(yb[0], ya[0], "control ports (lib.rs:122-161)")


@pytest.mark.parametrize("mode,exact", [("Triangle", True), ("Square", True), ("Constant", True), ("Sine", False)])
def test_signal_gen(oracle_mod, mode, exact):
    g = GraphSpec().node(0, "signal_gen", mode=mode, amplitude=0.5, frequency=997.0).node(1, "gain", level=1.5).node(11, "output")
    g.link(0, "out", 1, "in").link(1, "out", 11, "in")
    ya, yb = both(oracle_mod, g, [], 2, n=128 * 30, chunks=3)
    if exact:
        assert_bit_exact(yb[0], ya[0], f"signal_gen {mode}")
    else:
        assert_audio_close(yb[0], ya[0], what=f"signal_gen {mode}")


def test_config5_graph_with_tanh(oracle_mod):
    x = S.noise(2, 128 * 100)
    ya, yb = both(oracle_mod, S.config5(n_taps=32), x, 2)
    assert_audio_close(yb[0], ya[0], what="config5 graph")   # path A holds a Tanh distortion (libm vs numpy)


@pytest.mark.parametrize("target", [44100.0, 96000.0, 48000.0, 22050.0])
def test_resampler_restatements_agree(oracle_mod, target):
    """dasp Converter + Sinc<[f32; 16]> (devices.rs:443-500, 550-556): the C++ and the Python restatement, two calls each."""
    from oracle.np_oracle import NpResampler

    x = S.noise(2, 700, seed=11)
    a, b = oracle_mod.Resampler(2, target), NpResampler(2, target)
    off_a = off_b = 0
    for n_out in (300, 150):
        ya, ua = a.process(x[:, off_a:], n_out)
        yb, ub = b.process(x[:, off_b:], n_out)
        assert ua == ub
        assert_bit_exact(yb.reshape(2, -1), ya.reshape(2, -1), f"resampler {target}")
        off_a += ua
        off_b += ub


@pytest.mark.parametrize("seed", range(24))
def test_random_graphs_bit_exact(oracle_mod, seed):
    """Random DAGs (fan-in of 1-3 links, fan-out, control ports, demux zeros, nested recurrences): the two restatements agree
    bit for bit, sample for sample, across three calls."""
    g = S.random_graph(seed)
    xs = [S.noise(2, 128 * 9, seed=seed + 1), S.sweep(2, 128 * 9) * 1.5]
    ya, yb = both(oracle_mod, g, xs, 2, chunks=3)
    for k in range(2):
        assert_bit_exact(yb[k], ya[k], f"random graph {seed}, sink {k}")


@pytest.mark.parametrize("seed,n_nodes", [(2000, 30), (2000, 43), (2001, 57), (2001, 69)])
def test_large_random_graphs_bit_exact(oracle_mod, seed, n_nodes):
    """The graphs tests/test_gpu_zz_random_graphs.py runs on the GPU (cut into up to 15 segments there): the two restatements
    agree on them first."""
    g = S.random_graph(seed, n_nodes)
    xs = [S.noise(3, 128 * 9, seed=seed + 1), S.sweep(3, 128 * 9) * 1.5]
    ya, yb = both(oracle_mod, g, xs, 3, chunks=3)
    for k in range(2):
        assert np.isfinite(ya[k]).all() and 0.05 < np.abs(ya[k]).max() < 100
        assert_bit_exact(yb[k], ya[k], f"{n_nodes}-node random graph {seed}, sink {k}")
