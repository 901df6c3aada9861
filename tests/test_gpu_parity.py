"""GPU parity: the CUDA engine (through the C ABI) against the CPU oracle on identical seeded inputs.

Bit-exact: everything whose reference arithmetic is FMA-free +,-,*,/ (gain, clip shapers, add, mix,
mux/demux, fan-in averages, biquad / one-pole / envelope recurrences, the comb, the direct FIR) and
all index work.  Tolerance (tests/util.py: 1e-5 peak-relative and -100 dBFS rms): nodes that call
libm transcendentals (CUDA's tanhf/sinf/atanf/expf differ from glibc's by an ulp or two) and the
FFT FIR path."""
import numpy as np
import pytest

from dsp_stuff_b200 import GraphSpec
from dsp_stuff_b200 import signals as S
from tests.test_oracle_kat import single
from tests.util import assert_audio_close, assert_bit_exact, make_oracle

pytestmark = pytest.mark.gpu

FIR_DIRECT = 1


def make_engine(spec, channels, max_samples, ring_granule=1024, fir_mode=FIR_DIRECT):
    from dsp_stuff_b200.engine import Engine

    e = Engine(channels, block=128, max_samples=max_samples, ring_granule=ring_granule, fir_mode=fir_mode)
    spec.apply(e)
    return e


def run_both(oracle_mod, spec, x, chunks=None, ring_granule=1024, fir_mode=FIR_DIRECT):
    """x: array or list of arrays [C x n].  chunks: list of chunk lengths the GPU call sequence uses."""
    xs = [x] if isinstance(x, np.ndarray) else list(x)
    C, n = xs[0].shape
    ref = make_oracle(oracle_mod, spec, C, ring_granule=ring_granule).process(xs)
    chunks = chunks or [n]
    assert sum(chunks) == n
    eng = make_engine(spec, C, max(chunks), ring_granule, fir_mode)
    outs = None
    off = 0
    for c in chunks:
        part = eng.process([a[:, off:off + c] for a in xs])
        outs = part if outs is None else [np.concatenate([o, p], axis=1) for o, p in zip(outs, part)]
        off += c
    return outs, ref, eng


EXACT_NODES = [
    ("gain", dict(level=2.5)),
    ("distort", dict(mode="HardClip", level=4.0)),
    ("distort", dict(mode="SoftClip", level=4.0)),
    ("distort", dict(mode="SoftClip", level=0.0)),
    ("distort", dict(mode="RecipSoftClip", level=7.0)),
    ("distort", dict(mode="Square", level=3.0)),
    ("distort", dict(mode="Chebyshev4", level=1.5)),
    ("biquad", dict()),
    ("biquad", S.rbj_biquad("hp", 200.0)),
    ("biquad", dict(a0=2.0, a1=-0.5, a2=0.25, b0=1.0, b1=0.5, b2=0.25)),
    ("low_pass", dict(ratio=0.9)),
    ("high_pass", dict(ratio=0.99)),
    ("reverb", dict(seconds=0.01, decay=0.7)),
    ("envelope", dict(attack=20.0, release=400.0)),
    ("envelope", dict()),
    ("gate", dict(threshold=0.3, attack=5.0, release=200.0)),   # extension node: parity is against the oracle's definition only
    ("gate", dict(threshold=0.45)),
    ("gate", dict()),
    ("fir", dict()),
]
APPROX_NODES = [
    ("distort", dict(mode="Tanh", level=6.0)),
    ("distort", dict(mode="Sin", level=9.0)),
    ("distort", dict(mode="Atan", level=30.0)),
    ("distort", dict(mode="Fuzz", level=4.0)),
    ("overdrive", dict(boost=12.0, drive=0.7, level=0.8)),
    ("chebyshev", dict(level_pos=5.0, level_neg=2.0)),
]


@pytest.mark.parametrize("typename,params", EXACT_NODES, ids=lambda v: str(v) if isinstance(v, str) else "-".join(f"{k}{v}" for k, v in v.items())[:40])
def test_node_bit_exact(oracle_mod, typename, params):
    x = S.noise(5, 128 * 37)   # odd channel count, several tiles, partial last tile
    got, ref, _ = run_both(oracle_mod, single(typename, **params), x)
    assert_bit_exact(got[0], ref[0], f"{typename} {params}")


@pytest.mark.parametrize("typename,params", APPROX_NODES, ids=lambda v: str(v) if isinstance(v, str) else "-".join(f"{k}{v}" for k, v in v.items())[:40])
def test_node_within_tolerance(oracle_mod, typename, params):
    x = S.noise(5, 128 * 37)
    got, ref, _ = run_both(oracle_mod, single(typename, **params), x)
    assert_audio_close(got[0], ref[0], what=f"{typename} {params}")


def test_fuzz_zero_block_nan_positions_match(oracle_mod):
    x = S.noise(3, 1024)
    x[1, 256:384] = 0.0        # an all-zero reference block => NaN for exactly those 128 samples (distort.rs:158)
    got, ref, _ = run_both(oracle_mod, single("distort", mode="Fuzz", level=4.0), x)
    assert np.array_equal(np.isnan(got[0]), np.isnan(ref[0]))
    assert np.isnan(got[0][1, 256:384]).all() and np.isnan(got[0]).sum() == 128
    m = ~np.isnan(ref[0])
    assert_audio_close(got[0][m], ref[0][m], what="fuzz")


def test_fir_direct_warmup_and_steady_state_bit_exact(oracle_mod):
    g = S.config4(n_taps=300)
    x = S.noise(3, 128 * 9)
    got, ref, _ = run_both(oracle_mod, g, x, chunks=[128, 256, 128 * 6])   # warm-up spans several calls
    assert_bit_exact(got[0], ref[0], "fir direct")
    ga = single("fir", mode="Average")
    ga.nodes[0].taps = S.reverb_ir(64)[::-1].copy()
    got, ref, _ = run_both(oracle_mod, ga, x)
    assert_bit_exact(got[0], ref[0], "fir average")


@pytest.mark.parametrize("name,C,n", [("config1", 2, 128 * 200), ("config3", 64, 128 * 120), ("config2", 16, 128 * 64)])
def test_config_chains_bit_exact(oracle_mod, name, C, n):
    spec = S.WORKLOADS[name][0]()
    x = S.noise(C, n)
    got, ref, eng = run_both(oracle_mod, spec, x)
    assert_bit_exact(got[0], ref[0], name)
    assert eng.kernel_launches >= 1


def test_config2_one_pole_bit_exact(oracle_mod):
    x = S.sweep(8, 128 * 64)
    got, ref, _ = run_both(oracle_mod, S.config2(one_pole=True), x)
    assert_bit_exact(got[0], ref[0], "one-pole cascade")


def test_target_chain_with_direct_fir_bit_exact(oracle_mod):
    spec = S.target_chain(n_taps=512)
    x = S.noise(6, 128 * 110)
    got, ref, _ = run_both(oracle_mod, spec, x, chunks=[128 * 10, 128 * 100])
    assert_bit_exact(got[0], ref[0], "target chain (direct FIR)")


def test_config5_full_graph(oracle_mod):
    spec = S.config5(n_taps=256)
    x = S.noise(9, 128 * 100)
    got, ref, eng = run_both(oracle_mod, spec, x)
    # path A has a Tanh distortion (libm vs CUDA tanhf) so the graph is compared within tolerance
    assert_audio_close(got[0], ref[0], what="config5 graph")
    assert "fir step" in eng.describe_plan()


def test_block_size_invariance(oracle_mod):
    spec = S.config3()
    x = S.noise(4, 128 * 96)
    a, ref, _ = run_both(oracle_mod, spec, x, chunks=[128] * 96)
    b, _, _ = run_both(oracle_mod, spec, x, chunks=[1024] * 12)
    c, _, _ = run_both(oracle_mod, spec, x, chunks=[128 * 96])
    assert_bit_exact(a[0], b[0], "128 vs 1024 blocks")
    assert_bit_exact(a[0], c[0], "128 blocks vs one call")
    assert_bit_exact(a[0], ref[0], "vs oracle")


@pytest.mark.parametrize("seconds,granule,D", [(0.25, 1024, 12288), (0.25, 1, 12000), (0.5, 1024, 24576),
                                               (0.125, 1024, 6144), (0.001, 1, 128), (0.3333, 1, 15998)])
def test_reverb_index_work_bit_exact(oracle_mod, seconds, granule, D):
    spec = single("reverb", seconds=seconds, decay=0.5)
    n = 128 * (3 * D // 128 + 2)
    x = S.impulse(2, n)
    x[1] = S.noise(1, n)[0]
    got, ref, eng = run_both(oracle_mod, spec, x, chunks=[128 * 3, n - 128 * 3], ring_granule=granule)
    assert eng.get_i64(0, "delay_samples") == D
    assert eng.get_i64(0, "ring_pos") == n % D
    assert list(np.flatnonzero(got[0][0])) == [k * D for k in range(n // D + (1 if n % D else 0)) if k * D < n]
    assert_bit_exact(got[0], ref[0], "comb")


def test_fresh_reverb_uses_make_buffer_ring(oracle_mod):
    g = GraphSpec().node(0, "reverb").node(10, "input").node(11, "output")
    g.link(10, "out", 0, "in").link(0, "out", 11, "in")
    x = S.noise(2, 128 * 40)
    got, ref, eng = run_both(oracle_mod, g, x)
    assert eng.get_i64(0, "delay_samples") == 1024     # circular_buffer(128) rounded to the page granule
    assert_bit_exact(got[0], ref[0])


def test_fan_in_fan_out_mix_add_mux_demux(oracle_mod):
    g = GraphSpec().node(10, "input").node(12, "input").node(11, "output").node(13, "output")
    g.node(0, "gain", level=2.0).node(1, "mix", ratio=0.25).node(2, "add").node(3, "demux", out_port="B").node(4, "mux", in_port="B")
    g.link(10, "out", 0, "in").link(10, "out", 1, "a").link(12, "out", 1, "b").link(0, "out", 2, "a").link(1, "out", 2, "b")
    g.link(2, "out", 3, "in").link(3, "a", 4, "a").link(3, "b", 4, "b").link(4, "out", 11, "in").link(0, "out", 11, "in")
    g.link(12, "out", 11, "in").link(3, "a", 13, "in")
    a, b = S.noise(3, 1024), S.noise(3, 1024, seed=5)
    got, ref, _ = run_both(oracle_mod, g, [a, b])
    assert_bit_exact(got[0], ref[0], "three-link fan-in sink")
    assert_bit_exact(got[1], ref[1], "demux unselected port")
    assert not got[1].any()


def test_modulated_parameters(oracle_mod):
    g = GraphSpec().node(10, "input").node(12, "input").node(11, "output")
    g.node(0, "gain").node(1, "distort", mode="HardClip").node(2, "mix")
    g.link(10, "out", 0, "in").link(12, "out", 0, "level").link(0, "out", 1, "in").link(12, "out", 1, "level")
    g.link(1, "out", 2, "a").link(10, "out", 2, "b").link(12, "out", 2, "ratio").link(2, "out", 11, "in")
    a, c = S.noise(3, 1024), S.sweep(3, 1024) * 2.5
    got, ref, _ = run_both(oracle_mod, g, [a, c])
    assert_bit_exact(got[0], ref[0], "control ports (lib.rs:122-161)")


def test_param_change_semantics(oracle_mod):
    from dsp_stuff_b200.engine import Engine

    spec = S.config3()
    x = S.noise(4, 2048)
    o = make_oracle(oracle_mod, spec, 4)
    e = make_engine(spec, 4, 2048)
    assert_bit_exact(e.process(x)[0], o.process(x)[0])
    for eng in (o, e):
        eng.set_f32(0, "level", 0.5)          # plain store
        eng.set_f32(2, "b0", 0.3)             # regenerate_filter: coefficients + state reset
        eng.set_f32(3, "decay", 0.25)         # refresh_seconds: ring replaced by zeros
    assert_bit_exact(e.process(x)[0], o.process(x)[0], "after setters")
    o.reset_state(); e.reset_state()
    assert_bit_exact(e.process(x)[0], o.process(x)[0], "after reset")


def test_errors_surface_as_status_codes():
    from dsp_stuff_b200.engine import Engine, EngineError

    e = Engine(2)
    with pytest.raises(EngineError) as ei:
        e.add_node("muff", 0)
    assert ei.value.code == -2
    e.add_node("gain", 0)
    with pytest.raises(EngineError) as ei:
        e.set_f32(0, "nope", 1.0)
    assert ei.value.code == -3
    e.add_node("gain", 1)
    e.link(0, "out", 1, "in"); e.link(1, "out", 0, "in")
    with pytest.raises(EngineError) as ei:
        e.compile()
    assert ei.value.code == -4


def test_graph_json_loader_matches_builder(oracle_mod):
    from dsp_stuff_b200.engine import Engine

    spec = S.config3()
    x = S.noise(4, 4096)
    ref = make_oracle(oracle_mod, spec, 4).process(x)[0]
    e = Engine(4, max_samples=4096)
    e.load_graph_json(spec.to_json())
    assert_bit_exact(e.process(x)[0], ref, "graph restored from the reference's JSON format")


def test_device_pointer_path_and_full_size_subset(oracle_mod):
    """BASELINE config 3 at full width (1024 channels x 1024-sample blocks): device-resident call;
    a subset of channels is checked against the oracle (channels are independent)."""
    import torch

    C, n = 1024, 1024 * 16
    spec = S.config3()
    x = S.noise(C, n)
    e = make_engine(spec, C, n)
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    e.process_device([xd], [yd], n)
    torch.cuda.synchronize()
    y = yd.cpu().numpy()
    sel = [0, 1, 17, 511, 1023]
    o = make_oracle(oracle_mod, spec, len(sel))
    assert_bit_exact(y[sel], o.process(x[sel])[0], "config3 full width, channel subset")
    # idempotence of reset: same input after reset gives the same output
    e.reset_state()
    yd2 = torch.empty_like(xd)
    e.process_device([xd], [yd2], n)
    torch.cuda.synchronize()
    assert torch.equal(yd, yd2)


def _siggen_graph(mode, amp=0.5, freq=100.0):
    g = GraphSpec().node(0, "signal_gen", mode=mode, amplitude=amp, frequency=freq).node(1, "gain", level=1.5).node(11, "output")
    return g.link(0, "out", 1, "in").link(1, "out", 11, "in")


@pytest.mark.parametrize("mode,exact", [("Triangle", True), ("Square", True), ("Constant", True), ("Sine", False)])
def test_signal_gen_source(oracle_mod, mode, exact):
    """SignalGen as an in-graph source, 128-sample block phase clock (signal_gen.rs:55-109)."""
    from dsp_stuff_b200.engine import Engine

    C, n = 3, 128 * 75
    spec = _siggen_graph(mode, freq=997.0)
    o = make_oracle(oracle_mod, spec, C)
    e = Engine(C, max_samples=n)
    spec.apply(e)
    got = np.concatenate([e.process([], 128 * 10)[0], e.process([], n - 128 * 10)[0]], axis=1)
    ref = o.process_n(n)[0]
    if exact:
        assert_bit_exact(got, ref, f"signal_gen {mode}")
    else:
        assert_audio_close(got, ref, what=f"signal_gen {mode}")


def test_signal_gen_modulated_by_lfo(oracle_mod):
    from dsp_stuff_b200.engine import Engine

    g = GraphSpec().node(0, "signal_gen", mode="Triangle", frequency=3.0, amplitude=1.0)      # LFO
    g.node(1, "signal_gen", mode="Triangle", amplitude=0.8, frequency=440.0).node(2, "biquad").node(11, "output")
    g.link(0, "out", 1, "frequency").link(0, "out", 1, "amplitude").link(1, "out", 2, "in").link(2, "out", 11, "in")
    C, n = 2, 128 * 70
    o = make_oracle(oracle_mod, g, C)
    e = Engine(C, max_samples=n)
    g.apply(e)
    assert_bit_exact(e.process([], n)[0], o.process_n(n)[0], "LFO-modulated signal_gen into a biquad")


def test_node_process_matches_simple_node_contract(oracle_mod):
    """dspb_node_process = one SimpleNode::process per 128-block on PRE-AVERAGED port buffers (node.rs:135-146,
    217-251): no fan-in division, `None` = unconnected port, node state carried across calls."""
    from dsp_stuff_b200.engine import Engine

    C, n = 4, 128 * 20
    x, ctl, b = S.noise(C, n), S.sweep(C, n), S.noise(C, n, seed=9)
    e = Engine(C, max_samples=n)
    o = oracle_mod.Oracle(C)
    for eng in (e, o):
        eng.add_node("gain", 0); eng.set_f32(0, "level", 3.0)
        eng.add_node("mix", 1); eng.set_f32(1, "ratio", 0.3)
        eng.add_node("demux", 2); eng.set_enum(2, "out_port", "B")
        eng.add_node("biquad", 3)
        eng.add_node("fir", 4); eng.set_taps(4, S.reverb_ir(200)[::-1].copy())
        eng.add_node("reverb", 5); eng.set_f32(5, "seconds", 0.01)
    assert_bit_exact(e.node_process(0, [x, None])[0], o.node_process(0, [x, None])[0], "gain, slider level")
    assert_bit_exact(e.node_process(0, [x, ctl])[0], o.node_process(0, [x, ctl])[0], "gain, level control port")
    assert_bit_exact(e.node_process(1, [x, b, None])[0], o.node_process(1, [x, b, None])[0], "mix")
    assert_bit_exact(e.node_process(1, [x, None, ctl])[0], o.node_process(1, [x, None, ctl])[0], "mix, b unconnected")
    ga, gb = e.node_process(2, [x], n_outputs=2)
    ra, rb = o.node_process(2, [x], n_outputs=2)
    assert_bit_exact(ga, ra); assert_bit_exact(gb, rb); assert not ga.any()
    for k in range(2):   # state carries across calls
        assert_bit_exact(e.node_process(3, [x])[0], o.node_process(3, [x])[0], f"biquad call {k}")
        assert_bit_exact(e.node_process(5, [x])[0], o.node_process(5, [x])[0], f"reverb call {k}")
    e2 = Engine(C, max_samples=n, fir_mode=FIR_DIRECT)
    e2.add_node("fir", 4); e2.set_taps(4, S.reverb_ir(200)[::-1].copy())
    assert_bit_exact(e2.node_process(4, [x])[0], o.node_process(4, [x])[0], "fir (direct)")
    o.reset_state()   # the oracle's fir already consumed one call above; compare the FFT path from a fresh state
    assert_audio_close(e.node_process(4, [x])[0], o.node_process(4, [x])[0], what="fir (fft)")


# ---- launch shapes at BASELINE widths -----------------------------------------------------------------------------
# The fused kernels pick their CTA geometry from the channel count of the launch (8-channel CTAs in one wave, the
# exclusive-R layout for <= 128 CTAs, two pipelined recurrence warps, shared-memory vregs next to the pipeline, the
# plain kernel above those limits).  Small-channel tests never reach most of these shapes, so each is run here at a
# width that selects it, device-resident (one launch over all channels), and a channel subset is checked against the
# oracle -- channels are independent.
def _full_width(oracle_mod, spec, C, n, exact=True, fir_mode=FIR_DIRECT, calls=2):
    import torch

    x = S.noise(C, n * calls)
    e = make_engine(spec, C, n, fir_mode=fir_mode)
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    for k in range(calls):  # state carried across calls
        xin = xd[:, k * n:(k + 1) * n].contiguous()
        yout = torch.empty_like(xin)
        e.process_device([xin], [yout], n)
        yd[:, k * n:(k + 1) * n] = yout
    torch.cuda.synchronize()
    y = yd.cpu().numpy()
    sel = sorted({0, 1, 7, 8, 13, 14, 15, 16, 255 % C, C // 2 + 3, C - 2, C - 1})
    o = make_oracle(oracle_mod, spec, len(sel))
    ref = np.concatenate([o.process(x[sel][:, k * n:(k + 1) * n])[0] for k in range(calls)], axis=1)
    if exact:
        assert_bit_exact(y[sel], ref, f"C={C}")
    else:
        assert_audio_close(y[sel], ref, what=f"C={C}")
    return e


@pytest.mark.parametrize("name,C,n", [
    ("config3", 4096, 128 * 24),           # one recurrence, 8-channel CTAs (512 CTAs, four per SM)
    ("config3", 2048, 128 * 24),           # exclusive-R layout, G = 16
    ("config2", 256, 128 * 40),            # two pipelined recurrence warps, G = 2 (BASELINE config 2 width)
    ("config2", 1024, 128 * 24),           # same, G = 8
    ("config2", 4096, 128 * 16),           # above 128 CTAs: the plain kernel, static biquad-biquad chain
    ("config2_one_pole", 256, 128 * 40),   # low_pass -> high_pass on two recurrence warps (x kept for y = x - z)
    ("config2_one_pole", 4096, 128 * 16),
    ("config1", 4096, 128 * 24),           # no recurrence
])
def test_full_width_launch_shapes_bit_exact(oracle_mod, name, C, n):
    spec = S.WORKLOADS[name][0]()
    _full_width(oracle_mod, spec, C, n, exact=True)


def test_full_width_config5_graph(oracle_mod):
    """BASELINE config 5 per-GPU width (1024 channels): the 17-op segment with shared-memory vregs runs in the
    warp-specialised kernel (G = 4); Tanh on path A -> tolerance."""
    e = _full_width(oracle_mod, S.config5(n_taps=256), 1024, 128 * 24, exact=False)
    assert "G=4" in e.describe_plan()


def test_full_width_target_chain(oracle_mod):
    """north_star target width (4096 channels), short taps through the exact FIR path: bit-exact end to end."""
    _full_width(oracle_mod, S.target_chain(n_taps=128), 4096, 128 * 16, exact=True)
