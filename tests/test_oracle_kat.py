"""Known-answer tests that pin the CPU oracle to the reference SOURCE TEXT (the reference has no tests
of its own: SURVEY.md §4).  Every expected value here is derived by hand from the cited lines."""
import numpy as np
import pytest

from dsp_stuff_b200 import GraphSpec
from dsp_stuff_b200 import signals as S
from tests.util import NF1, NF2, NF3, assert_bit_exact, make_oracle

f32 = np.float32


def single(typename, **params):
    g = GraphSpec().node(0, typename, **params).node(10, "input").node(11, "output")
    ins, outs = __import__("dsp_stuff_b200").NODE_PORTS[typename]
    return g.link(10, "out", 0, ins[0]).link(0, outs[0], 11, "in")


def test_fan_in_divisors_bit_patterns():
    # node.rs:166,179: num_frames = 0.0001f32, += 1.0 per delivering link
    assert NF1.view(np.uint32) == 0x3F800347
    assert NF2.view(np.uint32) == 0x400001A4
    assert NF3.view(np.uint32) == 0x404001A4


def test_gain_unity_is_two_divisions(oracle_mod):
    x = S.noise(2, 256)
    y = make_oracle(oracle_mod, single("gain"), 2).process(x)[0]   # level default 1.0 (gain.rs:21)
    expect = ((f32(0.0) + x) / NF1 * f32(1.0)) / NF1                # gain input port + Output terminal
    assert_bit_exact(y, expect)


def test_distort_default_level_is_bypass(oracle_mod):
    x = S.noise(1, 128)
    y = make_oracle(oracle_mod, single("distort"), 1).process(x)[0]  # level 0.0 < 0.001 (distort.rs:46,72)
    assert_bit_exact(y, (x / NF1) / NF1)


def test_distort_softclip_formula(oracle_mod):
    x = S.noise(1, 128)
    y = make_oracle(oracle_mod, single("distort", level=4.0, mode="SoftClip"), 1).process(x)[0]
    s = (x / NF1) * f32(4.0)
    mid = s - ((s * s) * s) / f32(3.0)
    shaped = np.where(s > 1, f32(2.0) / f32(3.0), np.where(s >= -1, mid, f32(-2.0) / f32(3.0))).astype(f32)
    expect = (np.clip(shaped, -1, 1) / f32(4.0)) / NF1
    assert_bit_exact(y, expect)


def test_fuzz_zero_block_is_nan_and_output_nonpositive(oracle_mod):
    o = make_oracle(oracle_mod, single("distort", level=4.0, mode="Fuzz"), 1)
    x = np.concatenate([np.zeros((1, 128), f32), S.noise(1, 128)], axis=1)
    y = o.process(x)[0]
    assert np.all(np.isnan(y[0, :128]))          # clip(0)/0 (distort.rs:158)
    assert np.all(y[0, 128:] <= 0)               # sign discarded by copysign(-1) (distort.rs:159)
    assert np.max(np.abs(y[0, 128:] * NF1)) == pytest.approx(float(np.max(np.abs(x[0, 128:] / NF1))), rel=1e-6)


def test_biquad_defaults_one_pole(oracle_mod):
    x = S.impulse(1, 128)
    y = make_oracle(oracle_mod, single("biquad"), 1).process(x)[0][0] * NF1
    # defaults: y = 0.758 x + 0.24 y1 (biquad.rs:49-55); impulse amplitude 1/nf
    e = np.zeros(128, f32)
    acc = f32(0)
    xin = (x[0] / NF1).astype(f32)
    for i in range(128):
        acc = f32(f32(f32(f32(f32(0.758) * xin[i]) + f32(0)) + f32(0)) - f32(f32(-0.24) * acc)) - f32(0)
        e[i] = acc
    assert_bit_exact(y, (e / NF1) * NF1)


def test_biquad_param_change_resets_state(oracle_mod):
    o = make_oracle(oracle_mod, single("biquad"), 1)
    x = S.noise(1, 128)
    a = o.process(x)[0]
    o.set_f32(0, "b0", 0.758)   # any slider change -> regenerate_filter -> reset_state (biquad.rs:74)
    b = o.process(x)[0]
    assert_bit_exact(a, b)


def test_one_pole_formulas(oracle_mod):
    x = S.noise(1, 256)
    xin = (x[0] / NF1).astype(f32)
    r = f32(0.9)
    lp = make_oracle(oracle_mod, single("low_pass", ratio=0.9), 1).process(x)[0][0]
    hp = make_oracle(oracle_mod, single("high_pass", ratio=0.9), 1).process(x)[0][0]
    z = f32(0)
    zl = f32(0)
    el, eh = np.zeros(256, f32), np.zeros(256, f32)
    for i in range(256):
        zl = f32(f32(xin[i] * f32(f32(1) - r)) + f32(r * zl)); el[i] = zl           # low_pass.rs:37-38
        z = f32(f32(xin[i] * f32(f32(1) - r)) + f32(r * z)); eh[i] = f32(xin[i] - z)  # high_pass.rs:37-38
    assert_bit_exact(lp, el / NF1)
    assert_bit_exact(hp, eh / NF1)


@pytest.mark.parametrize("seconds,granule,D", [(0.25, 1024, 12288), (0.25, 1, 12000), (0.5, 1024, 24576),
                                               (0.125, 1024, 6144), (0.001, 1024, 1024), (0.001, 1, 128),
                                               (1.0, 1024, 48128)])
def test_reverb_delay_length(oracle_mod, seconds, granule, D):
    o = oracle_mod.Oracle(1, ring_granule=granule)
    o.add_node("reverb", 0)
    o.set_f32(0, "seconds", seconds)   # refresh_seconds: max((s*48000.0) as usize,128) (reverb.rs:58)
    assert o.get_i64(0, "delay_samples") == D


def test_reverb_fresh_node_ring(oracle_mod):
    o = oracle_mod.Oracle(1)
    o.add_node("reverb", 0)            # make_buffer(): circular_buffer(128) (reverb.rs:44-52)
    assert o.get_i64(0, "delay_samples") == 1024
    o = oracle_mod.Oracle(1, ring_granule=1)
    o.add_node("reverb", 0)
    assert o.get_i64(0, "delay_samples") == 128


def test_reverb_impulse_train(oracle_mod):
    D = 1024
    o = make_oracle(oracle_mod, single("reverb", seconds=0.001, decay=0.5), 1)
    n = 4 * D + 128
    y = o.process(S.impulse(1, n))[0][0]
    nz = np.flatnonzero(y)
    assert list(nz) == [0, D, 2 * D, 3 * D, 4 * D]            # y[n] = x[n] + decay*y[n-D] (reverb.rs:87-103)
    a = f32(1.0) / NF1
    for k in range(5):
        assert y[k * D] == (a * f32(0.5) ** k) / NF1


def test_fir_default_is_identity_and_warmup_prefix_sum(oracle_mod):
    x = S.noise(1, 256)
    y = make_oracle(oracle_mod, single("fir"), 1).process(x)[0]     # taps [1.0] (fir.rs:61)
    assert_bit_exact(y, (x / NF1) / NF1)
    taps = np.array([0.5, -0.25, 2.0, 1.0, 3.0], dtype=np.float64)   # stored reversed: taps[i] = h[N-1-i]
    g = single("fir", mode="Balanced")
    g.nodes[0].taps = taps
    y = make_oracle(oracle_mod, g, 1).process(x)[0][0]
    xin = (x[0] / NF1).astype(np.float64)
    N = len(taps)
    e = np.zeros(256, f32)
    for n in range(256):
        if n < N - 1:   # warm-up: oldest sample pairs with taps[0] (fir.rs:192-216)
            e[n] = f32(np.sum(xin[: n + 1] * taps[: n + 1]))
        else:
            e[n] = f32(np.sum(xin[n - N + 1: n + 1] * taps))
    np.testing.assert_allclose(y, e / NF1, rtol=2e-7, atol=1e-9)
    ga = single("fir", mode="Average")
    ga.nodes[0].taps = taps
    ya = make_oracle(oracle_mod, ga, 1).process(x)[0][0]
    np.testing.assert_allclose(ya, (e * (f32(1.0) / f32(N))) / NF1, rtol=2e-7, atol=1e-9)


def test_demux_unselected_zero_and_unconnected_input(oracle_mod):
    g = GraphSpec().node(0, "demux", out_port="B").node(1, "add").node(10, "input").node(11, "output").node(12, "output")
    g.link(10, "out", 0, "in").link(0, "a", 11, "in").link(0, "b", 1, "a").link(1, "out", 12, "in")
    x = S.noise(2, 128)
    ya, yb = make_oracle(oracle_mod, g, 2).process(x)
    assert not ya.any()                                   # unselected output stays zero (node.rs:272, demux.rs:50-57)
    assert_bit_exact(yb, (((x / NF1) / NF1) + f32(0.0)) / NF1)   # add.b unconnected => zeros (node.rs:162-194)


def test_fan_in_two_links_and_fan_out(oracle_mod):
    g = GraphSpec().node(0, "gain", level=2.0).node(10, "input").node(11, "output")
    g.link(10, "out", 0, "in").link(0, "out", 11, "in").link(10, "out", 11, "in")
    x = S.noise(1, 128)
    y = make_oracle(oracle_mod, g, 1).process(x)[0]
    expect = (((x / NF1) * f32(2.0)) + x) / NF2           # links summed in creation order, / 2.0001
    assert_bit_exact(y, expect)


def test_mix_and_modulated_gain(oracle_mod):
    # level control port fed by a constant 0.0 signal: level = 0 + 10*clamp((c/nf+1)/2) = 5 (lib.rs:138-146)
    g = GraphSpec().node(0, "gain").node(1, "signal_gen", mode="Constant", amplitude=0.0)
    g.node(10, "input").node(11, "output")
    g.link(10, "out", 0, "in").link(1, "out", 0, "level").link(0, "out", 11, "in")
    x = S.noise(1, 128)
    y = make_oracle(oracle_mod, g, 1).process(x)[0]
    assert_bit_exact(y, ((x / NF1) * f32(5.0)) / NF1)
    g = GraphSpec().node(0, "mix", ratio=0.25).node(10, "input").node(12, "input").node(11, "output")
    g.link(10, "out", 0, "a").link(12, "out", 0, "b").link(0, "out", 11, "in")
    a, b = S.noise(1, 128), S.noise(1, 128, seed=3)
    y = make_oracle(oracle_mod, g, 1).process([a, b])[0]
    r = f32(0.25)
    assert_bit_exact(y, (((b / NF1) * r) + ((a / NF1) * (f32(1) - r))) / NF1)   # mix.rs:45


def test_port_index_order(oracle_mod):
    o = oracle_mod.Oracle(1)
    o.add_node("overdrive", 0)
    o.add_node("mix", 1)
    o.add_node("demux", 2)
    assert [o.port_index(0, p) for p in ("in", "boost", "drive", "level")] == [0, 1, 2, 3]   # lib.rs:214-216
    assert [o.port_index(1, p) for p in ("a", "b", "ratio")] == [0, 1, 2]
    assert [o.port_index(2, p, True) for p in ("a", "b")] == [0, 1]


def test_errors(oracle_mod):
    o = oracle_mod.Oracle(1)
    with pytest.raises(oracle_mod.OracleError):
        o.add_node("muff", 0)          # GPL dependency, excluded
    o.add_node("gain", 0)
    with pytest.raises(oracle_mod.OracleError):
        o.set_f32(0, "nope", 1.0)
    o.add_node("gain", 1)
    o.link(0, "out", 1, "in")
    o.link(1, "out", 0, "in")
    with pytest.raises(oracle_mod.OracleError):
        o.compile()                    # cycle


def test_gate_extension_known_answers(oracle_mod):
    """`gate` is an EXTENSION (BASELINE north_star names a noise gate; the reference has none): hard gate keyed by
    the Envelope node's detector.  Pinned by hand-derivable answers only."""
    x = S.noise(2, 256)
    # default threshold 0.0: the envelope is >= 0 everywhere -> pass-through (two fan-in divisions)
    y = make_oracle(oracle_mod, single("gate"), 2).process(x)[0]
    assert_bit_exact(y, (x / NF1) / NF1)
    # attack = release = 0 frames: gains 0 -> the envelope is |x| itself -> per-sample gate
    y = make_oracle(oracle_mod, single("gate", threshold=0.2), 2).process(x)[0]
    xin = x / NF1
    assert_bit_exact(y, np.where(np.abs(xin) >= f32(0.2), xin, f32(0.0)) / NF1)
    # a slow release keeps the gate open after a burst: impulse 1.0 then silence, release 100 frames
    imp = np.zeros((1, 256), dtype=np.float32)
    imp[0, 0] = 1.0
    imp[0, 1:] = 1e-3
    y = make_oracle(oracle_mod, single("gate", threshold=0.5, release=100.0), 1).process(imp)[0]
    env = float(f32(1.0) / NF1)
    open_len = 1
    g = float(np.exp(f32(-1.0) / f32(100.0)))
    while env * g >= 0.5:   # e^(-k/100) decay: about 69 samples (f64 estimate, checked with a +-2 window)
        env *= g
        open_len += 1
    nz = int(np.count_nonzero(y[0]))
    assert abs(nz - open_len) <= 2, (nz, open_len)
    assert np.all(y[0, nz:] == 0.0)


# ---- playback-side converter (devices.rs:443-500, 550-556; dasp Converter + Sinc<[f32; 16]>, restated: parity unpinned) ----
def test_resampler_unity_rate_is_an_eight_frame_delay(oracle_mod):
    """target = 48 kHz: interpolation_value is 0 at every output, so only the left n = 0 tap has weight 1 (sinc(0)); the right
    taps sit at multiples of pi (sin ~ 1e-16).  Sinc::idx saturates at depth = 8, i.e. the output is frames[8] of 16: the
    input delayed by 8 frames -- derived by hand from the restated Sinc::interpolate."""
    r = oracle_mod.Resampler(2, 48000.0)
    x = S.noise(2, 256)
    y, used = r.process(x, 200)
    assert used == 199                       # the first next() pushes nothing (value starts at 0.0)
    assert np.array_equal(y[:, :, 0], y[:, :, 1])          # o.fill(x): both slots of a frame
    assert np.allclose(y[:, 16:, 0], x[:, 8:192], rtol=0, atol=1e-7)
    assert np.abs(y[:, :8, 0]).max() < 1e-7   # start-up: idx < depth, the unit-weight tap still points at the zero-filled ring


def test_resampler_44k1_tracks_a_sine_and_carries_state(oracle_mod):
    n = 4800
    x = np.sin(2 * np.pi * 1000.0 * np.arange(2 * n) / 48000.0).astype(np.float32)[None]
    r = oracle_mod.Resampler(1, 44100.0)
    y1, u1 = r.process(x[:, :n], 4000)
    y2, u2 = r.process(x[:, u1:], 4000)      # the caller re-presents what was not consumed (source.release(index))
    y = np.concatenate([y1, y2], axis=1)[0, :, 0]
    t = (np.arange(8000) * 48000.0 / 44100.0 - 8.0) / 48000.0
    assert np.max(np.abs(y[100:] - np.sin(2 * np.pi * 1000.0 * t[100:]))) < 3e-3   # 16-tap Hann-windowed sinc
    one, _ = oracle_mod.Resampler(1, 44100.0).process(x, 8000)
    assert np.array_equal(one[0, :, 0], y)   # two calls == one call
    assert abs(u1 - 4000 * 48000 / 44100) <= 2
