"""Independent-math cross-checks of the oracle (scipy in f64), SURVEY.md §4 row 2.

These compare the f32 oracle with IDEAL (f64) filters, so the bound is the reference arithmetic's own
rounding noise, not the GPU parity tolerance: the f32 DirectForm1 high-pass at 200 Hz sits ~1.3e-5 of
its output peak away from exact math.  That is why the CUDA engine runs recurrences bit-exactly
(DESIGN.md "IIR exactness") instead of re-ordering them."""
import numpy as np
from scipy import signal as sps

from dsp_stuff_b200 import signals as S
from tests.test_oracle_kat import single
from tests.util import NF1, assert_audio_close, make_oracle


def test_biquad_cascade_vs_lfilter(oracle_mod):
    x = S.noise(4, 12032)
    o = make_oracle(oracle_mod, S.config2(), 4)
    y = o.process(x)[0]
    lp, hp = S.rbj_biquad("lp", 1000.0), S.rbj_biquad("hp", 200.0)
    nf = float(NF1)
    v = x.astype(np.float64) / nf
    v = sps.lfilter([lp["b0"], lp["b1"], lp["b2"]], [1.0, lp["a1"], lp["a2"]], v, axis=1) / nf
    v = sps.lfilter([hp["b0"], hp["b1"], hp["b2"]], [1.0, hp["a1"], hp["a2"]], v, axis=1) / nf
    assert_audio_close(y, v, rel_tol=1e-4, what="oracle biquad cascade vs scipy f64")


def test_one_pole_vs_lfilter(oracle_mod):
    x = S.noise(2, 4096)
    y = make_oracle(oracle_mod, S.config2(one_pole=True), 2).process(x)[0]
    nf = float(NF1)
    r1, r2 = float(np.float32(0.9)), float(np.float32(0.99))
    v = x.astype(np.float64) / nf
    v = sps.lfilter([1 - r1], [1.0, -r1], v, axis=1) / nf
    z = sps.lfilter([1 - r2], [1.0, -r2], v, axis=1)
    assert_audio_close(y, (v - z) / nf, what="oracle one-pole vs scipy f64")


def test_fir_steady_state_is_convolution(oracle_mod):
    n_taps = 512
    h = S.reverb_ir(n_taps)
    x = S.noise(2, 4096)
    y = make_oracle(oracle_mod, S.config4(n_taps), 2).process(x)[0]
    nf = float(NF1)
    full = sps.fftconvolve(x.astype(np.float64) / nf, h[None, :], axes=1)[:, : x.shape[1]] / nf
    # n >= N-1: true convolution; the first N-1 samples are the warm-up prefix sums (fir.rs:192-216)
    assert_audio_close(y[:, n_taps - 1:], full[:, n_taps - 1:], what="oracle FIR steady state")
    taps = h[::-1]
    xin = x[0].astype(np.float64) / nf
    warm = np.cumsum(xin[: n_taps - 1] * taps[: n_taps - 1]) / nf
    np.testing.assert_allclose(y[0, : n_taps - 1], warm, rtol=1e-6, atol=1e-8)


def test_reverb_vs_comb_lfilter(oracle_mod):
    D = 1024
    x = S.noise(1, 8 * D)
    y = make_oracle(oracle_mod, single("reverb", seconds=0.001, decay=0.5), 1).process(x)[0]
    nf = float(NF1)
    a = np.zeros(D + 1)
    a[0], a[D] = 1.0, -0.5
    v = sps.lfilter([1.0], a, x.astype(np.float64) / nf, axis=1) / nf
    assert_audio_close(y, v, what="oracle comb vs scipy")
