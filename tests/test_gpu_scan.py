"""Opt-in time-parallel recurrences (dspb_config::iir_mode = 1, north_star's "warp-shuffle associative scan over the block").

Default (iir_mode 0) stays bit-exact -- every other GPU test covers that.  Here: scan mode is enabled per filter only when the
error measured at compile time on the device probe stays below 5e-6, the verdict is visible in the plan, a filter that
fails the probe keeps the exact path, and whatever path runs stays inside the float-audio tolerance against the oracle,
across calls (state carried) and across the CTA geometries (one warp ... sixteen warps per channel)."""
import numpy as np
import pytest

from dsp_stuff_b200 import GraphSpec
from dsp_stuff_b200 import signals as S
from tests.test_oracle_kat import single
from tests.util import assert_audio_close, assert_bit_exact, make_oracle

pytestmark = pytest.mark.gpu


def engine(spec, C, n, iir_mode=1):
    from dsp_stuff_b200.engine import Engine

    e = Engine(C, block=128, max_samples=n, fir_mode=1, iir_mode=iir_mode)
    spec.apply(e)
    return e


def run(oracle_mod, spec, C, n, calls=2, sweep=False):
    x = (S.sweep if sweep else S.noise)(C, n * calls)
    e = engine(spec, C, n)
    got = np.concatenate([e.process(x[:, k * n:(k + 1) * n])[0] for k in range(calls)], axis=1)
    sel = sorted({0, 1, C // 2, C - 1})
    o = make_oracle(oracle_mod, spec, len(sel))
    ref = np.concatenate([o.process(x[sel][:, k * n:(k + 1) * n])[0] for k in range(calls)], axis=1)
    return got[sel], ref, e


@pytest.mark.parametrize("typename,params,expect_scan", [
    ("biquad", S.rbj_biquad("lp", 1000.0), True),
    ("biquad", dict(), True),                               # the default filter: y = 0.758 x + 0.24 y1
    ("low_pass", dict(ratio=0.9), True),
    ("high_pass", dict(ratio=0.9), True),
    ("biquad", S.rbj_biquad("hp", 30.0, q=8.0), None),       # poles next to the unit circle: whatever the probe says
])
@pytest.mark.parametrize("C", [3, 300, 5000])   # G = 1 (16 warps per channel), 2, 32 (half a warp per channel)
def test_single_filter_scan_within_tolerance(oracle_mod, typename, params, expect_scan, C):
    n = 128 * 33   # one full 4096-sample tile + a partial one at G = 1
    got, ref, e = run(oracle_mod, single(typename, **params), C, n, calls=3)
    plan = e.describe_plan()
    if expect_scan is True:
        assert "time-parallel scan (probe error" in plan, plan
    assert ("time-parallel scan" in plan) != ("exact, lane=channel" in plan), plan
    rel, dbfs = assert_audio_close(got, ref, what=f"{typename} {params} C={C}")
    print(f"{typename} C={C}: {'scan' if 'time-parallel' in plan else 'exact'} rel {rel:.2e} dbfs {dbfs:.1f}")


def test_filter_that_fails_the_probe_stays_exact(oracle_mod):
    """A marginally stable resonator amplifies rounding differences far beyond the gate: it must keep the exact path and
    then match the oracle bit for bit."""
    spec = single("biquad", a0=1.0, a1=-1.99990, a2=0.99995, b0=1e-4, b1=0.0, b2=0.0)
    got, ref, e = run(oracle_mod, spec, 4, 128 * 40, calls=2)
    plan = e.describe_plan()
    assert "exact, lane=channel (scan probe error" in plan, plan
    assert_bit_exact(got, ref, "exact fallback")


@pytest.mark.parametrize("C", [256, 1024, 4096])
def test_config3_chain_in_scan_mode(oracle_mod, C):
    n = 128 * 110
    got, ref, e = run(oracle_mod, S.config3(), C, n, calls=2)
    assert "time-parallel scan" in e.describe_plan()
    assert_audio_close(got, ref, what=f"config3 scan C={C}")


def test_config2_all_or_nothing(oracle_mod):
    """biquad LP 1 kHz -> biquad HP 200 Hz: if the high-pass fails the probe, the low-pass goes back to exact too (one
    sequential chain bounds the segment anyway) and the chain is bit-exact; if both pass, tolerance."""
    got, ref, e = run(oracle_mod, S.config2(), 256, 128 * 64, calls=2)
    plan = e.describe_plan()
    print(plan)
    if "exact after all" in plan or "time-parallel scan" not in plan:
        assert_bit_exact(got, ref, "config2 exact")
    else:
        assert_audio_close(got, ref, what="config2 scan")


def test_one_pole_cascade_scan_and_sweep_input(oracle_mod):
    got, ref, e = run(oracle_mod, S.config2(one_pole=True), 256, 128 * 64, calls=2, sweep=True)
    assert_audio_close(got, ref, what="one-pole cascade")


def test_default_mode_is_unchanged(oracle_mod):
    spec = S.config3()
    x = S.noise(8, 128 * 40)
    e = engine(spec, 8, 128 * 40, iir_mode=0)
    assert "time-parallel" not in e.describe_plan()
    assert_bit_exact(e.process(x)[0], make_oracle(oracle_mod, spec, 8).process(x)[0], "iir_mode 0")
