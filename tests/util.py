"""Shared helpers for the parity tests.

Float-audio tolerance (BASELINE.json north_star: "within 1e-5 relative / -100 dBFS RMS error"):
  rel  = max|got - ref| / max|ref|           <= 1e-5   (peak-relative: a per-sample ratio is
                                                         meaningless at zero crossings)
  rms  = 20*log10(rms(got - ref)), FS = 1.0  <= -100 dBFS
Integer / index work and the FMA-free elementwise nodes are compared bit-exactly.
"""
import math

import numpy as np

REL_TOL = 1e-5
RMS_DBFS_TOL = -100.0


def err_metrics(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    d = got - ref
    peak = max(float(np.max(np.abs(ref))), 1e-30)
    rel = float(np.max(np.abs(d))) / peak
    rms = float(np.sqrt(np.mean(d * d)))
    dbfs = 20.0 * math.log10(rms) if rms > 0 else -math.inf
    return rel, dbfs


def assert_audio_close(got, ref, rel_tol=REL_TOL, dbfs_tol=RMS_DBFS_TOL, what=""):
    assert np.all(np.isfinite(ref)), "reference has non-finite samples; compare those cases bit-wise"
    assert np.all(np.isfinite(got)), f"{what}: non-finite samples"
    rel, dbfs = err_metrics(got, ref)
    assert rel <= rel_tol, f"{what}: peak-relative error {rel:.3e} > {rel_tol:.0e} (rms {dbfs:.1f} dBFS)"
    assert dbfs <= dbfs_tol, f"{what}: rms error {dbfs:.1f} dBFS > {dbfs_tol} dBFS (rel {rel:.3e})"
    return rel, dbfs


def assert_bit_exact(got, ref, what=""):
    got = np.asarray(got, dtype=np.float32)
    ref = np.asarray(ref, dtype=np.float32)
    same = (got == ref) | (np.isnan(got) & np.isnan(ref))  # -0.0 == +0.0
    if not np.all(same):
        idx = np.argwhere(~same)[0]
        raise AssertionError(f"{what}: {np.count_nonzero(~same)} samples differ; first at {tuple(idx)}: "
                             f"got {got[tuple(idx)]!r} ref {ref[tuple(idx)]!r}")


def make_oracle(oracle_mod, spec, channels, ring_granule=1024, threads=0):
    o = oracle_mod.Oracle(channels, ring_granule=ring_granule, threads=threads)
    spec.apply(o)
    return o


NF1 = np.float32(0.0001) + np.float32(1.0)  # collect_and_average divisor for one link (node.rs:166-179)
NF2 = NF1 + np.float32(1.0)
NF3 = NF2 + np.float32(1.0)
