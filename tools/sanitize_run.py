#!/usr/bin/env python
"""Small invocations of every fused-kernel launch shape, for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool racecheck python tools/sanitize_run.py
Covers the warp-specialised pipeline (named barriers), the exclusive-R layout, the two-recurrence pipeline, the
high_pass keep buffers, shared-memory vregs next to the pipeline, the plain kernel, and the FIR kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from dsp_stuff_b200 import signals as S  # noqa: E402
from dsp_stuff_b200.engine import Engine  # noqa: E402

CASES = [
    ("config3", S.config3(), 24, 128 * 6, 0),
    ("config2", S.config2(), 12, 128 * 40, 0),
    ("config2_one_pole", S.config2(one_pole=True), 12, 128 * 40, 0),
    ("config5", S.config5(n_taps=64), 10, 128 * 10, 0),
    ("target fft (narrow kernel, single segments)", S.target_chain(n_taps=4096), 4, 128 * 72, 0, 0),
    ("target fft (wide kernel: one double segment + a single one)", S.target_chain(n_taps=4096), 6, 128 * 104, 0, 0),
    ("target toeplitz", S.target_chain(n_taps=300), 4, 128 * 8, 2, 0),
    ("config3 scan mode, 16 warps per channel", S.config3(), 3, 128 * 40, 0, 1),
    ("config3 scan mode, one warp per channel", S.config3(), 5000, 128 * 3, 0, 1),
    ("one-pole cascade scan mode", S.config2(one_pole=True), 300, 128 * 12, 0, 1),
]
CASES = [c if len(c) == 6 else c + (0,) for c in CASES]
for name, spec, C, n, fir_mode, iir_mode in CASES:
    e = Engine(C, block=128, max_samples=n, fir_mode=fir_mode, iir_mode=iir_mode)
    spec.apply(e)
    x = S.noise(C, n)
    y = e.process(x)[0]
    y2 = e.process(x)[0]  # second call: carried state
    assert np.isfinite(y).all() and np.isfinite(y2).all(), name
    print(f"{name}: ok ({e.kernel_launches} launches per call)")
e = Engine(3, block=128, max_samples=128)
out, used = e.resample_dup_stereo(S.noise(3, 2048), 1800, 44100.0)
assert np.isfinite(out).all()
print(f"resample_dup_stereo: ok (consumed {used})")
print("sanitize_run done")
