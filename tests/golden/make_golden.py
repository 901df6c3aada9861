#!/usr/bin/env python
"""Regenerates tests/golden/*.  The reference ships no fixtures, so these vectors are produced
by the CPU oracle (oracle/dsp_oracle.cpp) on seeded inputs; they pin the ORACLE against accidental change (CPU
test) and give the GPU parity tests a second, file-based target.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from dsp_stuff_b200 import signals as S  # noqa: E402
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    # name: (graph factory, channels, samples)
    "config1": (S.config1, 2, 128 * 110),
    "config2": (S.config2, 2, 128 * 20),
    "config2_one_pole": (lambda: S.config2(one_pole=True), 2, 128 * 20),
    "config3": (S.config3, 2, 128 * 110),
    "target_256taps": (lambda: S.target_chain(256), 2, 128 * 110),
    "config5_64taps": (lambda: S.config5(64), 2, 128 * 110),
}


def main():
    for name, (factory, C, n) in CASES.items():
        spec = factory()
        o = oracle.Oracle(C, threads=1)
        spec.apply(o)
        y = o.process(S.noise(C, n))[0]
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), y=y, channels=C, samples=n)
        open(os.path.join(HERE, f"{name}_graph.json"), "w").write(spec.to_json())
        print(name, y.shape, float(np.abs(y).max()))


if __name__ == "__main__":
    main()
