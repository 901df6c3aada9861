"""The host scheduler's output, executed: tests/plan_emulator.py walks the lowered plan of a graph (planning mode, no GPU)
with the numpy node arithmetic and must reproduce the oracle's run of the ORIGINAL graph bit for bit -- across calls, so
state slots and rings are bound to the right nodes too.  Covers what a listing check cannot: the value in a reused
shared-memory slot, the fan-in divisor folded into the right consumer, cut values read back from the right scratch."""
import numpy as np
import pytest

from dsp_stuff_b200 import GraphSpec
from dsp_stuff_b200 import signals as S
from oracle.np_oracle import NpOracle
from tests.plan_emulator import PlanEmulator
from tests.test_scheduler_fuzz import plan_of
from tests.util import assert_bit_exact


def run(spec, C, xs, calls=2, n=None, channels_for_plan=None):
    emu = PlanEmulator(plan_of(spec, channels_for_plan or C), spec, C)
    ref = NpOracle(C)
    spec.apply(ref)
    for call in range(calls):
        part = [x[:, call * n:(call + 1) * n] for x in xs] if xs else []
        got = emu.process(part, n)
        want = ref.process(part, n)
        assert len(got) == len(want)
        for k in range(len(want)):
            assert_bit_exact(got[k], want[k], f"sink {k}, call {call}")
    return emu


@pytest.mark.parametrize("name", ["config1", "config2", "config3", "config5", "target"])
def test_baseline_workloads(name):
    mk = S.WORKLOADS[name][0]
    spec = mk(n_taps=24) if name in ("config5", "target") else mk()
    n = 128 * 2
    run(spec, 2, [S.noise(2, 2 * n)], n=n, channels_for_plan=4096)


@pytest.mark.parametrize("seed", range(30))
def test_small_random_graphs(seed):
    n = 128 * 2
    run(S.random_graph(seed), 2, [S.noise(2, 2 * n, seed=seed + 1), S.sweep(2, 2 * n) * 1.5], n=n)


@pytest.mark.parametrize("seed,n_nodes", [(2000, 30), (2000, 43), (2001, 57), (2001, 69), (1010, 20), (1025, 35), (1044, 54)])
def test_large_random_graphs_cut_into_segments(seed, n_nodes):
    n = 128 * 2
    emu = run(S.random_graph(seed, n_nodes), 2, [S.noise(2, 2 * n, seed=seed + 1), S.sweep(2, 2 * n) * 1.5], n=n)
    assert sum(1 for st in emu.steps if st["kind"] == "fused") > sum(1 for st in emu.steps if st["kind"] == "fir") + 1


def test_modulated_parameters_and_generators():
    """control ports (range-mapped tiles feeding gain / distort / mix / signal_gen) and a source with no inputs"""
    g = (GraphSpec().node(100, "input").node(101, "output").node(102, "output")
         .node(0, "signal_gen", mode="Triangle", frequency=3.0, amplitude=0.8)
         .node(1, "gain", level=2.0).node(2, "distort", mode="HardClip", level=4.0).node(3, "mix", ratio=0.3)
         .node(4, "signal_gen", mode="Square", frequency=500.0)
         .link(100, "out", 1, "in").link(0, "out", 1, "level")
         .link(1, "out", 2, "in").link(0, "out", 2, "level").link(100, "out", 2, "level")
         .link(2, "out", 3, "a").link(100, "out", 3, "b").link(0, "out", 3, "ratio")
         .link(0, "out", 4, "frequency").link(4, "out", 102, "in").link(3, "out", 101, "in"))
    n = 128 * 3
    run(g, 3, [S.noise(3, 2 * n)], n=n)


def _golden_graphs():
    import glob
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    return sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(here, "golden", "*.json"))
                  if p.endswith("_graph.json") or os.path.basename(p).startswith("ref_shape_"))


@pytest.mark.parametrize("name", _golden_graphs())
def test_saved_graphs_through_the_cpp_loader(name):
    """tests/golden/*.json -> dspb_load_graph_json (C++) -> lowered plan -> executed, against the oracle fed by the Python
    parser of the same text: the C++ loader's reading of every parameter, enum, tap and link, without a GPU."""
    import os
    from dsp_stuff_b200.engine import Engine

    text = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".json")).read()
    spec = GraphSpec.from_json(text)
    C, n = 2, 128 * 2
    e = Engine(C, block=128, max_samples=n, device=-1)
    e.load_graph_json(text)
    emu = PlanEmulator(e.describe_plan(), spec, C)
    ref = NpOracle(C)
    spec.apply(ref)
    n_in = len(ref.in_terms)
    for call in range(2):
        xs = [S.noise(C, n, seed=7 + call + 10 * k) for k in range(n_in)]
        got, want = emu.process(xs, n), ref.process(xs, n)
        assert len(got) == len(want) > 0
        for k in range(len(want)):
            assert_bit_exact(got[k], want[k], f"{name}: sink {k}, call {call}")
