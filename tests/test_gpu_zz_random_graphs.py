"""GPU parity on random graphs, including graphs too large for one fused segment (the scheduler cuts them and routes the cut
values through scratch: csrc/engine.cpp lower_graph).  Named to run after every other GPU file.  Every node type the
generator draws (dsp_stuff_b200.signals.random_graph) is FMA-free and the FIR runs on the exact f64 path, so the comparison
with the oracle is bit for bit; state is carried across three calls of different lengths.  The structure of the same plans is
checked without a GPU in tests/test_scheduler_fuzz.py."""
import numpy as np
import pytest

from dsp_stuff_b200 import GraphSpec
from dsp_stuff_b200 import signals as S
from tests.test_gpu_parity import run_both
from tests.util import assert_bit_exact

pytestmark = pytest.mark.gpu
CHUNKS = [128 * 2, 128 * 6, 128 * 1]


def _inputs(C, seed):
    n = sum(CHUNKS)
    return [S.noise(C, n, seed=seed + 1), S.sweep(C, n) * 1.5]


@pytest.mark.parametrize("seed", range(12))
def test_small_random_graph_bit_exact(oracle_mod, seed):
    got, ref, _ = run_both(oracle_mod, S.random_graph(seed), _inputs(5, seed), chunks=CHUNKS)
    for k in range(2):
        assert_bit_exact(got[k], ref[k], f"random graph {seed}, sink {k}")


@pytest.mark.parametrize("seed,n_nodes", [(2000, 30), (2000, 43), (2001, 57), (2001, 69)])
def test_large_random_graph_cut_into_segments_bit_exact(oracle_mod, seed, n_nodes):
    got, ref, eng = run_both(oracle_mod, S.random_graph(seed, n_nodes), _inputs(3, seed), chunks=CHUNKS)
    plan = eng.describe_plan()
    assert plan.count("fused segment:") > plan.count("fir step:") + 1      # at least one cut no FIR node forced
    for k in range(2):
        assert_bit_exact(got[k], ref[k], f"{n_nodes}-node random graph {seed}, sink {k}")


def test_forty_biquads_in_a_row_bit_exact(oracle_mod):
    """12 state slots per Program: the chain runs as four segments, one scratch value between neighbours."""
    g = GraphSpec().node(100, "input").node(101, "output")
    prev = 100
    for i in range(40):
        g.node(i, "biquad", **S.rbj_biquad("lp" if i % 2 else "hp", 300.0 + 150 * i))
        g.link(prev, "out", i, "in")
        prev = i
    g.link(prev, "out", 101, "in")
    x = S.noise(40, sum(CHUNKS))
    got, ref, eng = run_both(oracle_mod, g, x, chunks=CHUNKS)
    assert eng.describe_plan().count("fused segment:") >= 4
    assert_bit_exact(got[0], ref[0], "40 biquads")


def test_fifteen_reverbs_in_a_row_bit_exact(oracle_mod):
    """6 comb rings per Program: three segments."""
    g = GraphSpec().node(100, "input").node(101, "output")
    prev = 100
    for i in range(15):
        g.node(i, "reverb", seconds=0.01 + 0.004 * i, decay=0.4)
        g.link(prev, "out", i, "in")
        prev = i
    g.link(prev, "out", 101, "in")
    x = S.noise(7, 128 * 40)
    got, ref, eng = run_both(oracle_mod, g, x, chunks=[128 * 9, 128 * 31])
    assert eng.describe_plan().count("fused segment:") == 3
    assert_bit_exact(got[0], ref[0], "15 reverbs")
    assert np.count_nonzero(ref[0]) > 0


def test_ten_one_pole_filters_in_scan_mode(oracle_mod):
    """iir_mode = 1: four scan tables per Program, so the chain runs as three all-time-parallel segments (it used to fall
    back to ten sequential recurrences).  Scan mode is the tolerance contract, not bit-exactness."""
    from dsp_stuff_b200.engine import Engine
    from tests.util import assert_audio_close, make_oracle

    g = GraphSpec().node(100, "input").node(101, "output")
    prev = 100
    for i in range(10):
        g.node(i, "low_pass" if i % 2 == 0 else "high_pass", ratio=0.5 + 0.04 * i)
        g.link(prev, "out", i, "in")
        prev = i
    g.link(prev, "out", 101, "in")
    C, n = 64, 128 * 40
    x = S.noise(C, 2 * n)
    e = Engine(C, block=128, max_samples=n, iir_mode=1)
    g.apply(e)
    plan = e.describe_plan()
    o = make_oracle(oracle_mod, g, C)
    for call in range(2):
        got = e.process(x[:, call * n:(call + 1) * n])[0]
        ref = o.process(x[:, call * n:(call + 1) * n])[0]
        if "time-parallel scan" in plan and "exact" not in plan:   # every filter passed the device probe (expected: ~2e-7)
            assert plan.count("fused segment:") == 3
        assert_audio_close(got, ref, what=f"10 one-pole filters, iir_mode 1, call {call}")
