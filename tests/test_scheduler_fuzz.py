"""Host scheduler (csrc/engine.cpp Lowerer) on random graphs, in planning mode (no GPU).

The scheduler turns the graph region between two FIR nodes into ONE fused segment when it fits a Program (47 ops, 12
global buffers, 12 states, 6 rings, 200 KB of shared memory) and cuts it into several otherwise; values that cross a cut
travel through global scratch.  These tests read the plan listing (`dspb_describe_plan`: the "lowered:" and "buffers:"
lines of every segment) and check the DATAFLOW of what was lowered:

* every shared-memory value read was written earlier in the same segment,
* every scratch buffer / FIR output read was written by an earlier step (or earlier in the same segment, in which case the
  read must not carry the one-tile-ahead prefetch flag); a scratch buffer is reused only by a step later than every reader
  of its previous value,
* every FIR step finds its input ring stored by the segment in front of it,
* every output terminal is written exactly once,
* every segment respects the Program limits.

The numeric side of the same graphs is tests/test_gpu_zz_random_graphs.py (GPU, against the oracle).
"""
import re

import pytest

from dsp_stuff_b200 import signals as S

OP_NOP, OP_ZERO, OP_LOADG, OP_ADDG, OP_LOADV, OP_ADDV, OP_COPYV, OP_COPYG = 1, 2, 3, 4, 5, 6, 7, 8
OP_SAVEV, OP_STOREG, OP_ADD, OP_MIX, OP_GATE = 10, 11, 17, 18, 26
READS_G = {OP_LOADG, OP_ADDG, OP_COPYG}
READS_V = {OP_LOADV, OP_ADDV, OP_COPYV, OP_ADD, OP_MIX, OP_GATE}


def parse_plan(plan):
    """-> list of steps: ("fused", ops, binds, header) or ("fir", node_id, out_term)."""
    steps = []
    for line in plan.splitlines():
        m = re.match(r"\[\d+\] fused segment: (.*)", line)
        if m:
            steps.append(["fused", None, None, m.group(1)])
            continue
        m = re.match(r"\[\d+\] fir step: fir#(\d+),(.*)", line)
        if m:
            t = re.search(r"-> output terminal (\d+)", m.group(2))
            steps.append(["fir", int(m.group(1)), int(t.group(1)) if t else None])
            continue
        if line.startswith("    lowered:"):
            ops = []
            for tok in line.split()[1:]:
                parts = tok.split(":")
                op = dict(code=int(parts[0]), v=None, g=None, pf=False, pv=[])
                for q in parts[1:]:
                    if q.startswith("v"):
                        op["v"] = int(q[1:])
                    elif q.startswith("g"):
                        op["pf"] = q.endswith("*")
                        op["g"] = int(q[1:].rstrip("*"))
                    elif q.startswith("p"):
                        op["pv"].append(int(q.split("=v")[1]))
                ops.append(op)
            steps[-1][1] = ops
        if line.startswith("    buffers:"):
            steps[-1][2] = {int(k): v for k, v in re.findall(r"g(\d+)=(\S+)", line)}
    return steps


def check_dataflow(plan, n_out_terms):
    written = set()          # scratchN / firU#id / firY#id written so far
    last_read = {}           # scratchN -> index of the last step that read it (scratch buffers are reused)
    outs = {}
    for step_idx, st in enumerate(parse_plan(plan)):
        if st[0] == "fir":
            _, nid, term = st
            assert f"firU#{nid}" in written, f"fir#{nid} runs before its input was stored"
            written.add(f"firY#{nid}")
            if term is not None:
                outs[term] = outs.get(term, 0) + 1
            continue
        _, ops, binds, hdr = st
        m = re.match(r"G=(\d+) channels x S=(\d+) samples per CTA, (\d+) ops, (\d+) smem vregs", hdr)
        G, S_, n_ops, n_vregs = map(int, m.groups())
        assert G * S_ == 4096 and n_ops == len(ops) <= 47 and len(binds) <= 12, hdr
        assert n_vregs * 16384 < 200 * 1024, hdr
        saved, stored_here = set(), set()
        for i, op in enumerate(ops):
            c = op["code"]
            if c in READS_V and op["v"] != -1:
                assert op["v"] in saved, f"op {i} reads v{op['v']} before it is saved: {hdr}"
            for v in op["pv"]:
                assert v in saved, f"op {i} reads parameter tile v{v} before it is saved"
            if c == OP_SAVEV:
                assert 0 <= op["v"] < n_vregs
                saved.add(op["v"])
            if c in READS_G:
                what = binds[op["g"]]
                assert not what.startswith("out") and not what.startswith("firU"), what
                if not what.startswith("in"):
                    assert what in written, f"op {i} reads {what} before any step wrote it"
                    last_read[what] = step_idx
                    if what in stored_here:
                        assert not op["pf"], f"op {i} prefetches {what}, which this very segment writes"
            if c == OP_STOREG:
                what = binds[op["g"]]
                assert not what.startswith("in") and not what.startswith("firY"), what
                if what.startswith("out"):
                    t = int(what[3:])
                    outs[t] = outs.get(t, 0) + 1
                else:
                    # a scratch buffer may be given a new value only by a LATER step than every reader of the old one
                    assert what not in stored_here, f"{what} written twice in one segment"
                    assert last_read.get(what, -1) < step_idx, f"{what} overwritten in the step that still reads its old value"
                    written.add(what)
                    stored_here.add(what)
    assert outs == {t: 1 for t in range(n_out_terms)}, outs


def plan_of(graph, channels, iir_mode=0, max_samples=128 * 9):
    from dsp_stuff_b200.engine import Engine

    e = Engine(channels, block=128, max_samples=max_samples, device=-1, iir_mode=iir_mode)
    graph.apply(e)
    return e.describe_plan()


@pytest.mark.parametrize("seed", range(40))
def test_small_random_graphs_lower_with_sound_dataflow(seed):
    check_dataflow(plan_of(S.random_graph(seed), 64 if seed % 2 else 4096), 2)


@pytest.mark.parametrize("seed", range(60))
def test_large_random_graphs_are_cut_into_segments(seed):
    """10 - 69 nodes with up to three links per port: far more than one Program holds.  Round 2 and before this failed with
    "segment has too many ops"; now the segment is cut and the cut values go through scratch."""
    n_nodes = 10 + seed
    plan = plan_of(S.random_graph(1000 + seed, n_nodes), 64 if seed % 2 else 4096, iir_mode=seed % 3 == 0)
    check_dataflow(plan, 2)
    if n_nodes >= 30:
        assert plan.count("fused segment:") > plan.count("fir step:") + 1   # at least one cut that no FIR node forced


def test_long_chain_of_stateful_nodes_is_cut_at_the_state_limit():
    """40 biquads in a row: 12 state slots per Program -> at least four segments, each handing one value to the next."""
    from dsp_stuff_b200 import GraphSpec

    g = GraphSpec().node(100, "input").node(101, "output")
    prev = (100, "out")
    for i in range(40):
        g.node(i, "biquad", **S.rbj_biquad("lp", 500.0 + 100 * i))
        g.link(prev[0], prev[1], i, "in")
        prev = (i, "out")
    g.link(prev[0], prev[1], 101, "in")
    plan = plan_of(g, 256)
    check_dataflow(plan, 1)
    assert plan.count("fused segment:") >= 4
    assert plan.count("DF1(") == 40
    # one value crosses each cut; scratch buffers are reused once their reader's step is over: two buffers, ping-pong
    assert len(set(re.findall(r"=scratch(\d+)", plan))) == 2


def test_many_reverbs_are_cut_at_the_ring_limit():
    from dsp_stuff_b200 import GraphSpec

    g = GraphSpec().node(100, "input").node(101, "output")
    prev = (100, "out")
    for i in range(15):
        g.node(i, "reverb", seconds=0.01 + 0.002 * i, decay=0.5)
        g.link(prev[0], prev[1], i, "in")
        prev = (i, "out")
    g.link(prev[0], prev[1], 101, "in")
    plan = plan_of(g, 64)
    check_dataflow(plan, 1)
    assert plan.count("fused segment:") == 3        # 6 rings per Program


def test_a_single_node_that_cannot_fit_still_fails_loudly():
    """One node with 200 links into one port cannot be cut (a cut goes in front of a node): error code, not a wrong plan."""
    from dsp_stuff_b200 import GraphSpec
    from dsp_stuff_b200.engine import EngineError

    g = GraphSpec().node(100, "input").node(101, "output").node(0, "gain", level=1.0)
    for _ in range(200):
        g.link(100, "out", 0, "in")
    g.link(0, "out", 101, "in")
    with pytest.raises(EngineError) as ei:
        plan_of(g, 4)
    assert "too many ops" in str(ei.value)


@pytest.mark.parametrize("seed", range(30))
def test_saved_graph_round_trip_lowers_to_the_same_plan(seed):
    """GraphSpec -> saved-graph JSON -> the C++ loader (dspb_load_graph_json) and -> the Python parser -> builder calls: all
    three routes must produce the same schedule, parameter for parameter (the listing prints every coefficient)."""
    from dsp_stuff_b200 import GraphSpec
    from dsp_stuff_b200.engine import Engine

    g = S.random_graph(seed, None if seed < 15 else 10 + seed)
    text = g.to_json()
    plans = [plan_of(g, 64), plan_of(GraphSpec.from_json(text), 64)]
    e = Engine(64, block=128, max_samples=128 * 9, device=-1)
    e.load_graph_json(text)
    plans.append(e.describe_plan())
    assert plans[0] == plans[1] == plans[2]


def test_scan_mode_cuts_at_the_scan_table_limit():
    """iir_mode = 1: a Program holds four scan tables.  A fifth qualifying filter used to stay sequential and, all or nothing,
    take the other four back to exact with it; now the segment is cut so that every filter stays time-parallel."""
    from dsp_stuff_b200 import GraphSpec

    g = GraphSpec().node(100, "input").node(101, "output")
    prev = 100
    for i in range(10):
        g.node(i, "low_pass", ratio=0.5 + 0.04 * i)
        g.link(prev, "out", i, "in")
        prev = i
    g.link(prev, "out", 101, "in")
    plan = plan_of(g, 256, iir_mode=1)
    check_dataflow(plan, 1)
    assert plan.count("fused segment:") == 3
    assert plan.count("time-parallel scan") == 10 and "exact" not in plan
    assert plan_of(g, 256, iir_mode=0).count("fused segment:") == 1     # the exact chain still fits one Program
