"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header declares,
and refuses to run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from dsp_stuff_b200 import build, engine

    build.build()
    return engine.load_library()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "dspb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dspb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from dsp_stuff_b200 import engine

    declared = header_symbols()
    assert declared, "no declarations found in include/dspb200.h"
    for name in declared:
        assert hasattr(lib, name), f"libdspb200.so does not export {name}"
    assert sorted(engine.ABI_SYMBOLS) == declared


def test_abi_version(lib):
    assert lib.dspb_abi_version() == 2


def test_config_struct_matches_header():
    from dsp_stuff_b200 import engine

    assert ctypes.sizeof(engine.Config) == 6 * 4 + 8 + 4 * 4   # 3 x int32 + tail padding to the int64 alignment
    assert [f[0] for f in engine.Config._fields_] == ["channels", "block", "sample_rate", "ref_block", "ring_granule",
                                                     "device", "max_samples", "fir_fft_log2", "fir_mode", "iir_mode"]


def test_engine_fails_loudly_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dsp_stuff_b200.engine import Engine, EngineError

    with pytest.raises(EngineError) as ei:
        Engine(4)
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "dsp_stuff_b200")
    for dp, _, files in os.walk(pkg):
        if "build" in dp.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "liboracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_planning_mode_describes_every_fir_kernel_and_extension_node(lib):
    """device = -1 builds schedules without a GPU (and refuses to process): the host scheduler is CPU-testable."""
    from dsp_stuff_b200 import GraphSpec
    from dsp_stuff_b200 import signals as S
    from dsp_stuff_b200.engine import Engine, EngineError

    want = {0: "overlap-save FFT 2^13 (+ 2^14 double segments", 1: "direct f64 sum", 2: "Toeplitz-tiled tcgen05 GEMM", 3: "packed f32x2 variant"}
    for mode, text in want.items():
        e = Engine(4096, block=1024, max_samples=16384, device=-1, fir_mode=mode)
        S.target_chain(4096).apply(e)
        plan = e.describe_plan()
        assert text in plan, (mode, plan[-300:])
        assert "fused segment: G=16" in plan
    with pytest.raises(EngineError):
        Engine(4, device=-1, fir_mode=7)
    # the gate extension lowers to save -> envelope -> select
    g = (GraphSpec().node(10, "input").node(11, "output").node(0, "gate", threshold=0.25, release=50.0)
         .link(10, "out", 0, "in").link(0, "out", 11, "in"))
    e = Engine(64, device=-1)
    g.apply(e)
    plan = e.describe_plan()
    assert "envelope(" in plan and "(extension)" in plan


@pytest.mark.parametrize("workload,channels,expect", [
    ("target", 4096, "G=16"),    # many channels per SM: >= 256 CTAs, the launcher then trims to 8-channel CTAs
    ("config3", 1024, "G=8"),    # light chain, <= 2048 channels: one wave of <= 128 CTAs (exclusive-R layout)
    ("config2", 256, "G=2"),
    ("config5", 1024, "G=4"),    # heavy graph segment (17 ops, shared-memory vregs): 256 CTAs at two per SM
])
def test_tile_geometry_heuristic(lib, workload, channels, expect):
    """The scheduler's CTA geometry (csrc/engine.cpp close_fused) for the BASELINE configs, without a GPU."""
    from dsp_stuff_b200 import signals as S
    from dsp_stuff_b200.engine import Engine

    e = Engine(channels, block=1024, max_samples=16384, device=-1)
    S.WORKLOADS[workload][0]().apply(e)
    plan = e.describe_plan()
    segs = [l for l in plan.splitlines() if "fused segment" in l]
    assert any(expect + " " in l for l in segs), segs


def test_engine_registry_mirrors_the_dsp_attributes(lib):
    """Product-side node registry (csrc/engine.cpp kNodeTypes) against SURVEY.md Appendix A: port names in index order
    (declared input= attrs, then slider(as_input) fields in struct order, lib.rs:191-219), through the C ABI in
    planning mode (no GPU)."""
    from dsp_stuff_b200 import NODE_PORTS
    from dsp_stuff_b200.engine import Engine, EngineError

    appendix_a = {
        "gain": (("in", "level"), ("out",)), "distort": (("in", "level"), ("out",)),
        "overdrive": (("in", "boost", "drive", "level"), ("out",)), "chebyshev": (("in",), ("out",)),
        "biquad": (("in",), ("out",)), "low_pass": (("in",), ("out",)), "high_pass": (("in",), ("out",)),
        "reverb": (("in",), ("out",)), "fir": (("in",), ("out",)), "add": (("a", "b"), ("out",)),
        "mix": (("a", "b", "ratio"), ("out",)), "mux": (("a", "b"), ("out",)), "demux": (("in",), ("a", "b")),
        "envelope": (("in",), ("out",)), "signal_gen": (("amplitude", "frequency"), ("out",)),
        "input": ((), ("out",)), "output": (("in",), ()),
    }
    e = Engine(4, device=-1)
    for i, (typename, (ins, outs)) in enumerate(appendix_a.items()):
        e.add_node(typename, 100 + i)
        assert NODE_PORTS[typename] == (ins, outs), typename          # the Python mirror
        assert e.get_i64(100 + i, "n_inputs") == len(ins) and e.get_i64(100 + i, "n_outputs") == len(outs), typename
        assert [e.port_index(100 + i, p) for p in ins] == list(range(len(ins))), typename
        assert [e.port_index(100 + i, p, True) for p in outs] == list(range(len(outs))), typename
        with pytest.raises(EngineError):
            e.port_index(100 + i, "no_such_port")
    with pytest.raises(EngineError):
        e.add_node("muff", 999)        # GPL dependency, out of scope: unknown typename is an error code, not a panic
    with pytest.raises(EngineError):
        e.add_node("gain", 100)        # duplicate id


def test_engine_rejects_cycles_and_unknown_ports_with_status_codes(lib):
    from dsp_stuff_b200.engine import Engine, EngineError

    e = Engine(2, device=-1)
    for i, t in enumerate(["input", "gain", "gain", "output"]):
        e.add_node(t, i)
    e.link(0, "out", 1, "in")
    e.link(1, "out", 2, "in")
    e.link(2, "out", 1, "in")          # cycle: a deadlock in the reference (runtime.rs:568), DSPB_ERR_GRAPH here
    e.link(2, "out", 3, "in")
    with pytest.raises(EngineError) as ei:
        e.compile()
    assert "cycl" in str(ei.value).lower()
    with pytest.raises(EngineError):
        e.link(0, "nope", 1, "in")
    with pytest.raises(EngineError):
        e.set_f32(1, "no_such_field", 1.0)


@pytest.mark.parametrize("name,expect", [
    ("config1", ["gain#", "distort[SoftClip]", "ring*0.5"]),
    ("config2", ["DF1(", "biquad#"]),
    ("config3", ["gain#", "distort[SoftClip]", "DF1(", "D=12288"]),
    ("config5_64taps", ["fir step", "mix#", "add#", "D=6144", "acc /= 2.00010014"]),
    ("target_256taps", ["fir step", "256 taps", "D=12288"]),
])
def test_saved_graph_json_lowers_in_planning_mode(lib, name, expect):
    """The C++ saved-graph loader (dspb_load_graph_json, reference DSPConfig format: runtime.rs:44-48, 606-612) and
    the scheduler on the committed fixture graphs, without a GPU."""
    from dsp_stuff_b200.engine import Engine, EngineError

    text = open(os.path.join(ROOT, "tests", "golden", f"{name}_graph.json")).read()
    e = Engine(64, block=128, max_samples=128 * 8, device=-1)
    e.load_graph_json(text)
    plan = e.describe_plan()
    for s in expect:
        assert s in plan, (s, plan)
    with pytest.raises(EngineError):          # planning mode never processes: no CPU fallback
        import numpy as np
        e.process(np.zeros((64, 128), np.float32))
    with pytest.raises(EngineError):
        Engine(4, device=-1).load_graph_json('{"nodes": [{"id": 0, "typename": "no_such_node", "position": [0, 0], "cfg": {}}], "links": []}')
    with pytest.raises(EngineError):
        Engine(4, device=-1).load_graph_json('{"nodes": [')


@pytest.mark.parametrize("name,expect,absent", [
    ("ref_shape_pedalboard", ["gain#3", "distort[SoftClip]", "DF1(", "high_pass(acc; ratio=0.75)", "D=12288", "fir step: fir#13, 16 taps"],
     ["wave_view", "spectrogram", "pitch"]),
    ("ref_shape_fir_default", ["fir step: fir#5, 1 taps", "signal_gen[Triangle](amp=0.25, freq=440)", "mix#7"], []),
])
def test_reference_shaped_saved_graph_loads_in_planning_mode(lib, name, expect, absent):
    """dspb_load_graph_json on graphs written the way the reference's serde structs write them (Input / Output carry
    selected_host / selected_device, GUI sinks are attached): tests/golden/ref_shape_*.json are hand-written."""
    from dsp_stuff_b200.engine import Engine

    text = open(os.path.join(ROOT, "tests", "golden", f"{name}.json")).read()
    e = Engine(64, block=128, max_samples=128 * 8, device=-1)
    e.load_graph_json(text)
    plan = e.describe_plan()
    for s in expect:
        assert s in plan, (s, plan)
    for s in absent:
        assert s not in plan, (s, plan)


def test_saved_graph_with_muff_is_an_unknown_node_error(lib):
    import json

    from dsp_stuff_b200.engine import Engine, EngineError

    doc = json.loads(open(os.path.join(ROOT, "tests", "golden", "ref_shape_fir_default.json")).read())
    doc["nodes"].append({"id": 40, "typename": "muff", "position": [0.0, 0.0],
                         "cfg": {"id": 40, "inputs": {"in": 41}, "outputs": {"out": 42}, "toan": 0.5, "level": 0.5, "sustain": 0.5}})
    with pytest.raises(EngineError) as ei:
        Engine(4, device=-1).load_graph_json(json.dumps(doc))
    assert ei.value.code == -2 and "muff" in str(ei.value)


def _lowered(plan):
    """'lowered:' lines of describe_plan(): per fused step the ops as the kernel sees them, [(code, attrs...)]"""
    out = []
    for line in plan.splitlines():
        if line.strip().startswith("lowered:"):
            out.append([tok.split(":") for tok in line.split("lowered:")[1].split()])
    return out


OP_LOADG, OP_SAVEV, OP_STOREG, OP_GATE = "3", "10", "11", "25"   # csrc/plan.h OpCode


@pytest.mark.parametrize("front", ["gain", "biquad", "distort"])
def test_gate_operand_is_remapped_to_an_allocated_vreg(lib, front):
    """the gate's saved input must be renamed to its physical shared-memory slot like every other vreg
    operand (it used to keep its virtual id, i.e. read a slot that was never allocated once a node sat in front)."""
    import re

    from dsp_stuff_b200 import GraphSpec
    from dsp_stuff_b200.engine import Engine

    g = (GraphSpec().node(10, "input").node(11, "output").node(0, front).node(1, "gate", threshold=0.3, release=40.0)
         .link(10, "out", 0, "in").link(0, "out", 1, "in").link(1, "out", 11, "in"))
    e = Engine(64, device=-1)
    g.apply(e)
    plan = e.describe_plan()
    n_vregs = int(re.search(r"(\d+) smem vregs", plan).group(1))
    ops = _lowered(plan)[0]
    gates = [o for o in ops if o[0] == OP_GATE]
    saves = {o[1] for o in ops if o[0] == OP_SAVEV}
    assert len(gates) == 1 and n_vregs >= 1
    assert int(gates[0][1][1:]) < n_vregs and gates[0][1] in saves, (plan, ops)


def test_value_spilled_and_reused_in_one_step_is_not_prefetched(lib):
    """signal_gen -> {fir, gain}: the generator's value is stored to scratch for the FIR step and re-read by
    the gain in the same kernel; a one-tile-ahead register prefetch of that buffer would read the previous call's data."""
    from dsp_stuff_b200 import GraphSpec
    from dsp_stuff_b200.engine import Engine

    g = (GraphSpec().node(11, "output").node(12, "output").node(0, "signal_gen", frequency=440.0)
         .node(1, "fir", taps=[0.5, 0.25, 0.125]).node(2, "gain", level=2.0)
         .link(0, "out", 1, "in").link(0, "out", 2, "in").link(1, "out", 11, "in").link(2, "out", 12, "in"))
    e = Engine(64, device=-1)
    g.apply(e)
    steps = _lowered(e.describe_plan())
    for ops in steps:
        stored = set()
        for o in ops:
            if o[0] == OP_STOREG:
                stored.add(o[1].rstrip("*"))
            if o[0] == OP_LOADG and o[1].rstrip("*") in stored:
                assert not o[1].endswith("*"), ("prefetched although stored earlier in the same program", ops)


def test_iir_mode_lowers_to_scan_segments_in_planning_mode(lib):
    """dspb_config::iir_mode = 1 (opt-in time-parallel recurrences): in planning mode every filter is assumed to pass the
    probe, so the schedule shows the scan verdict, no transposition tile, and as many CTAs as fill the machine."""
    import re

    from dsp_stuff_b200 import signals as S
    from dsp_stuff_b200.engine import Engine, EngineError

    for C, G in [(256, 1), (512, 2), (4096, 16)]:
        e = Engine(C, block=1024, max_samples=24576, device=-1, iir_mode=1)
        S.config3().apply(e)
        plan = e.describe_plan()
        assert "time-parallel scan (probe error" in plan and "exact, lane=channel" not in plan, plan
        assert re.search(rf"fused segment: G={G} channels", plan), plan
    e = Engine(256, block=1024, max_samples=24576, device=-1)      # default: exact
    S.config3().apply(e)
    assert "time-parallel" not in e.describe_plan()
    e = Engine(64, device=-1, iir_mode=1)                           # an envelope stays sequential -> the biquad next to it too
    g = S.config3()
    g.nodes.insert(3, type(g.nodes[0])(7, "envelope"))
    g.links = [l for l in g.links if not (l[0] == 2 and l[2] == 3)] + [(2, "out", 7, "in"), (7, "out", 3, "in")]
    g.apply(e)
    assert "exact after all" in e.describe_plan()
    with pytest.raises(EngineError):
        Engine(4, device=-1, iir_mode=5)
