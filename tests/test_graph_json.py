"""Saved-graph JSON (reference DSPConfig, runtime.rs:44-48): host-side round trip, CPU only."""
import json

from dsp_stuff_b200 import GraphSpec
from dsp_stuff_b200 import signals as S
from tests.util import assert_bit_exact, make_oracle


def test_round_trip_preserves_graph(oracle_mod):
    g = S.config5(n_taps=64)
    text = g.to_json()
    doc = json.loads(text)
    assert set(doc) == {"nodes", "links"}
    assert all(set(n) == {"id", "typename", "position", "cfg"} for n in doc["nodes"])
    g2 = GraphSpec.from_json(text)
    assert g2.links == g.links
    x = S.noise(2, 512)
    a = make_oracle(oracle_mod, g, 2).process(x)[0]
    b = make_oracle(oracle_mod, g2, 2).process(x)[0]
    assert_bit_exact(a, b)


def test_appendix_c_example_parses():
    text = """{ "nodes": [
        { "id": 0, "typename": "gain", "position": [100.0, 100.0],
          "cfg": { "id": 0, "inputs": {"in": 0, "level": 1}, "outputs": {"out": 2}, "level": 2.0 } },
        { "id": 1, "typename": "distort", "position": [300.0, 100.0],
          "cfg": { "id": 1, "inputs": {"in": 3, "level": 4}, "outputs": {"out": 5}, "level": 4.0, "mode": "SoftClip" } },
        { "id": 2, "typename": "reverb", "position": [500.0, 100.0],
          "cfg": { "id": 2, "inputs": {"in": 6}, "outputs": {"out": 7}, "seconds": 0.25, "decay": 0.5 } } ],
      "links": [ { "lhs": [0, 2], "rhs": [1, 3] }, { "lhs": [1, 5], "rhs": [2, 6] } ] }"""
    g = GraphSpec.from_json(text)
    assert [n.typename for n in g.nodes] == ["gain", "distort", "reverb"]
    assert g.links == [(0, "out", 1, "in"), (1, "out", 2, "in")]
    assert g.nodes[1].enums == {"mode": "SoftClip"} and g.nodes[2].f32 == {"seconds": 0.25, "decay": 0.5}


# ---- graphs in the exact shape the reference GUI saves (hand-written, NOT produced by GraphSpec.to_json) -----------
# tests/golden/ref_shape_*.json follow the serde output of the reference field by field:
#   NodeConfig{id, typename, position:(f32, f32), cfg}                       runtime.rs:606-612
#   InputConfig{id, selected_host, selected_device: Option<String>, outputs} nodes/input.rs:33-38
#   OutputConfig{id, selected_host, selected_device, inputs}                 nodes/output.rs:33-38
#   derive-generated <Name>Config: id, inputs, outputs, every #[dsp(save)] field (lib.rs:233-293); serde_json's
#   default map is a BTreeMap, so keys come out sorted; Option::None is null (Fir.file_name, selected_device)
#   WaveViewConfig{id, inputs}, SpectrogramConfig{id, inputs, buffer_size, fft_size, upper_bound, lower_bound}
#   PortIds come from one global counter (ids.rs), so they are unique across the graph and have gaps.
# A saved "Low Pass" node carries typename "high_pass" (its cfg_name() is wrong: nodes/low_pass.rs:9) and restores as a
# HighPass in the reference too -- node 9 of the pedalboard graph is such an entry.
import os

import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_SHAPED = ["ref_shape_pedalboard", "ref_shape_fir_default"]


def test_reference_shaped_graph_parses_and_drops_gui_sinks():
    g = GraphSpec.from_json(open(os.path.join(GOLDEN, "ref_shape_pedalboard.json")).read())
    assert [n.typename for n in g.nodes] == ["input", "gain", "distort", "biquad", "high_pass", "reverb", "fir", "output"]
    assert (3, "out", 7, "in") in g.links and (13, "out", 16, "in") in g.links
    assert all(d not in (4, 14, 15) for (_, _, d, _) in g.links)       # wave_view / spectrogram / pitch links are gone
    assert g.nodes[0].enums == {} and g.nodes[-1].enums == {}           # selected_host is not an enum field
    fir = [n for n in g.nodes if n.typename == "fir"][0]
    assert len(fir.taps) == 16 and fir.enums == {"mode": "Average"}


@pytest.mark.parametrize("name", REF_SHAPED)
def test_reference_shaped_graph_runs_on_the_oracle(oracle_mod, name):
    g = GraphSpec.from_json(open(os.path.join(GOLDEN, f"{name}.json")).read())
    y = make_oracle(oracle_mod, g, 2).process(S.noise(2, 1024))[0]
    assert y.shape == (2, 1024) and float(abs(y).max()) > 0


def test_muff_and_unknown_typenames_are_errors():
    doc = json.loads(open(os.path.join(GOLDEN, "ref_shape_fir_default.json")).read())
    doc["nodes"].append({"id": 40, "typename": "muff", "position": [0.0, 0.0],
                         "cfg": {"id": 40, "inputs": {"in": 41}, "outputs": {"out": 42}, "toan": 0.5, "level": 0.5, "sustain": 0.5}})
    with pytest.raises(KeyError):
        GraphSpec.from_json(json.dumps(doc))
