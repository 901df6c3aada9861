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
