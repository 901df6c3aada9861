"""Graph description: the host-side mirror of the reference's DSPConfig.

Reference: `DSPConfig{nodes: Vec<NodeConfig>, links: Vec<LinkConfig>}` (dsp-stuff/src/runtime.rs:44-48),
`NodeConfig{id, typename, position, cfg}` (runtime.rs:606-612), `LinkConfig{lhs, rhs}`
(runtime.rs:560-564); per-node `cfg` is the derive-generated `<Name>Config`
(dsp-stuff-derive/src/lib.rs:266-293): id, inputs/outputs name->PortId maps and every
`#[dsp(save)]` field.

A GraphSpec is pure data.  `apply(builder)` replays it onto anything exposing
add_node / set_f32 / set_enum / set_taps / link / compile — the CUDA Engine in production.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

# Port order per typename = declared input=/output= attrs, then slider(as_input) fields in struct
# order (dsp-stuff-derive/src/lib.rs:191-219).  Mirrors SURVEY.md Appendix A.
NODE_PORTS: Dict[str, Tuple[Tuple[str, ...], Tuple[str, ...]]] = {
    "gain": (("in", "level"), ("out",)),
    "distort": (("in", "level"), ("out",)),
    "overdrive": (("in", "boost", "drive", "level"), ("out",)),
    "chebyshev": (("in",), ("out",)),
    "biquad": (("in",), ("out",)),
    "low_pass": (("in",), ("out",)),
    "high_pass": (("in",), ("out",)),
    "reverb": (("in",), ("out",)),
    "fir": (("in",), ("out",)),
    "add": (("a", "b"), ("out",)),
    "mix": (("a", "b", "ratio"), ("out",)),
    "mux": (("a", "b"), ("out",)),
    "demux": (("in",), ("a", "b")),
    "envelope": (("in",), ("out",)),
    "signal_gen": (("amplitude", "frequency"), ("out",)),
    "gate": (("in",), ("out",)),  # extension: not a reference node (see csrc/engine.cpp kNodeTypes)
    "input": ((), ("out",)),
    "output": (("in",), ()),
}

# GUI-only sink typenames (nodes/mod.rs:111-122): restored by the reference, irrelevant to any audio value.
GUI_SINKS = ("wave_view", "spectrogram", "pitch")

# Saved enum fields per typename (serialised as the variant name string).
ENUM_FIELDS = {
    "distort": ("mode",),
    "fir": ("mode",),
    "mux": ("in_port",),
    "demux": ("out_port",),
    "signal_gen": ("mode",),
}


@dataclass
class NodeSpec:
    id: int
    typename: str
    f32: Dict[str, float] = field(default_factory=dict)
    enums: Dict[str, str] = field(default_factory=dict)
    taps: Optional[Sequence[float]] = None  # fir only, stored reversed like the reference
    restored: bool = False  # came from saved JSON: after_settings_change runs even with no field change


@dataclass
class GraphSpec:
    nodes: List[NodeSpec] = field(default_factory=list)
    links: List[Tuple[int, str, int, str]] = field(default_factory=list)

    # ---- building -------------------------------------------------------------------------------
    def node(self, id: int, typename: str, taps=None, **params) -> "GraphSpec":
        if typename not in NODE_PORTS:
            raise KeyError(f"unknown typename {typename!r}")
        n = NodeSpec(id, typename, taps=taps)
        for k, v in params.items():
            if isinstance(v, str):
                n.enums[k] = v
            else:
                n.f32[k] = float(v)
        self.nodes.append(n)
        return self

    def link(self, src: int, out_port: str, dst: int, in_port: str) -> "GraphSpec":
        self.links.append((src, out_port, dst, in_port))
        return self

    def chain(self, ids: Sequence[int]) -> "GraphSpec":
        """Link consecutive single-in/single-out nodes through their first ports."""
        by_id = {n.id: n for n in self.nodes}
        for a, b in zip(ids[:-1], ids[1:]):
            self.link(a, NODE_PORTS[by_id[a].typename][1][0], b, NODE_PORTS[by_id[b].typename][0][0])
        return self

    def apply(self, builder) -> None:
        for n in self.nodes:
            builder.add_node(n.typename, n.id)
            if n.taps is not None:
                builder.set_taps(n.id, n.taps)
            for k, v in n.enums.items():
                builder.set_enum(n.id, k, v)
            for k, v in n.f32.items():
                builder.set_f32(n.id, k, v)
        for l in self.links:
            builder.link(*l)
        builder.compile()

    # ---- the reference's saved-graph JSON ----------------------------------
    def to_json(self) -> str:
        port_ids: Dict[Tuple[int, str, bool], int] = {}
        next_pid = 0
        nodes = []
        for i, n in enumerate(self.nodes):
            ins, outs = NODE_PORTS[n.typename]
            cfg: Dict[str, object] = {"id": n.id, "inputs": {}, "outputs": {}}
            for p in ins:
                port_ids[(n.id, p, False)] = next_pid
                cfg["inputs"][p] = next_pid
                next_pid += 1
            for p in outs:
                port_ids[(n.id, p, True)] = next_pid
                cfg["outputs"][p] = next_pid
                next_pid += 1
            cfg.update(n.f32)
            cfg.update(n.enums)
            if n.taps is not None:
                cfg["taps"] = [float(t) for t in n.taps]
                cfg.setdefault("file_name", None)
            nodes.append({"id": n.id, "typename": n.typename, "position": [100.0 + 200.0 * i, 100.0], "cfg": cfg})
        links = [{"lhs": [s, port_ids[(s, op, True)]], "rhs": [d, port_ids[(d, ip, False)]]}
                 for (s, op, d, ip) in self.links]
        return json.dumps({"nodes": nodes, "links": links})

    @staticmethod
    def from_json(text: str) -> "GraphSpec":
        doc = json.loads(text)
        g = GraphSpec()
        pid_name: Dict[Tuple[int, int], Tuple[str, bool]] = {}
        dropped = set()
        for nd in doc["nodes"]:
            typename = nd["typename"]
            if typename in GUI_SINKS:  # oscilloscope / spectrogram / pitch read-out: no output port, dropped with their links
                dropped.add(int(nd["id"]))
                continue
            if typename not in NODE_PORTS:
                raise KeyError(f"unknown typename {typename!r}")  # reference panics: runtime.rs:634-637
            cfg = nd["cfg"]
            spec = NodeSpec(int(nd["id"]), typename, restored=True)
            for name, pid in cfg.get("inputs", {}).items():
                pid_name[(spec.id, int(pid))] = (name, False)
            for name, pid in cfg.get("outputs", {}).items():
                pid_name[(spec.id, int(pid))] = (name, True)
            for k, v in cfg.items():
                if k in ("id", "inputs", "outputs", "file_name"):
                    continue
                if typename in ("input", "output") and k in ("selected_host", "selected_device"):
                    continue  # cpal host / device names (nodes/input.rs:33-38, nodes/output.rs:33-38)
                if k == "taps":
                    spec.taps = [float(t) for t in v]
                elif isinstance(v, str):
                    spec.enums[k] = v
                elif isinstance(v, (int, float)):
                    spec.f32[k] = float(v)
            g.nodes.append(spec)
        for l in doc["links"]:
            (sn, sp), (dn, dp) = l["lhs"], l["rhs"]
            if int(dn) in dropped:
                continue
            g.links.append((int(sn), pid_name[(int(sn), int(sp))][0], int(dn), pid_name[(int(dn), int(dp))][0]))
        return g
