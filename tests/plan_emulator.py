"""Executes a LOWERED plan (the listing `dspb_describe_plan` prints) on the CPU.  TEST INFRASTRUCTURE ONLY.

What it is for: the host scheduler (csrc/engine.cpp Lowerer) decides the order of the ops, folds the fan-in averages into
their consumers, allocates shared-memory slots by liveness, cuts segments and routes values through scratch.  None of that
needs a GPU to be wrong.  This module walks the lowered op list exactly as the kernels do -- one accumulator, physical
shared-memory slots, global buffers by their launch-time binding -- and takes the ARITHMETIC of every op from the numpy
restatement of the reference nodes (oracle/np_oracle.py `_Node`).  If the scheduler is right, the result equals the
oracle's run of the original graph bit for bit (tests/test_plan_semantics.py); a slot reused while still live, a fan-in
average applied to the wrong value, a cut value read from the wrong scratch buffer all show up as different samples.

It does not model the kernels (tiling, prefetch, pipelining): those are the GPU parity tests' job.
"""
import re

import numpy as np

from oracle.np_oracle import BUF_SIZE, NpOracle, F

OP_NOP, OP_ZERO, OP_LOADG, OP_ADDG, OP_LOADV, OP_ADDV, OP_COPYV, OP_COPYG = 1, 2, 3, 4, 5, 6, 7, 8
OP_SAVEV, OP_STOREG, OP_MODMAP = 10, 11, 12
OP_GAIN, OP_DISTORT, OP_OVERDRIVE, OP_CHEBY, OP_ADD, OP_MIX, OP_COMB = 13, 14, 15, 16, 17, 18, 19
OP_BIQUAD, OP_LP1, OP_HP1, OP_ENVELOPE, OP_SIGGEN, OP_GATE = 20, 21, 22, 23, 24, 26
NODE_OPS = {OP_GAIN, OP_DISTORT, OP_OVERDRIVE, OP_CHEBY, OP_ADD, OP_MIX, OP_COMB, OP_BIQUAD, OP_LP1, OP_HP1, OP_ENVELOPE, OP_SIGGEN}
# which parameter P0 / P1 / P2 of an op is, per node type (csrc/engine.cpp ctl_param calls)
PARAM_ORDER = {"gain": ["level"], "distort": ["level"], "overdrive": ["boost", "drive", "level"], "mix": ["ratio"],
               "signal_gen": ["amplitude", "frequency"]}


def parse(plan):
    """-> steps: dict(kind="fused", ops=[...], text=[...], binds={slot: name}) | dict(kind="fir", node=id, term=t|None)"""
    steps = []
    for line in plan.splitlines():
        m = re.match(r"\[\d+\] fused segment:", line)
        if m:
            steps.append(dict(kind="fused", ops=[], text=[], binds={}))
            continue
        m = re.match(r"\[\d+\] fir step: fir#(\d+),(.*)", line)
        if m:
            t = re.search(r"epilogue \(0\.0 \+ y\)/(\S+) -> output terminal (\d+)", m.group(2))
            steps.append(dict(kind="fir", node=int(m.group(1)), term=int(t.group(2)) if t else None, nf=F(t.group(1)) if t else None))
            continue
        if not steps or steps[-1]["kind"] != "fused" or not line.startswith("    "):
            continue
        st = steps[-1]
        if line.startswith("    lowered:"):
            for tok in line.split()[1:]:
                parts = tok.split(":")
                op = dict(code=int(parts[0]), v=None, g=None, pk={})
                for q in parts[1:]:
                    if q.startswith("v"):
                        op["v"] = int(q[1:])
                    elif q.startswith("g"):
                        op["g"] = int(q[1:].rstrip("*"))
                    elif q.startswith("p"):
                        k, v = q[1:].split("=v")
                        op["pk"][int(k)] = int(v)
                st["ops"].append(op)
        elif line.startswith("    buffers:"):
            st["binds"] = {int(k): v for k, v in re.findall(r"g(\d+)=(\S+)", line)}
        else:
            st["text"].append(line[4:])
    for st in steps:
        if st["kind"] == "fused":
            assert len(st["ops"]) == len(st["text"]), (len(st["ops"]), len(st["text"]))
    return steps


class PlanEmulator:
    def __init__(self, plan, spec, channels, ring_granule=1024):
        self.steps = parse(plan)
        self.C = channels
        self.np = NpOracle(channels, ring_granule=ring_granule)   # only its nodes (parameters + state) are used
        spec.apply(self.np)
        self.nodes = self.np.nodes

    def _node_of(self, text):
        m = re.findall(r"; (\w+)#(\d+)", text)
        assert m, text
        return int(m[-1][1])

    def _blocks(self, n):
        return [slice(b * BUF_SIZE, (b + 1) * BUF_SIZE) for b in range(n // BUF_SIZE)]

    def _run_node(self, nid, ins, params, n):
        """node.process block by block; `params` = {field: [C x n] tile} for connected control ports (already range-mapped)."""
        nd = self.nodes[nid]
        out = np.empty((self.C, n), F)
        keep = {k: nd.p[k] for k in params}
        try:
            for sl in self._blocks(n):
                for k, tile in params.items():
                    nd.p[k] = tile[:, sl]          # _Node.param broadcasts the scalar; a tile goes through unchanged
                with np.errstate(all="ignore"):
                    out[:, sl] = nd.process({k: v[:, sl] for k, v in ins.items()}, {})["out"]
        finally:
            nd.p.update(keep)
        return out

    def process(self, inputs, n=None):
        xs = [np.ascontiguousarray(x, dtype=F) for x in ([inputs] if isinstance(inputs, np.ndarray) else inputs)]
        n = xs[0].shape[1] if xs else int(n)
        bufs = {f"in{i}": x for i, x in enumerate(xs)}
        zero = F(0.0)
        for st in self.steps:
            if st["kind"] == "fir":
                nd = self.nodes[st["node"]]
                y = np.empty((self.C, n), F)
                u = bufs[f"firU#{st['node']}"]
                for sl in self._blocks(n):
                    y[:, sl] = nd.process({"in": u[:, sl]}, {})["out"]
                if st["term"] is not None:
                    bufs[f"out{st['term']}"] = ((zero + y) / st["nf"]).astype(F)
                else:
                    bufs[f"firY#{st['node']}"] = y
                continue
            acc = None
            vregs = {}
            for op, text in zip(st["ops"], st["text"]):
                c = op["code"]
                m = re.match(r"\[(acc = 0\.0 \+ acc; )?acc /= ([^\]]+)\] ", text)
                with np.errstate(all="ignore"):
                    if m:                                            # folded fan-in prologue
                        if m.group(1):
                            acc = (zero + acc).astype(F)
                        acc = (acc / F(m.group(2))).astype(F)
                    if c == OP_NOP:
                        if not m:
                            assert text.startswith("acc = 0.0 + acc"), text
                            acc = (zero + acc).astype(F)
                    elif c == OP_ZERO:
                        acc = np.zeros((self.C, n), F)
                    elif c in (OP_LOADG, OP_ADDG):
                        g = bufs[st["binds"][op["g"]]]
                        acc = (zero + g).astype(F) if c == OP_LOADG else (acc + g).astype(F)
                    elif c in (OP_LOADV, OP_ADDV):
                        v = acc if op["v"] == -1 else vregs[op["v"]]
                        acc = (zero + v).astype(F) if c == OP_LOADV else (acc + v).astype(F)
                    elif c == OP_SAVEV:
                        vregs[op["v"]] = acc.copy()
                    elif c == OP_STOREG:
                        bufs[st["binds"][op["g"]]] = acc.copy()
                    elif c == OP_MODMAP:
                        mm = re.search(r"; (\w+)#(\d+)\.(\w+) control port", text)
                        nd = self.nodes[int(mm.group(2))]
                        acc = nd.param(mm.group(3), {mm.group(3): acc}, {mm.group(3): True})
                    elif c in NODE_OPS:
                        nid = self._node_of(text)
                        nd = self.nodes[nid]
                        params = {PARAM_ORDER[nd.t][k]: vregs[v] for k, v in op["pk"].items()}
                        if c in (OP_ADD, OP_MIX):
                            ins = {"a": acc, "b": vregs[op["v"]]}
                        elif c == OP_SIGGEN:
                            ins = {}
                        else:
                            ins = {"in": acc}
                        acc = self._run_node(nid, ins, params, n)
                    else:
                        raise AssertionError(f"op code {c} is not modelled: {text}")
        n_out = len(self.np.out_terms) if hasattr(self.np, "out_terms") else None
        outs = []
        t = 0
        while f"out{t}" in bufs:
            outs.append(bufs[f"out{t}"])
            t += 1
        assert n_out is None or n_out == len(outs)
        return outs
