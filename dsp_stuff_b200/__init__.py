"""dsp_stuff_b200 — B200-native batch engine for simmsb/dsp-stuff's effect-node processing path.

Only what the path needs: `csrc/` (CUDA kernels + the C-ABI library, include/dspb200.h), `engine.py`
(host-side mirror of the reference's node/graph interface over ctypes), `graph.py` (the saved-graph
data model), `signals.py` (synthetic inputs and the BASELINE workloads), `shard.py` (channel
sharding across ranks).  There is no CPU fallback: creating an Engine without the built CUDA
library or without a GPU raises.
"""
from .graph import GraphSpec, NodeSpec, NODE_PORTS  # noqa: F401

__all__ = ["GraphSpec", "NodeSpec", "NODE_PORTS", "Engine", "EngineError"]


def __getattr__(name):
    if name in ("Engine", "EngineError", "load_library"):
        from . import engine

        return getattr(engine, name)
    raise AttributeError(name)
