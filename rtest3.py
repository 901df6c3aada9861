import sys, os, ctypes
sys.path.insert(0, '.')
import torch
from dsp_stuff_b200 import GraphSpec, signals as S
from dsp_stuff_b200.engine import Engine, load_library
L = load_library()
def run(spec, C, n, tag):
    e = Engine(C, block=1024, max_samples=n)
    spec.apply(e)
    x = torch.from_numpy(S.noise(C, n)).cuda(); y = torch.empty_like(x)
    for _ in range(3): e.process_device([x],[y],n)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 8)()
    L.dspb_debug_ws_timing(buf, 1)
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): e.process_device([x],[y],n)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)/10
    L.dspb_debug_ws_timing(buf, 1)
    tiles = 10 * 64
    names = ["R wait FULL", "R loop", "E pre", "E wait DONE", "E post", "E top barrier", "pre: ops before rec", "pre: rec prologue+edge barrier"]
    print(f"{tag}: {C*n/ms/1e6:.1f} Gs/s {ms:.4f} ms; per tile cycles: " + ", ".join(f"{nm}={buf[i]/tiles:.0f}" for i, nm in enumerate(names)))
C, n = 4096, 16384
run(S.config3(), C, n, "config3")
g = GraphSpec().node(10,"input").node(11,"output").node(0,"biquad", **S.rbj_biquad("lp",1000.0)).link(10,"out",0,"in").link(0,"out",11,"in")
run(g, C, n, "in->biquad->out")
g = GraphSpec().node(10,"input").node(11,"output").node(0,"gain", level=2.0).link(10,"out",0,"in").link(0,"out",11,"in")
run(g, C, n, "gain only")
