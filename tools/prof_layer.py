"""Time / profile one quantized conv layer at a BASELINE shape.
usage: python tools/prof_layer.py [shrink1|shrink0|s0|s1|s2|d0|d1|d2] [iters] [--graph] [--debug=N]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)
from _layers import build  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "shrink1"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
use_graph = "--graph" in sys.argv
dbg = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--debug=")]
if dbg:
    from quantv2x_b200 import _lib
    _lib.lib().qv2x_set_debug_flags(dbg[0])
dev = torch.device("cuda:0")
run, ops = build(which, dev)
for _ in range(10):
    run()
torch.cuda.synchronize()
if use_graph:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters):
                run()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
else:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
print(f"{which} debug={dbg} graph={use_graph}: {ms * 1e3:.1f} us/launch  {ops / ms / 1e9:.1f} TOP/s", flush=True)
