"""Time / profile one quantized conv layer at a BASELINE shape.
usage: python tools/prof_layer.py [shrink1|shrink0|s0|s1|s2] [iters] [--graph]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantv2x_b200.engine import QLayer, rowsum_u8  # noqa: E402
from tests.layer_cases import make_conv, make_input  # noqa: E402

CFG = {
    "shrink1": (4, 100, 352, 256, 256, 1),
    "shrink0": (4, 100, 352, 384, 256, 3),
    "s0": (4, 100, 352, 64, 64, 1),
    "s1": (4, 50, 176, 128, 128, 1),
    "s2": (4, 25, 88, 256, 256, 1),
}
which = sys.argv[1] if len(sys.argv) > 1 else "shrink1"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
use_graph = "--graph" in sys.argv
dbg = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--debug=")]
if dbg:
    from quantv2x_b200 import _lib
    _lib.lib().qv2x_set_debug_flags(dbg[0])
n, H, W, cin, cout, groups = CFG[which]
dev = torch.device("cuda:0")
rng = np.random.default_rng(1)
p = make_conv(rng, cin, cout, 3, 8, groups)
x = torch.from_numpy(make_input(rng, n, H, W, cin)).to(dev)
layer = QLayer(kind=0, w_int=p["w_int"], w_delta=p["w_delta"], w_zp=p["w_zp"], bias=p["bias"], ksize=3, stride=1,
               pad=1, w_bits=8, relu=True, in_delta=p["in_delta"], out_delta=p["out_delta"])
cg = cin // groups
rs = [rowsum_u8(x, i * cg, cg) for i in range(groups)]
out = torch.empty((n, H, W, cout), dtype=torch.uint8, device=dev)
for _ in range(10):
    layer.forward(x, rowsum_in=rs, out=out)
torch.cuda.synchronize()
ops = 2.0 * n * H * W * cout * cin * 9
if use_graph:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters):
                layer.forward(x, rowsum_in=rs, out=out)
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
        layer.forward(x, rowsum_in=rs, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
print(f"{which} debug={dbg} graph={use_graph}: {ms * 1e3:.1f} us/launch  {ops / ms / 1e9:.1f} TOP/s", flush=True)
