"""Time the codebook encode kernel at the BASELINE shape (C=256, m=1, k=128 x 3 levels).
usage: python tools/prof_encode.py [agents] [--debug=N] [--k=128]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantv2x_b200 import _lib  # noqa: E402
from quantv2x_b200.engine import CodebookEngine  # noqa: E402
from tests.codebook_cases import make_codebook_params, make_features  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else 4
dbg = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--debug=")]
kk = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--k=")]
k = kk[0] if kk else 128
dev = torch.device("cuda:0")
cbs, heads = make_codebook_params(5, 256, 1, [k] * 3)
eng = CodebookEngine(cbs, heads)
rows = n * 35200
q = torch.from_numpy(make_features(1, rows, 256)).to(dev)
out = torch.empty((3, 1, rows), dtype=torch.uint8, device=dev)
for _ in range(3):
    eng.encode(q, 0.173, out=out)
torch.cuda.synchronize()
if dbg:
    _lib.lib().qv2x_set_debug_flags(dbg[0])
iters = 20
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    eng.encode(q, 0.173, out=out)
    with torch.cuda.graph(g, stream=s):
        for _ in range(iters):
            eng.encode(q, 0.173, out=out)
for _ in range(2):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / iters * 1e3
gmac = rows * 3 * k * 256 * 3 / 1e9
print(f"encode k={k} agents={n} debug={dbg}: {us:.1f} us/launch  ({2 * gmac / us / 1e3:.1f} TOP/s of int8 digit MMA)", flush=True)
