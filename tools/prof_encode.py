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

if "--trace" in sys.argv:
    from ctypes import c_void_p

    cta = 40
    buf = torch.zeros((148, 32, 16), dtype=torch.int64, device=dev)
    _lib.lib().qv2x_debug_trace(c_void_p(buf.data_ptr()))
    _lib.lib().qv2x_set_debug_flags((dbg[0] if dbg else 0) | 32)
    eng.encode(q, 0.173, out=out)
    torch.cuda.synchronize()
    _lib.lib().qv2x_debug_trace(c_void_p(0))
    tr = buf.cpu().numpy()[cta]
    t0 = tr[0][:14][tr[0][:14] > 0].min()
    names = ["P:start", "P:issued", "M:start", "M:slot", "M:full0", "M:issued", "E0:start", "E0:begun", "E0:tfull",
             "E0:chunk0", "E0:end", "E0:lvl0end", "E0:tfull1", "E0:chunk1"]
    print("tile " + " ".join(f"{nm:>9s}" for nm in names))
    for i in range(20):
        if tr[i, 6] == 0:
            break
        print(f"{i:4d} " + " ".join(f"{(tr[i, k] - t0) if tr[i, k] else -1:9d}" for k in range(len(names))) +
              f"   Pwait {tr[i, 14]:6d} Mwait {tr[i, 15]:6d}")
