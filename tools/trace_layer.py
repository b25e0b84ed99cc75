"""Per-role timeline of the igemm kernel (qv2x_debug_trace): cycles per tile spent in each role.
usage: python tools/trace_layer.py [shrink1|shrink0|s0|s1|s2] [cta]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)
from ctypes import c_void_p  # noqa: E402

from _layers import build  # noqa: E402
from quantv2x_b200 import _lib  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "s0"
cta = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device("cuda:0")
run, _ = build(which, dev)
for _ in range(5):
    run()
torch.cuda.synchronize()
buf = torch.zeros((148, 32, 16), dtype=torch.int64, device=dev)
_lib.lib().qv2x_debug_trace(c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
_lib.lib().qv2x_debug_trace(c_void_p(0))
tr = buf.cpu().numpy()[cta]
t0 = tr[0][:14][tr[0][:14] > 0].min()
names = ["P:start", "P:issued", "M:start", "M:slot", "M:full0", "M:issued", "E0:start", "E0:begun", "E0:tfull",
         "E0:chunks", "E0:end", "E1:start", "E1:tfull", "E1:end"]
print(f"{which}: CTA {cta}, cycles relative to the first stamp")
print("tile " + " ".join(f"{nm:>9s}" for nm in names))
for i in range(32):
    if tr[i, 6] == 0:
        break
    print(f"{i:4d} " + " ".join(f"{(tr[i, k] - t0) if tr[i, k] else -1:9d}" for k in range(len(names))) +
          f"   Pwait {tr[i, 14]:6d} Mwait {tr[i, 15]:6d}")
