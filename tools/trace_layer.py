"""Per-role timeline of the igemm kernel (qv2x_debug_trace): cycles per tile spent in each role.
usage: python tools/trace_layer.py [shrink1|shrink0|s0|s1|s2] [cta]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctypes import c_void_p  # noqa: E402

from quantv2x_b200 import _lib  # noqa: E402
from quantv2x_b200.engine import QLayer, rowsum_u8  # noqa: E402
from tests.layer_cases import make_conv, make_input  # noqa: E402

CFG = {
    "shrink1": (4, 100, 352, 256, 256, 1),
    "shrink0": (4, 100, 352, 384, 256, 3),
    "s0": (4, 100, 352, 64, 64, 1),
    "s1": (4, 50, 176, 128, 128, 1),
    "s2": (4, 25, 88, 256, 256, 1),
}
which = sys.argv[1] if len(sys.argv) > 1 else "s0"
cta = int(sys.argv[2]) if len(sys.argv) > 2 else 40
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
for _ in range(5):
    layer.forward(x, rowsum_in=rs, out=out)
torch.cuda.synchronize()
buf = torch.zeros((148, 32, 16), dtype=torch.int64, device=dev)
_lib.lib().qv2x_debug_trace(c_void_p(buf.data_ptr()))
layer.forward(x, rowsum_in=rs, out=out)
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
