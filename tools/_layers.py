"""Shared setup of the representative layers at BASELINE shapes (4 agents) for the timing / tracing tools."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantv2x_b200.engine import QLayer, rowsum_u8  # noqa: E402
from tests.layer_cases import make_conv, make_deconv, make_input  # noqa: E402

# name: (n_img, H, W, cin, cout, groups, kind, k/stride)
CFG = {
    "shrink1": (4, 100, 352, 256, 256, 1, 0, 3),
    "shrink0": (4, 100, 352, 384, 256, 3, 0, 3),
    "s0": (4, 100, 352, 64, 64, 1, 0, 3),
    "s1": (4, 50, 176, 128, 128, 1, 0, 3),
    "s2": (4, 25, 88, 256, 256, 1, 0, 3),
    "d0": (4, 100, 352, 64, 128, 1, 1, 1),
    "d1": (4, 50, 176, 128, 128, 1, 1, 2),
    "d2": (4, 25, 88, 256, 128, 1, 1, 4),
}


def build(which, dev):
    """Returns (run, ops): run() launches the layer once on static buffers; ops = 2*MACs per launch.
    QV2X_NIMG overrides the number of agents (default 4; the bench frame has 8)."""
    n, H, W, cin, cout, groups, kind, k = CFG[which]
    n = int(os.environ.get("QV2X_NIMG", n))
    rng = np.random.default_rng(1)
    x = torch.from_numpy(make_input(rng, n, H, W, cin)).to(dev)
    if kind == 0:
        p = make_conv(rng, cin, cout, 3, 8, groups)
        layer = QLayer(kind=0, w_int=p["w_int"], w_delta=p["w_delta"], w_zp=p["w_zp"], bias=p["bias"], ksize=3,
                       stride=1, pad=1, w_bits=8, relu=True, in_delta=p["in_delta"], out_delta=p["out_delta"])
        cg = cin // groups
        rs = [rowsum_u8(x, i * cg, cg) for i in range(groups)]
        out = torch.empty((n, H, W, cout), dtype=torch.uint8, device=dev)
        ops = 2.0 * n * H * W * cout * cin * 9
        return (lambda: layer.forward(x, rowsum_in=rs, out=out)), ops
    p = make_deconv(rng, cin, cout, k)
    layer = QLayer(kind=1, w_int=p["w_int"], w_delta=p["w_delta"], w_zp=p["w_zp"], bias=p["bias"], ksize=k, stride=k,
                   pad=0, w_bits=8, relu=True, in_delta=p["in_delta"], out_delta=p["out_delta"])
    out = torch.empty((n, H * k, W * k, 384), dtype=torch.uint8, device=dev)   # a slice of the concat buffer
    ops = 2.0 * n * H * W * cout * cin * k * k
    return (lambda: layer.forward(x, out=out, out_cbase=128)), ops
