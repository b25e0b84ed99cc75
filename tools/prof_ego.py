"""Time the ego stage at the BASELINE shape (100 x 352 map, C=256, m=1, k x 3 levels, 72 head channels): the one-kernel
qv2x_ego_att against the three-kernel chain decode -> fuse -> heads.
usage: python tools/prof_ego.py [agents] [--k=128] [--random-codes]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantv2x_b200 import engine as E  # noqa: E402
from quantv2x_b200.collab_model import normalize_pairwise_tfm  # noqa: E402
from quantv2x_b200.synthetic import synthetic_poses  # noqa: E402
from tests.codebook_cases import make_codebook_params  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else 8
kk = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--k=")]
k = kk[0] if kk else 128
ho, wo, C = 100, 352, 256
hw = ho * wo
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
cb = E.CodebookEngine(*make_codebook_params(5, C, 1, [k] * 3))
hd = E.HeadsEngine(rng.normal(size=(72, C)).astype(np.float32) / 16, rng.normal(size=72).astype(np.float32))
eng = E.EgoAttEngine(cb, hd)
t = torch.from_numpy(synthetic_poses(n, max(n, 5)))
aff = normalize_pairwise_tfm(t, 80.0, 281.6, 1)[0, 0, :n].to(torch.float32).contiguous().to(dev)
codes = torch.from_numpy(rng.integers(0, k, size=(3, 1, n * hw), dtype=np.uint8)).to(dev)
out = torch.empty((72, hw), dtype=torch.float32, device=dev)
feat = torch.empty((n, ho, wo, C), dtype=torch.float32, device=dev)
fused = torch.empty((ho, wo, C), dtype=torch.float32, device=dev)
out2 = torch.empty((72, hw), dtype=torch.float32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def chain():
    cb.decode(codes, out=feat.view(n * hw, C))
    E.fuse(feat, aff, "att", out=fused)
    hd.forward(fused, out=out2)


def time_cold(fn, iters=10):
    ts = []
    for _ in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts[2:]))


one = time_cold(lambda: eng.forward(codes, aff, n, ho, wo, out=out))
three = time_cold(chain)
err = float((out - out2).abs().max()) / float(out2.abs().max())
alg = n * 3 * hw + 72 * hw * 4
print(f"ego stage, {n} agents, k={k}: qv2x_ego_att {one:.1f} us ({alg / one / 1e3:.1f} GB/s of {alg / 1e6:.2f} MB "
      f"algorithmic), chain decode+fuse+heads {three:.1f} us; max |diff| / max |y| = {err:.2e}")
