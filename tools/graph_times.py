"""CUDA-graph replay time of the per-agent stage (backbone + shrinker + encode) and the ego stage, with and without
programmatic dependent launch (debug flag 128 = plain launches).  usage: graph_times.py [agents]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from quantv2x_b200 import _lib  # noqa: E402
from quantv2x_b200.collab_model import normalize_pairwise_tfm  # noqa: E402
from quantv2x_b200.export import attach_engines  # noqa: E402
from quantv2x_b200.synthetic import synthetic_bev, synthetic_poses  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
q, bev_delta = bench.build_calibrated_model(dev, "att", 8)
attach_engines(q, bev_delta=bev_delta, device=dev)
pipe = q.model._pipelines["m1"]
bev = torch.from_numpy(synthetic_bev(0, n)).to(dev)
aff = normalize_pairwise_tfm(torch.from_numpy(synthetic_poses(8)).float(), 80.0, 281.6, 1)[0, 0, :8].contiguous().to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(g, iters=20):
    for _ in range(3):
        g.replay()
    tot = 0.0
    for _ in range(iters):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


for flags in (0, 128):
    _lib.lib().qv2x_set_debug_flags(flags)
    g_enc, codes = pipe.capture_encode(bev)
    print(f"agents={n} flags={flags}: encode graph {timeit(g_enc):8.1f} us")
_lib.lib().qv2x_set_debug_flags(0)
