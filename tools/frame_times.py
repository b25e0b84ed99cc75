"""Per-stage CUDA-event timing of one bench frame (not under a profiler).  usage: frame_times.py [agents]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from quantv2x_b200 import engine as E  # noqa: E402
from quantv2x_b200.collab_model import normalize_pairwise_tfm  # noqa: E402
from quantv2x_b200.export import attach_engines  # noqa: E402
from quantv2x_b200.synthetic import synthetic_bev, synthetic_poses  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
q, bev_delta = bench.build_calibrated_model(dev, "att", 8)
attach_engines(q, bev_delta=bev_delta, device=dev)
pipe = q.model._pipelines["m1"]
bev = torch.from_numpy(synthetic_bev(0, n)).to(dev)
aff = normalize_pairwise_tfm(torch.from_numpy(synthetic_poses(n)).float(), 80.0, 281.6, 1)[0, 0, :n].contiguous().to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


eb, gb = pipe.encode_buffers(n), pipe.ego_buffers(n)
print(f"agents={n}")
print(f"backbone+shrinker plan : {timeit(lambda: pipe.fused.forward_u8(bev, out=eb['feat'])):9.1f} us")
print(f"codebook encode        : {timeit(lambda: pipe.codebook.encode(eb['feat'], pipe.feat_delta, out=eb['codes'])):9.1f} us")
print(f"codebook decode        : {timeit(lambda: pipe.codebook.decode(eb['codes'], out=gb['feat'].view(n * pipe.hw, pipe.c_feat))):9.1f} us")
print(f"warp + att fusion      : {timeit(lambda: E.fuse(gb['feat'], aff, 'att', out=gb['fused'])):9.1f} us")
print(f"warp + max fusion      : {timeit(lambda: E.fuse(gb['feat'], aff, 'max', out=gb['fused'])):9.1f} us")
print(f"heads                  : {timeit(lambda: pipe.heads.forward(gb['fused'], out=gb['preds'])):9.1f} us")
print(f"whole frame            : {timeit(lambda: pipe.forward(bev, aff)):9.1f} us")
