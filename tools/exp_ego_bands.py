"""Ego stage in horizontal bands: decode only the source rows a band samples, fuse the band, heads on the band -- so that
the decoded maps of a band (<= L2 size) are read back from L2 instead of HBM.  Times CUDA-graph replays against the
whole-map ego stage and checks the outputs are identical.  usage: python tools/exp_ego_bands.py [agents]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from quantv2x_b200.collab_model import normalize_pairwise_tfm  # noqa: E402
from quantv2x_b200.export import attach_engines  # noqa: E402
from quantv2x_b200.synthetic import synthetic_poses  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
q, bev_delta = bench.build_calibrated_model(dev, "att", 8)
attach_engines(q, bev_delta=bev_delta, device=dev)
pipe = q.model._pipelines["m1"]
aff = normalize_pairwise_tfm(torch.from_numpy(synthetic_poses(n)).float(), 80.0, 281.6, 1)[0, 0, :n].contiguous().to(dev)
aff_host = aff.cpu().numpy()
codes = torch.randint(0, 128, (3, 1, n * pipe.hw), dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(graph, iters=20):
    tot = 0.0
    for _ in range(iters):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


g0, ref = pipe.capture_ego(codes, aff, slot=0)
print(f"whole map: {timed(g0):7.1f} us")
ref = ref.clone()
for bands in (2, 4, 5, 10):
    if pipe.ho % bands:
        continue
    out = torch.zeros_like(ref)
    th = pipe.ho // bands

    def run(out=out, bands=bands, th=th):
        for b in range(bands):
            tile = (b * th, (b + 1) * th, 0, pipe.wo)
            pipe.decode_fuse_heads_tile_to(codes, aff, aff_host, tile, out.data_ptr(), slot=10 + bands)
        return out

    g, o = pipe._capture(run)
    t = timed(g)
    g.replay()
    torch.cuda.synchronize()
    print(f"{bands:2d} bands : {t:7.1f} us   identical: {bool(torch.equal(o, ref))}")
