"""Run W warm-up + K frames of the bench workload (for ncu launch lists).  usage: frame_once.py [agents] [K]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from quantv2x_b200.collab_model import normalize_pairwise_tfm  # noqa: E402
from quantv2x_b200.export import attach_engines  # noqa: E402
from quantv2x_b200.synthetic import synthetic_pillars, synthetic_poses  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
q, bev_delta = bench.build_calibrated_model(dev, "att", 8)
attach_engines(q, bev_delta=bev_delta, device=dev)
pipe = q.model._pipelines["m1"]
enc_args = q.hypes["model"]["args"]["m1"]["encoder_args"]
pil = [torch.from_numpy(t).to(dev) for t in synthetic_pillars(0, n, enc_args["lidar_range"], enc_args["voxel_size"], 6000)]
frame = lambda: pipe.decode_fuse_heads(pipe.encode_pillars(*pil, n), aff)
aff = normalize_pairwise_tfm(torch.from_numpy(synthetic_poses(n)).float(), 80.0, 281.6, 1)[0, 0, :n].contiguous().to(dev)
for _ in range(2):
    frame()
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("timed")
for _ in range(K):
    frame()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done")
