"""Time the GPU detection post-processing (qv2x_postprocess_*) on a bench-like frame: 72 head maps of 100 x 352,
threshold set so that ~300 anchors survive.  usage: python tools/prof_postprocess.py [candidates]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantv2x_b200 import yaml_utils  # noqa: E402
from quantv2x_b200.postprocess import PostProcessor  # noqa: E402

ncand = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = torch.device("cuda:0")
hypes = yaml_utils.load_yaml(yaml_utils.default_config("att"))
torch.manual_seed(0)
p = (torch.randn((72, 35200), device=dev) * 0.5).contiguous()
sc = torch.sigmoid(p[:18].float().flatten())
thr = float(torch.topk(sc, ncand).values[-1].item())
e = PostProcessor(hypes, (704, 200), score_threshold=thr).engine
outs = e.alloc_outputs(dev)
for _ in range(3):
    e.forward_into(p, outs)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("timed")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    e.forward_into(p, outs)
e1.record()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print(f"post-processing, {ncand} candidates: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per frame (4 kernels, back to back); "
      f"boxes kept {int(outs[4][0].item())}, candidates {int(outs[4][1].item())}")
