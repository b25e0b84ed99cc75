"""Frame time of the pyramid MODEL DRIVER (HeterPyramidCollabCodebookMC on libqv2x) at the V2X-Real shape: N agents,
6000 pillars each, 200 x 704 BEV -> agent backbone -> codebook (C = 64) | decode -> ResNeXt pyramid [3, 5, 8] over all
agents -> deblocks -> shrink conv -> heads.  CUDA events, eager launches (no plan / graph yet).
usage: python tools/prof_pyramid_model.py [agents] [out.json]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantv2x_b200 import _lib, yaml_utils  # noqa: E402
from quantv2x_b200.pyramid_model import attach_pyramid_engines  # noqa: E402
from quantv2x_b200.quant import QuantModel, set_weight_quantize_params  # noqa: E402
from quantv2x_b200.synthetic import seeded_init, seeded_init_codebook, synthetic_pillars, synthetic_poses  # noqa: E402

pos = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(pos[0]) if pos else 8
out_path = pos[1] if len(pos) > 1 else None
dev = torch.device("cuda:0")
here = os.path.dirname(os.path.abspath(yaml_utils.__file__))
hy = yaml_utils.load_yaml(os.path.join(here, "hypes_yaml/v2x_real/Codebook/Pyramid/lidar_pyramid_stage3.yaml"))
model = yaml_utils.create_model(hy).eval()
seeded_init(model, 1234)
seeded_init_codebook(model.codebook, 4321)
wq = dict(n_bits=8, channel_wise=True, scale_method="minmax")
aq = dict(n_bits=8, channel_wise=False, scale_method="minmax", leaf_param=True, prob=1.0)
q = QuantModel(model, wq, aq).eval()
q.disable_network_output_quantization()
q.to(dev)
set_weight_quantize_params(q)
enc = hy["model"]["args"]["m1"]["encoder_args"]


def frame(seed, agents):
    vf, vc, vn = synthetic_pillars(seed, agents, enc["lidar_range"], enc["voxel_size"], 6000)
    return {"inputs_m1": {"voxel_features": torch.from_numpy(vf).to(dev), "voxel_coords": torch.from_numpy(vc).to(dev),
                          "voxel_num_points": torch.from_numpy(vn).to(dev)},
            "agent_modality_list": ["m1"] * agents,
            "pairwise_t_matrix": torch.from_numpy(synthetic_poses(agents)).float(),
            "record_len": torch.tensor([agents])}


mods = [m for m in q.modules() if hasattr(m, "act_quantizer")]
q.set_quant_state(True, True)
for m in mods:
    m.act_quantizer.set_inited(False)
with torch.no_grad():
    q.model.calibration_forward(frame(99, 2))           # offline calibration: the torch body, two agents
for m in mods:
    m.act_quantizer.set_inited(True)
attach_pyramid_engines(q, device=dev)
data = frame(0, n)
mdl = q.model


def timed(fn, iters=5):
    for _ in range(2):
        r = fn()
    torch.cuda.synchronize()
    l0 = _lib.lib().qv2x_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (_lib.lib().qv2x_launch_count() - l0) // iters, r


if "--once" in sys.argv:          # one frame inside an NVTX range, for `ncu --nvtx --nvtx-include "timed/"`
    for _ in range(2):
        codes, _, info = mdl.encode_features(data)
        mdl.decode_features(codes, info)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("timed")
    codes, _, info = mdl.encode_features(data)
    mdl.decode_features(codes, info)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    sys.exit(0)
enc_ms, enc_l, (codes, _, info) = timed(lambda: mdl.encode_features(data))
dec_ms, dec_l, out = timed(lambda: mdl.decode_features(codes, info))
codes_u8 = torch.stack([c.t() for c in codes]).to(torch.uint8).contiguous()
graph, gout = mdl.capture_decode(codes_u8, info)
graph_ms, _, _ = timed(lambda: graph.replay())
same = bool(torch.equal(gout["preds_tensor"], out["preds_tensor"]))
res = {"agents": n, "decode_features_graph_ms": graph_ms, "graph_equals_eager": same,
       "frame_ms_with_graph": enc_ms + graph_ms, "frames_per_s_with_graph": 1e3 / (enc_ms + graph_ms), "bev": [200, 704], "encode_features_ms": enc_ms, "encode_launches": int(enc_l),
       "decode_features_ms": dec_ms, "decode_launches": int(dec_l), "frame_ms": enc_ms + dec_ms,
       "frames_per_s": 1e3 / (enc_ms + dec_ms), "preds_shape": list(out["preds_tensor"].shape),
       "preds_finite": bool(torch.isfinite(out["preds_tensor"]).all()),
       "note": "decode_features_ms: eager launches from Python with torch allocations in the loop; "
               "decode_features_graph_ms: the same stage replayed from a CUDA graph (capture_decode)"}
print(json.dumps(res))
if out_path:
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
