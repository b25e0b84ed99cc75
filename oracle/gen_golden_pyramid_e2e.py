"""Generates tests/golden/e2e_pyramid.npz by running the UNMODIFIED reference pyramid model
(opencood/models/heter_pyramid_collab_codebook_mc_encdec.py wrapped in opencood/quant/quant_model.QuantModel) on a
seeded 3-agent frame at a 32 x 64 BEV, through its own deterministic ``forward_with_encdec``.

Run in the build container only (the reference does not exist on the GPU box):
    python oracle/gen_golden_pyramid_e2e.py
TEST INFRASTRUCTURE: the fixture pins the mirror model's calibration (every quantizer's delta), the agent-side feature
codes, the codebook codes and the ego-side outputs of the pyramid driver to the reference's own arithmetic.  Weights
are rebuilt from numpy seeds by the tests (quantv2x_b200.synthetic), so the fixture holds only inputs, quantizer
parameters and reference outputs.
"""
from __future__ import annotations

import contextlib
import copy
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
from opencood.hypes_yaml import yaml_utils as ref_yaml  # noqa: E402
from opencood.quant.quant_model import QuantModel  # noqa: E402
from opencood.quant.set_weight_quantize_params import set_weight_quantize_params  # noqa: E402
from opencood.tools import train_utils  # noqa: E402

from quantv2x_b200.synthetic import (seeded_init, seeded_init_codebook, synthetic_pillars,  # noqa: E402
                                     synthetic_poses)

OUT = os.path.join(ROOT, "tests", "golden")
SMALL_RANGE = [-12.8, -6.4, -3, 12.8, 6.4, 1]
WQ = dict(n_bits=8, channel_wise=True, scale_method="minmax")
AQ = dict(n_bits=8, channel_wise=False, scale_method="minmax", leaf_param=True, prob=1.0)


def f32(t):
    return np.asarray(t.detach().cpu().numpy(), dtype=np.float32)


def main():
    with contextlib.redirect_stdout(io.StringIO()):
        hy = ref_yaml.load_yaml(os.path.join(ref_shim.REF_ROOT, "opencood/hypes_yaml/v2x_real/Codebook/Pyramid",
                                             "lidar_pyramid_stage3.yaml"))
        hy["model"]["core_method"] = "heter_pyramid_collab_codebook_mc_encdec"
        args = hy["model"]["args"]
        args["lidar_range"] = list(SMALL_RANGE)
        args["m1"]["encoder_args"]["lidar_range"] = list(SMALL_RANGE)
        model = train_utils.create_model(hy).eval()
    seeded_init(model, 1234)
    seeded_init_codebook(model.codebook, 4321)
    n = 3
    vf, vc, vn = synthetic_pillars(3, n, SMALL_RANGE, [0.4, 0.4, 4], pillars=300)
    poses = synthetic_poses(n)
    data = {"inputs_m1": {"voxel_features": torch.from_numpy(vf), "voxel_coords": torch.from_numpy(vc),
                          "voxel_num_points": torch.from_numpy(vn)},
            "agent_modality_list": ["m1"] * n, "pairwise_t_matrix": torch.from_numpy(poses).float(),
            "record_len": torch.tensor([n])}
    with contextlib.redirect_stdout(io.StringIO()):
        qt = QuantModel(model, WQ, AQ).eval()
    qt.disable_network_output_quantization()
    set_weight_quantize_params(qt)
    quantizers = {name: mod.act_quantizer for name, mod in qt.model.named_modules() if hasattr(mod, "act_quantizer")}
    for q in quantizers.values():
        q.set_inited(False)
    qt.set_quant_state(True, True)
    m = qt.model
    taps = {}
    hooks = [m.backbone_m1.register_forward_hook(lambda mod, i, o: taps.__setitem__("feat", o)),
             m.encoder_m1.register_forward_hook(lambda mod, i, o: taps.__setitem__("bev", o)),
             m.pyramid_backbone.register_forward_hook(lambda mod, i, o: taps.__setitem__("cat", o[0])),
             m.shrink_conv.register_forward_hook(lambda mod, i, o: taps.__setitem__("shrink", o))]
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        m.forward_with_encdec(copy.deepcopy(data))              # calibration pass: every quantizer initialises
        for q in quantizers.values():
            q.set_inited(True)
        codes, _, info = m.encode_features(copy.deepcopy(data))
        out = m.decode_features(codes, info)
    for h in hooks:
        h.remove()
    used = {k: q for k, q in quantizers.items() if isinstance(q.delta, torch.Tensor) or q.delta != 1.0}
    names = sorted(used)
    deltas = np.array([float(used[k].delta) for k in names], np.float64)
    zps = np.array([float(used[k].zero_point) for k in names], np.float64)
    bev_delta = float(m.encoder_m1.pillar_vfe.pfn_layers[-1].act_quantizer.delta)
    feat_delta = float(m.backbone_m1.resnet.layer0[-1].act_quantizer.delta)
    shrink_delta = float(m.shrink_conv.layers[-1].double_conv[1].act_quantizer.delta)
    bev, feat = taps["bev"], taps["feat"]
    assert float((torch.round(bev / bev_delta) * bev_delta - bev).abs().max()) < 1e-5
    assert float((torch.round(feat / feat_delta) * feat_delta - feat).abs().max()) < 1e-4
    res = dict(voxel_features=vf, voxel_coords=vc, voxel_num_points=vn, poses=poses,
               quantizer_names=np.array(names), quantizer_deltas=deltas, quantizer_zero_points=zps,
               bev_delta=np.float64(bev_delta), feat_delta=np.float64(feat_delta),
               shrink_delta=np.float64(shrink_delta),
               bev_codes=torch.round(bev / bev_delta).numpy().astype(np.uint8),
               feat_codes=torch.round(feat / feat_delta).numpy().astype(np.uint8),
               codes=np.stack([c.numpy() for c in codes]).astype(np.int16),
               cat=f32(taps["cat"]).astype(np.float16),
               shrink_codes=torch.round(taps["shrink"] / shrink_delta).numpy().astype(np.uint8),
               preds=f32(out["preds_tensor"]))
    for i, o in enumerate(out["occ_single_list"]):
        res[f"occ{i}"] = f32(o)
    np.savez_compressed(os.path.join(OUT, "e2e_pyramid.npz"), **res)
    print("e2e_pyramid.npz", os.path.getsize(os.path.join(OUT, "e2e_pyramid.npz")), "bytes;",
          {k: v.shape for k, v in res.items() if k in ("bev_codes", "feat_codes", "codes", "shrink_codes", "preds")},
          "quantizers", len(names), "bev_delta", bev_delta, "feat_delta", feat_delta,
          "preds max", float(np.abs(res["preds"]).max()), "feat nonzero", float((res["feat_codes"] > 0).mean()),
          "shrink nonzero", float((res["shrink_codes"] > 0).mean()))


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    main()
