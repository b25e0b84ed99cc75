"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the build container only (the reference does not exist on the GPU box):
    python oracle/gen_golden.py
The fixtures pin the oracle (and through it the CUDA path) to the reference's own arithmetic; the reference
ships no tests or golden vectors of its own.  Weights are rebuilt from numpy seeds by the tests
(quantv2x_b200.synthetic), so the fixtures hold only inputs, quantizer parameters and reference outputs.
"""
from __future__ import annotations

import contextlib
import copy
import io
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
from opencood.hypes_yaml import yaml_utils as ref_yaml  # noqa: E402
from opencood.models.fuse_modules.fusion_in_one import AttFusion, MaxFusion  # noqa: E402
from opencood.models.sub_modules.codebook import UMGMQuantizer  # noqa: E402
from opencood.quant.quant_layer import QuantModule  # noqa: E402
from opencood.quant.quant_model import QuantModel  # noqa: E402
from opencood.quant.set_weight_quantize_params import set_weight_quantize_params  # noqa: E402
from opencood.tools import train_utils  # noqa: E402
from opencood.utils.transformation_utils import normalize_pairwise_tfm  # noqa: E402

from quantv2x_b200.synthetic import (seeded_init, seeded_init_codebook, synthetic_pillars,  # noqa: E402
                                     synthetic_poses)

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
WQ = dict(n_bits=8, channel_wise=True, scale_method="minmax")
AQ = dict(n_bits=8, channel_wise=False, scale_method="minmax", leaf_param=True, prob=1.0)


def f32(t):
    return np.asarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.float32)


# ------------------------------------------------------------------------------------------ 1. single layers
LAYER_CASES = [
    # name, kind, cin, cout, k, stride, pad, zero_pad, w_bits, H, W, head
    ("conv_s1", 0, 64, 64, 3, 1, 1, 0, 8, 12, 20, False),
    ("conv_s2_zeropad", 0, 64, 128, 3, 2, 0, 1, 8, 12, 20, False),
    ("conv_w4", 0, 128, 128, 3, 1, 1, 0, 4, 10, 12, False),
    ("deconv_s2", 1, 128, 128, 2, 2, 0, 0, 8, 6, 10, False),
    ("deconv_s4", 1, 256, 128, 4, 4, 0, 0, 8, 3, 5, False),
    ("head_1x1", 0, 256, 18, 1, 1, 0, 0, 8, 6, 10, True),
]


def gen_layers():
    out = {}
    for idx, (name, kind, cin, cout, k, s, p, zp_pad, w_bits, H, W, head) in enumerate(LAYER_CASES):
        rng = np.random.default_rng(100 + idx)
        if kind == 0:
            mod = nn.Conv2d(cin, cout, k, stride=s, padding=p, bias=True)
            fan_in = cin * k * k
        else:
            mod = nn.ConvTranspose2d(cin, cout, k, stride=s, bias=True)
            fan_in = cin
        with torch.no_grad():
            mod.weight.copy_(torch.from_numpy(rng.normal(0, np.sqrt(2.0 / fan_in), size=tuple(mod.weight.shape)).astype(np.float32)))
            mod.bias.copy_(torch.from_numpy(rng.uniform(-0.1, 0.1, size=cout).astype(np.float32)))
        qm = QuantModule(mod, dict(WQ, n_bits=w_bits), AQ, disable_act_quant=head)
        if not head:
            qm.activation_function = nn.ReLU()
        in_delta = np.float32(0.11)
        q_in = rng.integers(0, 256, size=(2, cin, H, W)).astype(np.uint8)
        q_in[rng.random(q_in.shape) > 0.5] = 0
        x = torch.from_numpy(q_in.astype(np.float32) * in_delta)
        if zp_pad:
            x = nn.ZeroPad2d(zp_pad)(x)
        qm.set_quant_state(True, True)
        qm.weight_quantizer.set_inited(False)
        qm.act_quantizer.set_inited(False)
        with torch.no_grad():
            qm(x)
            qm.weight_quantizer.set_inited(True)
            qm.act_quantizer.set_inited(True)
            y = qm(x)
        out[f"{name}.q_in"] = q_in
        out[f"{name}.in_delta"] = in_delta
        out[f"{name}.w_delta"] = f32(qm.weight_quantizer.delta).reshape(-1)
        out[f"{name}.w_zp"] = f32(qm.weight_quantizer.zero_point).reshape(-1)
        if head:
            out[f"{name}.out"] = f32(y)
        else:
            d, z = float(qm.act_quantizer.delta), float(qm.act_quantizer.zero_point)
            out[f"{name}.act_delta"] = np.float32(d)
            out[f"{name}.act_zp"] = np.float32(z)
            codes = torch.round(y / d + z)
            assert float((codes * d - z * d - y).abs().max()) < 1e-4 * d
            out[f"{name}.out_codes"] = codes.numpy().astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "quant_layers.npz"), **out)
    print("quant_layers.npz", len(out), "arrays")


# ------------------------------------------------------------------------------------------ 2. codebook
def gen_codebook():
    out = {}
    for name, C, m, ks, rows in [("c64_m2_k32", 64, 2, [32] * 3, 512), ("c256_m1_k128", 256, 1, [128] * 3, 256)]:
        q = UMGMQuantizer(C, m, ks, 0.0, {n: (lambda: nn.Linear(C, C)) for n in
                                          ["latentStageEncoder", "quantizationHead", "latentHead",
                                           "dequantizationHead", "sideHead", "restoreHead"]}).eval()
        seeded_init_codebook(q, 77)
        rng = np.random.default_rng(5)
        xq = rng.integers(0, 256, size=(rows, C)).astype(np.uint8)
        xq[rng.random(xq.shape) > 0.5] = 0
        delta = np.float32(0.173)
        with torch.no_grad():
            codes = q.encode(torch.from_numpy(xq.astype(np.float32) * delta))
            dec = q.decode(codes)
        out[f"{name}.xq"] = xq
        out[f"{name}.delta"] = delta
        out[f"{name}.codes"] = np.stack([c.numpy() for c in codes]).astype(np.int16)      # [L, rows, m]
        out[f"{name}.decoded"] = f32(dec)
    np.savez_compressed(os.path.join(OUT, "codebook.npz"), **out)
    print("codebook.npz")


# ------------------------------------------------------------------------------------------ 3. fusion
def gen_fusion():
    out = {}
    rng = np.random.default_rng(9)
    N, C, H, W = 3, 16, 12, 20
    feat = (rng.random((N, C, H, W)) * 4 * (rng.random((N, C, H, W)) > 0.4)).astype(np.float32)
    poses = synthetic_poses(N)
    aff = normalize_pairwise_tfm(torch.from_numpy(poses).float(), 80.0, 281.6, 1)
    rl = torch.tensor([N])
    with torch.no_grad():
        out["att"] = f32(AttFusion(C)(torch.from_numpy(feat), rl, aff)[0])
        out["max"] = f32(MaxFusion()(torch.from_numpy(feat), rl, aff)[0])
    out["feat"] = feat
    out["poses"] = poses
    out["affine"] = f32(aff)
    np.savez_compressed(os.path.join(OUT, "fusion.npz"), **out)
    print("fusion.npz")


# ------------------------------------------------------------------------------------------ 4. end to end (small range)
SMALL_RANGE = [-12.8, -6.4, -3, 12.8, 6.4, 1]      # 64 x 32 BEV cells -> 32 x 16 feature map


def gen_e2e():
    for fusion in ("att", "max"):
        sub = "Attfuse/lidar_attfuse_stage3.yaml" if fusion == "att" else "Attfuse/lidar_attfuse_stage3.yaml"
        with contextlib.redirect_stdout(io.StringIO()):
            hy = ref_yaml.load_yaml(os.path.join(ref_shim.REF_ROOT, "opencood/hypes_yaml/v2x_real/Codebook", sub))
            args = hy["model"]["args"]
            args["lidar_range"] = list(SMALL_RANGE)
            args["m1"]["encoder_args"]["lidar_range"] = list(SMALL_RANGE)
            args["fusion_method"] = fusion
            hy["model"]["args"] = args
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
        qt = QuantModel(model, WQ, AQ).eval()
        qt.disable_network_output_quantization()
        set_weight_quantize_params(qt)
        for mod in qt.modules():
            if hasattr(mod, "act_quantizer"):
                mod.act_quantizer.set_inited(False)
        qt.set_quant_state(True, True)
        m = qt.model
        with torch.no_grad():
            with contextlib.redirect_stdout(io.StringIO()):
                qt(copy.deepcopy(data))                       # calibration forward (the stochastic forward is fine here)
            for mod in qt.modules():
                if hasattr(mod, "act_quantizer"):
                    mod.act_quantizer.set_inited(True)
            # deterministic path: encoder -> backbone -> shrinker -> codebook.encode -> decode -> fusion -> heads
            bev = m.encoder_m1(copy.deepcopy(data), "m1")
            feat = m.shrinker_m1(m.backbone_m1(bev))
            N, C, H, W = feat.shape
            flat = feat.permute(0, 2, 3, 1).contiguous().view(-1, C)
            codes = m.codebook.encode(flat)
            dec = m.codebook.decode(codes).view(N, H, W, C).permute(0, 3, 1, 2).contiguous()
            aff = normalize_pairwise_tfm(data["pairwise_t_matrix"], m.H, m.W, m.fake_voxel_size)
            fused = m.fusion_net(dec, data["record_len"], aff)
            preds = torch.cat([m.cls_head(fused), m.reg_head(fused), m.dir_head(fused)], dim=1)
        bev_delta = float(m.encoder_m1.pillar_vfe.pfn_layers[-1].act_quantizer.delta)
        feat_delta = float(m.shrinker_m1.layers[-1].double_conv[1].act_quantizer.delta)
        act_deltas = {name: float(mod.act_quantizer.delta) for name, mod in qt.model.named_modules()
                      if isinstance(mod, QuantModule) and not mod.disable_act_quant}
        out = dict(voxel_features=vf, voxel_coords=vc, voxel_num_points=vn, poses=poses,
                   bev_delta=np.float32(bev_delta), feat_delta=np.float32(feat_delta),
                   bev_codes=torch.round(bev / bev_delta).numpy().astype(np.uint8),
                   feat_codes=torch.round(feat / feat_delta).numpy().astype(np.uint8),
                   codes=np.stack([c.numpy() for c in codes]).astype(np.int16), fused=f32(fused).astype(np.float16),
                   preds=f32(preds), act_delta_names=np.array(sorted(act_deltas)),
                   act_delta_values=np.array([act_deltas[k] for k in sorted(act_deltas)], np.float32))
        assert float((torch.round(bev / bev_delta) * bev_delta - bev).abs().max()) < 1e-5
        np.savez_compressed(os.path.join(OUT, f"e2e_{fusion}.npz"), **out)
        print(f"e2e_{fusion}.npz", {k: getattr(v, "shape", None) for k, v in out.items() if k in ("bev_codes", "feat_codes", "preds")},
              "bev_delta", bev_delta, "feat_delta", feat_delta, "preds max", float(preds.abs().max()))


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    gen_layers()
    gen_codebook()
    gen_fusion()
    gen_e2e()
    print("sizes:", {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))})
