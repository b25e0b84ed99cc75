"""``HeterBaselineCollabCodebookMC`` (+ ``...EncDec``) -- mirror of the reference V2X-Real model
(opencood/models/heter_model_baseline_mc.py:28-278, heter_baseline_collab_codebook_mc.py:30-180) for the
LiDAR PointPillar modality, with the explicit ``encode_features / decode_features / forward_with_encdec``
split the reference ships for its pyramid model (heter_pyramid_collab_codebook_mc_encdec.py:33-208).

Same constructor argument (the yaml ``model.args`` dict), same attribute names (state_dict compatible), same
input dict and output dict.  The forward is the deterministic encode -> decode path; once the model has been
wrapped in ``QuantModel``, calibrated and ``attach_engines`` has run, it executes entirely in libqv2x.
"""
from __future__ import annotations

from collections import Counter

import torch
import torch.nn as nn

from .bev_modules import BaseBEVBackbone, DownsampleConv
from .codebook import UMGMQuantizer
from .fusion_modules import AttFusion, MaxFusion
from .pillar_modules import PointPillar


def normalize_pairwise_tfm(pairwise_t_matrix, H, W, discrete_ratio, downsample_rate=1):
    """[B, L, L, 4, 4] -> [B, L, L, 2, 3] normalized affine for affine_grid (reference
    opencood/utils/transformation_utils.py:68-92; H, W are metres on this path)."""
    aff = pairwise_t_matrix[:, :, :, [0, 1], :][:, :, :, :, [0, 1, 3]].clone()
    aff[..., 0, 1] = aff[..., 0, 1] * H / W
    aff[..., 1, 0] = aff[..., 1, 0] * W / H
    aff[..., 0, 2] = aff[..., 0, 2] / (downsample_rate * discrete_ratio * W) * 2
    aff[..., 1, 2] = aff[..., 1, 2] / (downsample_rate * discrete_ratio * H) * 2
    return aff


class HeterBaselineCollabCodebookMC(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.fusion_method = args["fusion_method"]
        self.modality_name_list = [k for k in args.keys() if k.startswith("m") and k[1:].isdigit()]
        self.ego_modality = args.get("ego_modality", "m1")
        self.cav_range = args["lidar_range"]
        self.sensor_type_dict = {}
        for name in self.modality_name_list:
            setting = args[name]
            self.sensor_type_dict[name] = setting["sensor_type"]
            if setting["core_method"].replace("_", "").lower() != "pointpillar":
                raise NotImplementedError("only the point_pillar LiDAR encoder is on the B200 path")
            setattr(self, f"encoder_{name}", PointPillar(setting["encoder_args"]))
            setattr(self, f"backbone_{name}", BaseBEVBackbone(setting["backbone_args"],
                                                              setting["backbone_args"].get("inplanes", 64)))
            setattr(self, f"shrinker_{name}", DownsampleConv(setting["shrink_header"]))
        self.H = self.cav_range[4] - self.cav_range[1]
        self.W = self.cav_range[3] - self.cav_range[0]
        self.fake_voxel_size = 1
        self.supervise_single = bool(args.get("supervise_single", False))
        if self.supervise_single:   # kept for checkpoint compatibility; unused at inference
            c = args["in_head_single"]
            self.cls_head_single = nn.Conv2d(c, args["anchor_number"], kernel_size=1)
            self.reg_head_single = nn.Conv2d(c, args["anchor_number"] * 7, kernel_size=1)
            self.dir_head_single = nn.Conv2d(c, args["anchor_number"] * args["dir_args"]["num_bins"], kernel_size=1)
        if self.fusion_method == "max":
            self.fusion_net = MaxFusion()
        elif self.fusion_method == "att":
            self.fusion_net = AttFusion(args["att"]["feat_dim"])
        else:
            raise NotImplementedError(f"fusion_method {self.fusion_method!r} is not on the B200 path (max | att)")
        self.shrink_flag = "shrink_header" in args
        if self.shrink_flag:
            self.shrink_conv = DownsampleConv(args["shrink_header"])
        nc, na = args["num_class"], args["anchor_number"]
        self.cls_head = nn.Conv2d(args["in_head"], na * nc * nc, kernel_size=1)
        self.reg_head = nn.Conv2d(args["in_head"], 7 * na * nc, kernel_size=1)
        self.dir_head = nn.Conv2d(args["in_head"], args["dir_args"]["num_bins"] * na * nc, kernel_size=1)
        self.channel = 256
        if "codebook" in args:
            self.seg_num = args["codebook"]["seg_num"]
            self.dict_size = [args["codebook"]["dict_size"]] * 3
        else:
            self.seg_num, self.dict_size = 2, [256] * 3
        self.p_rate = 0.0
        lin = lambda: nn.Linear(self.channel, self.channel)  # noqa: E731
        self.codebook = UMGMQuantizer(self.channel, self.seg_num, self.dict_size, self.p_rate,
                                      {k: lin for k in ("latentStageEncoder", "quantizationHead", "latentHead",
                                                        "restoreHead", "dequantizationHead", "sideHead")})
        self._pipelines = {}        # modality -> quantv2x_b200.pipeline.CollabPipeline (set by attach_engines)

    # ------------------------------------------------------------------ float (calibration) path pieces
    def _features_float(self, data_dict):
        agent_modality_list = data_dict["agent_modality_list"]
        count = Counter(agent_modality_list)
        feats = {}
        for name in self.modality_name_list:
            if name not in count:
                continue
            f = getattr(self, f"encoder_{name}")(data_dict, name)
            f = getattr(self, f"backbone_{name}")(f)
            feats[name] = getattr(self, f"shrinker_{name}")(f)
        idx = {n: 0 for n in self.modality_name_list}
        out = []
        for name in agent_modality_list:
            out.append(feats[name][idx[name]])
            idx[name] += 1
        return torch.stack(out)

    def calibration_forward(self, data_dict):
        """Float/fake-quant forward up to the shrinker output: what PTQ calibration needs to observe.
        (The codebook, fusion and heads carry no activation quantizers: quant_block.py:1599-1615,
        quant_model.py:129-136.)"""
        return self._features_float(data_dict)

    # ------------------------------------------------------------------ inference (libqv2x)
    def _pipeline(self, name="m1"):
        if name not in self._pipelines:
            raise RuntimeError("no libqv2x engines attached: wrap the model in QuantModel, calibrate, then call "
                               "quantv2x_b200.collab_model.attach_engines(qmodel) (there is no CPU fallback)")
        return self._pipelines[name]

    def encode_features(self, data_dict):
        agent_modality_list = data_dict["agent_modality_list"]
        if set(agent_modality_list) != {"m1"}:
            raise NotImplementedError("single-modality (m1) frames only")
        affine_matrix = normalize_pairwise_tfm(data_dict["pairwise_t_matrix"], self.H, self.W, self.fake_voxel_size)
        record_len = data_dict["record_len"]
        pipe = self._pipeline("m1")
        inp = data_dict["inputs_m1"]
        if "bev_u8" in inp:                                            # BEV-level callers: uint8 [N, H, W, 64]
            bev_u8 = pipe.bev_from_inputs(data_dict)
            n = bev_u8.shape[0]
            codes = pipe.encode_agents(bev_u8)                         # uint8 [levels, m, N*hw]
        else:                                                          # the reference's pillar-level dict
            n = len(agent_modality_list)
            dev = pipe.device
            codes = pipe.encode_pillars(inp["voxel_features"].to(dev), inp["voxel_coords"].to(dev),
                                        inp["voxel_num_points"].to(dev), n)
        other_info = {"affine_matrix": affine_matrix, "record_len": record_len,
                      "agent_modality_list": agent_modality_list,
                      "feature_shape": (n, pipe.c_feat, pipe.ho, pipe.wo)}
        code_list = [codes[l].t().long() for l in range(codes.shape[0])]   # reference layout: levels x [n*hw, m]
        return code_list, agent_modality_list, other_info

    def decode_features(self, codes, other_info):
        pipe = self._pipeline("m1")
        if isinstance(codes, (list, tuple)):
            codes = torch.stack([c.t() for c in codes]).to(torch.uint8).contiguous()
        record_len = other_info["record_len"]
        if int(record_len.numel()) != 1:
            raise NotImplementedError("batch size 1 at inference (as in the reference's test loader)")
        n = int(record_len[0])
        aff = other_info["affine_matrix"][0][0, :n].to(device=codes.device, dtype=torch.float32).contiguous()
        preds = pipe.decode_fuse_heads(codes, aff)
        nc, na = self.args["num_class"], self.args["anchor_number"]
        out = pipe.split_preds(preds, na * nc * nc, 7 * na * nc, self.args["dir_args"]["num_bins"] * na * nc)
        return dict(out)

    def forward_with_encdec(self, data_dict):
        codes, _, other_info = self.encode_features(data_dict)
        return self.decode_features(codes, other_info)

    def forward(self, data_dict):
        """Deterministic inference forward (encode -> decode).  The reference's forward() samples codes with
        Gumbel noise even in eval mode; parity is defined on the encode/decode path."""
        return self.forward_with_encdec(data_dict)


HeterBaselineCollabCodebookMCEncDec = HeterBaselineCollabCodebookMC
