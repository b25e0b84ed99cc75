"""Shared by the CPU and GPU tests of the pyramid model driver: build the mirror at the fixture's 32 x 64 BEV with the
seeded weights of oracle/gen_golden_pyramid_e2e.py and calibrate it on the fixture's frame."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL_RANGE = [-12.8, -6.4, -3, 12.8, 6.4, 1]
WQ = dict(n_bits=8, channel_wise=True, scale_method="minmax")
AQ = dict(n_bits=8, channel_wise=False, scale_method="minmax", leaf_param=True, prob=1.0)


def load_fixture():
    return np.load(os.path.join(GOLD, "e2e_pyramid.npz"))


def frame_dict(g, device="cpu"):
    n = int(g["voxel_coords"][:, 0].max()) + 1
    return {"inputs_m1": {"voxel_features": torch.from_numpy(g["voxel_features"]).to(device),
                          "voxel_coords": torch.from_numpy(g["voxel_coords"]).to(device),
                          "voxel_num_points": torch.from_numpy(g["voxel_num_points"]).to(device)},
            "agent_modality_list": ["m1"] * n, "pairwise_t_matrix": torch.from_numpy(g["poses"]).float(),
            "record_len": torch.tensor([n])}


def build_and_calibrate(g):
    """-> (QuantModel, calibration output dict of the SECOND (all quantizers initialised) torch pass)."""
    from quantv2x_b200 import yaml_utils
    from quantv2x_b200.quant import QuantModel, set_weight_quantize_params
    from quantv2x_b200.synthetic import seeded_init, seeded_init_codebook

    here = os.path.dirname(os.path.abspath(yaml_utils.__file__))
    hy = yaml_utils.load_yaml(os.path.join(here, "hypes_yaml/v2x_real/Codebook/Pyramid/lidar_pyramid_stage3.yaml"))
    hy["model"]["args"]["lidar_range"] = list(SMALL_RANGE)
    hy["model"]["args"]["m1"]["encoder_args"]["lidar_range"] = list(SMALL_RANGE)
    model = yaml_utils.create_model(hy).eval()
    seeded_init(model, 1234)
    seeded_init_codebook(model.codebook, 4321)
    qt = QuantModel(model, WQ, AQ).eval()
    qt.disable_network_output_quantization()
    set_weight_quantize_params(qt)
    mods = [m for m in qt.modules() if hasattr(m, "act_quantizer")]
    for m in mods:
        m.act_quantizer.set_inited(False)
    qt.set_quant_state(True, True)
    data = frame_dict(g)
    with torch.no_grad():
        qt.model.calibration_forward(data)
        for m in mods:
            m.act_quantizer.set_inited(True)
        out = qt.model.calibration_forward(data)
    return qt, out
