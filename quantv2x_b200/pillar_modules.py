"""PointPillars front end in PyTorch (checkpoint-compatible definitions + calibration path):

* ``PFNLayer`` / ``PillarVFE``   opencood/models/sub_modules/pillar_vfe.py:10-155
* ``PointPillarScatter``          opencood/models/sub_modules/point_pillar_scatter.py:19-75
* ``PointPillar``                 opencood/models/heter_encoders.py:22-51

SURVEY section 8(f)-1 lists the pillar front end as a "next" row: the B200 hot path starts at the uint8 BEV
map these modules produce.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class PFNLayer(nn.Module):
    def __init__(self, in_channels, out_channels, use_norm=True, last_layer=False):
        super().__init__()
        self.last_vfe = last_layer
        self.use_norm = use_norm
        if not self.last_vfe:
            out_channels = out_channels // 2
        if self.use_norm:
            self.linear = nn.Linear(in_channels, out_channels, bias=False)
            self.norm = nn.BatchNorm1d(out_channels, eps=1e-3, momentum=0.01)
        else:
            self.linear = nn.Linear(in_channels, out_channels, bias=True)
        self.part = 50000

    def forward(self, inputs):
        x = self.linear(inputs)
        if self.use_norm:
            x = self.norm(x.permute(0, 2, 1)).permute(0, 2, 1)
        x = F.relu(x)
        x_max = torch.max(x, dim=1, keepdim=True)[0]
        if self.last_vfe:
            return x_max
        return torch.cat([x, x_max.repeat(1, inputs.shape[1], 1)], dim=2)


def augment_pillars(voxel_features, voxel_num_points, coords, voxel_size, offsets, use_absolute_xyz=True,
                    with_distance=False):
    """Point decoration shared by the float and quantized VFE: raw xyzi + offsets to the pillar's point mean
    and to the pillar centre; padded points zeroed.  [M, P, 4] -> [M, P, 10]."""
    vx, vy, vz = voxel_size
    ox, oy, oz = offsets
    mean = voxel_features[:, :, :3].sum(dim=1, keepdim=True) / voxel_num_points.type_as(voxel_features).view(-1, 1, 1)
    f_cluster = voxel_features[:, :, :3] - mean
    f_center = torch.zeros_like(voxel_features[:, :, :3])
    dt = voxel_features.dtype
    f_center[:, :, 0] = voxel_features[:, :, 0] - (coords[:, 3].to(dt).unsqueeze(1) * vx + ox)
    f_center[:, :, 1] = voxel_features[:, :, 1] - (coords[:, 2].to(dt).unsqueeze(1) * vy + oy)
    f_center[:, :, 2] = voxel_features[:, :, 2] - (coords[:, 1].to(dt).unsqueeze(1) * vz + oz)
    feats = [voxel_features if use_absolute_xyz else voxel_features[..., 3:], f_cluster, f_center]
    if with_distance:
        feats.append(torch.norm(voxel_features[:, :, :3], 2, 2, keepdim=True))
    feats = torch.cat(feats, dim=-1)
    P = feats.shape[1]
    mask = voxel_num_points.int().unsqueeze(1) > torch.arange(P, dtype=torch.int, device=feats.device).view(1, -1)
    return feats * mask.unsqueeze(-1).type_as(feats)


class PillarVFE(nn.Module):
    def __init__(self, model_cfg, num_point_features, voxel_size, point_cloud_range):
        super().__init__()
        self.model_cfg = model_cfg
        self.use_norm = model_cfg["use_norm"]
        self.with_distance = model_cfg["with_distance"]
        self.use_absolute_xyz = model_cfg["use_absolute_xyz"]
        num_point_features += 6 if self.use_absolute_xyz else 3
        if self.with_distance:
            num_point_features += 1
        self.num_filters = model_cfg["num_filters"]
        filters = [num_point_features] + list(self.num_filters)
        self.pfn_layers = nn.ModuleList(
            PFNLayer(filters[i], filters[i + 1], self.use_norm, last_layer=(i >= len(filters) - 2))
            for i in range(len(filters) - 1))
        self.voxel_x, self.voxel_y, self.voxel_z = voxel_size
        self.x_offset = self.voxel_x / 2 + point_cloud_range[0]
        self.y_offset = self.voxel_y / 2 + point_cloud_range[1]
        self.z_offset = self.voxel_z / 2 + point_cloud_range[2]

    def forward(self, batch_dict):
        feats = augment_pillars(batch_dict["voxel_features"], batch_dict["voxel_num_points"],
                                batch_dict["voxel_coords"], (self.voxel_x, self.voxel_y, self.voxel_z),
                                (self.x_offset, self.y_offset, self.z_offset), self.use_absolute_xyz,
                                self.with_distance)
        for pfn in self.pfn_layers:
            feats = pfn(feats)
        batch_dict["pillar_features"] = feats.squeeze()
        return batch_dict


class PointPillarScatter(nn.Module):
    def __init__(self, model_cfg):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_bev_features = model_cfg["num_features"]
        self.nx, self.ny, self.nz = [int(v) for v in model_cfg["grid_size"]]
        assert self.nz == 1

    def forward(self, batch_dict):
        feats, coords = batch_dict["pillar_features"], batch_dict["voxel_coords"]
        batch_size = int(coords[:, 0].max().item()) + 1
        out = feats.new_zeros((batch_size, self.num_bev_features, self.ny * self.nx))
        idx = (coords[:, 1] + coords[:, 2] * self.nx + coords[:, 3]).long()
        out[coords[:, 0].long(), :, idx] = feats
        batch_dict["spatial_features"] = out.view(batch_size, self.num_bev_features * self.nz, self.ny, self.nx)
        return batch_dict


class PointPillar(nn.Module):
    def __init__(self, args):
        super().__init__()
        grid = (np.array(args["lidar_range"][3:6]) - np.array(args["lidar_range"][0:3])) / np.array(args["voxel_size"])
        args["point_pillar_scatter"]["grid_size"] = np.round(grid).astype(np.int64)
        self.pillar_vfe = PillarVFE(args["pillar_vfe"], num_point_features=4, voxel_size=args["voxel_size"],
                                    point_cloud_range=args["lidar_range"])
        self.scatter = PointPillarScatter(args["point_pillar_scatter"])

    def forward(self, data_dict, modality_name):
        inp = data_dict[f"inputs_{modality_name}"]
        batch_dict = {"voxel_features": inp["voxel_features"], "voxel_coords": inp["voxel_coords"],
                      "voxel_num_points": inp["voxel_num_points"]}
        batch_dict = self.scatter(self.pillar_vfe(batch_dict))
        return batch_dict["spatial_features"]
