"""FP32 PyTorch definitions of the pyramid-fusion backbone, so that reference checkpoints load
unchanged (same parameter names and shapes) and PTQ calibration has a float model to observe:

* ``BasicBlock``                    -- resblock.py:19-64, the block of the agent-side ResNetBEVBackbone
* ``Bottleneck`` / ``ResNeXtStages`` -- the ResNeXt trunk ``ResNetModified(Bottleneck, groups=32, width_per_group=4)``
  with ``Bottleneck.expansion = 1`` (opencood/models/sub_modules/resblock.py:67-122, 125-235 as configured by
  opencood/models/fuse_modules/pyramid_fuse.py:69-77)
* ``PyramidFusion``                 -- pyramid_fuse.py:64-180: trunk + per-level occupancy heads + deblocks
* ``weighted_fuse_torch``           -- pyramid_fuse.py:17-62, the calibration-time (float, any device) body

Inference does not run these bodies: see quantv2x_b200.pyramid (libqv2x engines) and
quantv2x_b200.quant.quant_block.QuantPyramidFusion.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class BasicBlock(nn.Module):
    """Two 3x3 convs with the shortcut added before the last ReLU (the agent-side ResNetBEVBackbone of the pyramid
    models: opencood/models/sub_modules/resblock.py:19-64, base_bev_backbone_resnet.py:40-45)."""

    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + shortcut)


class Bottleneck(nn.Module):
    """1x1 -> grouped 3x3 (carries the stride) -> 1x1, shortcut added before the last ReLU.  ``expansion`` is 1 (the
    pyramid model's setting): the block outputs ``planes`` channels, the inner width is planes * base_width/64 * groups."""

    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=32, base_width=4):
        super().__init__()
        width = int(planes * (base_width / 64.0)) * groups
        self.conv1 = nn.Conv2d(inplanes, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride=stride, padding=1, groups=groups, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, planes * self.expansion, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return self.relu(out + shortcut)


class ResNeXtStages(nn.Module):
    """``layer{i}`` = Sequential of ``layer_nums[i]`` bottlenecks; the first one of a stage carries the stride and, when
    the shape changes, a 1x1 conv + BN shortcut.  forward returns the output of every stage."""

    def __init__(self, layer_nums, layer_strides, num_filters, inplanes=64, groups=32, width_per_group=4):
        super().__init__()
        self.layernum = len(num_filters)
        for i, (n, stride, planes) in enumerate(zip(layer_nums, layer_strides, num_filters)):
            blocks = []
            for b in range(n):
                s = stride if b == 0 else 1
                down = None
                if b == 0 and (s != 1 or inplanes != planes * Bottleneck.expansion):
                    down = nn.Sequential(nn.Conv2d(inplanes, planes * Bottleneck.expansion, 1, stride=s, bias=False),
                                         nn.BatchNorm2d(planes * Bottleneck.expansion))
                blocks.append(Bottleneck(inplanes, planes, s, down, groups, width_per_group))
                inplanes = planes * Bottleneck.expansion
            setattr(self, f"layer{i}", nn.Sequential(*blocks))

    def forward(self, x):
        feats = []
        for i in range(self.layernum):
            x = getattr(self, f"layer{i}")(x)
            feats.append(x)
        return feats


class BasicBlockStages(nn.Module):
    """``layer{i}`` = Sequential of ``layer_nums[i]`` BasicBlocks (reference resblock.ResNetModified with BasicBlock,
    resblock.py:125-227: the first block of a stage carries the stride and, when the shape changes, a 1x1 conv + BN
    shortcut).  forward returns the output of every stage."""

    def __init__(self, layer_nums, layer_strides, num_filters, inplanes=64):
        super().__init__()
        self.layernum = len(num_filters)
        for i, (n, stride, planes) in enumerate(zip(layer_nums, layer_strides, num_filters)):
            blocks = []
            for b in range(n):
                s = stride if b == 0 else 1
                down = None
                if b == 0 and (s != 1 or inplanes != planes * BasicBlock.expansion):
                    down = nn.Sequential(nn.Conv2d(inplanes, planes * BasicBlock.expansion, 1, stride=s, bias=False),
                                         nn.BatchNorm2d(planes * BasicBlock.expansion))
                blocks.append(BasicBlock(inplanes, planes, s, down))
                inplanes = planes * BasicBlock.expansion
            setattr(self, f"layer{i}", nn.Sequential(*blocks))

    def forward(self, x):
        feats = []
        for i in range(self.layernum):
            x = getattr(self, f"layer{i}")(x)
            feats.append(x)
        return feats


class ResNetBEVBackbone(nn.Module):
    """The agent-side backbone of the pyramid models (reference base_bev_backbone_resnet.py:12-118): BasicBlock stages;
    the shipped configs (``layer_nums: [3]``, no ``upsample_strides``) have no deblocks, so the output is the last
    stage's feature.  Same attribute names (``resnet.layer{i}.{j}.conv1`` ...) as the reference."""

    def __init__(self, model_cfg, input_channels=64):
        super().__init__()
        self.model_cfg = model_cfg
        if model_cfg.get("upsample_strides"):
            raise NotImplementedError("agent-side deblocks (upsample_strides) are not built: the pyramid configs of the "
                                      "hot path have none")
        nums, strides, filters = (list(model_cfg[k]) for k in ("layer_nums", "layer_strides", "num_filters"))
        self.num_levels = len(nums)
        self.resnet = BasicBlockStages(nums, strides, filters, inplanes=model_cfg.get("inplanes", input_channels))
        self.deblocks = nn.ModuleList()
        self.num_bev_features = filters[-1]

    def get_multiscale_feature(self, spatial_features):
        return self.resnet(spatial_features)

    def forward(self, spatial_features):
        x = self.resnet(spatial_features)
        return torch.cat(x, dim=1) if len(x) > 1 else x[0]


class AlignNet(nn.Module):
    """reference feature_alignnet.py:12-43; only ``core_method: identity`` (the LiDAR configs) is on this path."""

    def __init__(self, args):
        super().__init__()
        if args.get("core_method", "identity") != "identity":
            raise NotImplementedError(f"aligner {args.get('core_method')!r} is not on the B200 path (identity only)")
        self.channel_align = nn.Identity()

    def forward(self, x):
        return self.channel_align(x)


def warp_affine_simple(src, m, dsize):
    """F.affine_grid + F.grid_sample (bilinear, zeros, align_corners=False): torch_transformation_utils.py:323-332."""
    grid = F.affine_grid(m, [src.shape[0], src.shape[1], dsize[0], dsize[1]], align_corners=False).to(src)
    return F.grid_sample(src, grid, align_corners=False)


def weighted_fuse_torch(x, score, record_len, affine_matrix):
    """x [sum(N), C, H, W], score [sum(N), 1, H, W] -> [B, C, H, W]: per sample, warp features and scores into the ego
    frame, exclude agents whose warped score is exactly 0, softmax over agents, weighted sum."""
    _, _, H, W = x.shape
    outs, start = [], 0
    for b, n in enumerate(int(v) for v in record_len):
        t = affine_matrix[b][0, :n]
        feat = warp_affine_simple(x[start:start + n], t, (H, W))
        sc = warp_affine_simple(score[start:start + n], t, (H, W))
        sc = sc.masked_fill(sc == 0, float("-inf"))
        w = torch.softmax(sc, dim=0)
        w = torch.where(torch.isnan(w), torch.zeros_like(w), w)
        outs.append((feat * w).sum(dim=0))
        start += n
    return torch.stack(outs)


class PyramidFusion(nn.Module):
    """model_cfg keys as the reference's yaml: layer_nums, layer_strides, num_filters, upsample_strides,
    num_upsample_filter, inplanes, stage ('single' | 'collab'); ResNeXt trunk only (``resnext: true``)."""

    def __init__(self, model_cfg, input_channels=64):
        super().__init__()
        self.model_cfg = model_cfg
        if not model_cfg.get("resnext", True):
            raise NotImplementedError("only the ResNeXt trunk (resnext: true) of the pyramid backbone is built")
        if model_cfg.get("align_corners", False):
            raise NotImplementedError("align_corners=True is not implemented by the warp kernel")
        nums, strides, filters = (list(model_cfg[k]) for k in ("layer_nums", "layer_strides", "num_filters"))
        ups, up_filters = list(model_cfg.get("upsample_strides", [])), list(model_cfg.get("num_upsample_filter", []))
        assert len(nums) == len(strides) == len(filters) and len(ups) == len(up_filters)
        self.stage = model_cfg["stage"]
        self.align_corners = False
        self.num_levels = len(nums)
        self.resnet = ResNeXtStages(nums, strides, filters, inplanes=model_cfg.get("inplanes", input_channels))
        self.deblocks = nn.ModuleList()
        for i in range(self.num_levels):
            if ups:
                if ups[i] < 1:
                    raise NotImplementedError("down-sampling deblocks (upsample stride < 1) are not built")
                self.deblocks.append(nn.Sequential(
                    nn.ConvTranspose2d(filters[i], up_filters[i], ups[i], stride=ups[i], bias=False),
                    nn.BatchNorm2d(up_filters[i], eps=1e-3, momentum=0.01), nn.ReLU()))
            setattr(self, f"single_head_{i}", nn.Conv2d(filters[i], 1, kernel_size=1))
        if len(ups) > self.num_levels:
            raise NotImplementedError("a final deblock over the concatenated levels is not built")
        self.num_bev_features = sum(up_filters)

    def get_multiscale_feature(self, spatial_features):
        return self.resnet(spatial_features)

    def decode_multiscale_feature(self, feats):
        ups = [self.deblocks[i](feats[i]) if len(self.deblocks) > 0 else feats[i] for i in range(self.num_levels)]
        return torch.cat(ups, dim=1) if len(ups) > 1 else ups[0]

    def forward_single(self, spatial_features):
        feats = self.get_multiscale_feature(spatial_features)
        occ = [getattr(self, f"single_head_{i}")(feats[i]) for i in range(self.num_levels)]
        return self.decode_multiscale_feature(feats), occ

    def forward_collab(self, spatial_features, record_len, affine_matrix, agent_modality_list=None, cam_crop_info=None):
        if cam_crop_info:
            raise NotImplementedError("the camera crop mask is not built (LiDAR agents only)")
        feats = self.get_multiscale_feature(spatial_features)
        fused, occ = [], []
        for i in range(self.num_levels):
            o = getattr(self, f"single_head_{i}")(feats[i])
            occ.append(o)
            fused.append(weighted_fuse_torch(feats[i], torch.sigmoid(o) + 1e-4, record_len, affine_matrix))
        return self.decode_multiscale_feature(fused), occ

    def forward(self, spatial_features, record_len=None, affine_matrix=None, agent_modality_list=None,
                cam_crop_info=None):
        if self.stage == "single":
            return self.forward_single(spatial_features)
        if record_len is None or affine_matrix is None:
            raise ValueError("record_len and affine_matrix are required for forward_collab()")
        return self.forward_collab(spatial_features, record_len, affine_matrix, agent_modality_list, cam_crop_info)
