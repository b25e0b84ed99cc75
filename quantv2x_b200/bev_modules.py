"""PyTorch definitions of the FP32 modules the quantized path wraps.  They exist so that (a) reference
checkpoints load unchanged (identical parameter names / shapes) and (b) PTQ calibration has a float
model to observe.  Structure follows the reference so state_dicts are interchangeable:

* ``BaseBEVBackbone``  -- opencood/models/sub_modules/base_bev_backbone.py:6-119
* ``DoubleConv`` / ``DownsampleConv`` -- opencood/models/sub_modules/downsample_conv.py:7-49
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


def _conv_bn_relu(cin, cout, **kw):
    return [nn.Conv2d(cin, cout, kernel_size=3, bias=False, **kw), nn.BatchNorm2d(cout, eps=1e-3, momentum=0.01),
            nn.ReLU()]


class BaseBEVBackbone(nn.Module):
    """Multi-stage 2-D BEV CNN: per stage ZeroPad2d(1) + strided 3x3 conv + n 3x3 convs, an up-sampling
    (transposed conv, kernel == stride) "deblock" per stage, channel concat of the deblock outputs."""

    def __init__(self, model_cfg, input_channels):
        super().__init__()
        self.model_cfg = model_cfg
        layer_nums = list(model_cfg.get("layer_nums", []))
        layer_strides = list(model_cfg.get("layer_strides", []))
        num_filters = list(model_cfg.get("num_filters", []))
        assert len(layer_nums) == len(layer_strides) == len(num_filters)
        upsample_strides = list(model_cfg.get("upsample_strides", []))
        num_upsample_filters = list(model_cfg.get("num_upsample_filter", []))
        assert len(upsample_strides) == len(num_upsample_filters)

        self.num_levels = len(layer_nums)
        c_in_list = [input_channels, *num_filters[:-1]]
        self.blocks = nn.ModuleList()
        self.deblocks = nn.ModuleList()
        for idx in range(self.num_levels):
            layers = [nn.ZeroPad2d(1)] + _conv_bn_relu(c_in_list[idx], num_filters[idx], stride=layer_strides[idx],
                                                       padding=0)
            for _ in range(layer_nums[idx]):
                layers += _conv_bn_relu(num_filters[idx], num_filters[idx], padding=1)
            self.blocks.append(nn.Sequential(*layers))
            if upsample_strides:
                s = upsample_strides[idx]
                if s >= 1:
                    up = nn.ConvTranspose2d(num_filters[idx], num_upsample_filters[idx], s, stride=s, bias=False)
                else:
                    s = int(np.round(1 / s))
                    up = nn.Conv2d(num_filters[idx], num_upsample_filters[idx], s, stride=s, bias=False)
                self.deblocks.append(nn.Sequential(up, nn.BatchNorm2d(num_upsample_filters[idx], eps=1e-3,
                                                                       momentum=0.01), nn.ReLU()))
        c_in = sum(num_upsample_filters)
        if len(upsample_strides) > self.num_levels:
            self.deblocks.append(nn.Sequential(
                nn.ConvTranspose2d(c_in, c_in, upsample_strides[-1], stride=upsample_strides[-1], bias=False),
                nn.BatchNorm2d(c_in, eps=1e-3, momentum=0.01), nn.ReLU()))
        self.num_bev_features = c_in

    def get_multiscale_feature(self, spatial_features):
        feats, x = [], spatial_features
        for blk in self.blocks:
            x = blk(x)
            feats.append(x)
        return feats

    def decode_multiscale_feature(self, x):
        ups = [self.deblocks[i](x[i]) if len(self.deblocks) > 0 else x[i] for i in range(self.num_levels)]
        out = torch.cat(ups, dim=1) if len(ups) > 1 else ups[0]
        if len(self.deblocks) > self.num_levels:
            out = self.deblocks[-1](out)
        return out

    def forward(self, x):
        return self.decode_multiscale_feature(self.get_multiscale_feature(x))


class DoubleConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1),
            nn.ReLU(inplace=True))

    def forward(self, x):
        return self.double_conv(x)


class DownsampleConv(nn.Module):
    """The "shrinker": a stack of DoubleConv blocks configured by kernal_size / dim / stride / padding lists
    (the misspelt key is the reference's yaml schema)."""

    def __init__(self, config):
        super().__init__()
        self.layers = nn.ModuleList()
        cin = config["input_dim"]
        for k, dim, s, p in zip(config["kernal_size"], config["dim"], config["stride"], config["padding"]):
            self.layers.append(DoubleConv(cin, dim, kernel_size=k, stride=s, padding=p))
            cin = dim

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x
