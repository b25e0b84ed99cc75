"""Block-level quantization wrappers -- mirror of the part of opencood/quant/quant_block.py that the
V2X-Real intermediate-fusion path uses:

* ``BaseQuantBlock``          reference quant_block.py:45-65
* ``QuantBaseBEVBackbone``    reference quant_block.py:243-335
* ``QuantDoubleConv`` / ``QuantDownsampleConv``   reference quant_block.py:552-586
* registries ``opencood_specials`` / ``specials_unquantized_names``   reference quant_block.py:1581-1615

The torch bodies are the calibration path.  With every quantizer initialised and quantization switched
on, ``forward`` hands the whole block to libqv2x (see ``attach_engine``): FP32 NCHW is converted to uint8
NHWC once at the block boundary, all layers of the block run as tcgen05 kernels, and the result is
de-quantized back to FP32 NCHW so the block stays a drop-in for callers that expect the reference's
tensors.  The model-level fast path (quantv2x_b200.collab_model) skips even those boundary copies.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..bev_modules import BaseBEVBackbone, DoubleConv, DownsampleConv
from .quant_layer import QuantModule


class BaseQuantBlock(nn.Module):
    def __init__(self):
        super().__init__()
        self.use_weight_quant = False
        self.use_act_quant = False
        self.ignore_reconstruction = False
        self.trained = False
        self._engine = None            # set by quantv2x_b200.export.attach_engines

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
        for m in self.modules():
            if isinstance(m, QuantModule):
                m.set_quant_state(weight_quant, act_quant)

    # ------------------------------------------------------------------ engine hand-off
    def quant_modules(self):
        return [m for m in self.modules() if isinstance(m, QuantModule)]

    def engine_ready(self) -> bool:
        """True when this block must run on the GPU library: quantization on and everything calibrated."""
        if not (self.use_weight_quant and self.use_act_quant):
            return False
        return all(m.weight_quantizer.inited and m.act_quantizer.inited and m.use_weight_quant and m.use_act_quant
                   for m in self.quant_modules())

    def attach_engine(self, engine):
        self._engine = engine

    def _run_engine(self, x):
        if self._engine is None:
            raise RuntimeError(
                f"{type(self).__name__}: quantized inference requested but no libqv2x engine is attached; call "
                "quantv2x_b200.export.attach_engines(model) after calibration (there is no CPU fallback)")
        return self._engine.forward_nchw(x)


class QuantBaseBEVBackbone(BaseQuantBlock):
    def __init__(self, basebevbackbone: BaseBEVBackbone, weight_quant_params={}, act_quant_params={}):
        super().__init__()
        self.num_levels = basebevbackbone.num_levels
        self.blocks = nn.ModuleList()
        self.deblocks = nn.ModuleList()
        for base_block in basebevbackbone.blocks:
            layers = list(base_block)
            wrapped = nn.Sequential(layers[0])          # ZeroPad2d stays a module
            for i in range(1, len(layers), 3):          # (conv, bn-or-identity, relu) triples
                qm = QuantModule(layers[i], weight_quant_params, act_quant_params)
                qm.norm_function = layers[i + 1]
                qm.activation_function = layers[i + 2]
                wrapped.add_module(str(len(wrapped)), qm)
            self.blocks.append(wrapped)
        for base_deblock in basebevbackbone.deblocks:
            qm = QuantModule(base_deblock[0], weight_quant_params, act_quant_params)
            qm.norm_function = base_deblock[1]
            qm.activation_function = base_deblock[2]
            self.deblocks.append(nn.Sequential(qm))
        self.num_bev_features = basebevbackbone.num_bev_features

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
        if self.engine_ready():
            return self._run_engine(x)
        return self.decode_multiscale_feature(self.get_multiscale_feature(x))


class QuantDoubleConv(BaseQuantBlock):
    def __init__(self, double_conv: DoubleConv, weight_quant_params={}, act_quant_params={}):
        super().__init__()
        seq = double_conv.double_conv
        self.double_conv = nn.Sequential(QuantModule(seq[0], weight_quant_params, act_quant_params),
                                         QuantModule(seq[2], weight_quant_params, act_quant_params))
        self.double_conv[0].activation_function = seq[1]
        self.double_conv[1].activation_function = seq[3]

    def forward(self, x):
        return self.double_conv[1](self.double_conv[0](x))


class QuantDownsampleConv(BaseQuantBlock):
    def __init__(self, downsample_conv: DownsampleConv, weight_quant_params={}, act_quant_params={}):
        super().__init__()
        self.layers = nn.ModuleList(QuantDoubleConv(layer, weight_quant_params, act_quant_params)
                                    for layer in downsample_conv.layers)

    def forward(self, x):
        if self.engine_ready():
            return self._run_engine(x)
        for layer in self.layers:
            x = layer(x)
        return x


opencood_specials = {
    BaseBEVBackbone: QuantBaseBEVBackbone,
    DownsampleConv: QuantDownsampleConv,
}

# modules the reference keeps in FP32 by attribute name (quant_block.py:1599-1615)
specials_unquantized_names = ["aligner_m1", "aligner_m2", "codebook"]


# --------------------------------------------------------------------------------------------------
# PointPillars front end (reference quant_block.py:589-741).  Produces the uint8-grid BEV map the integer
# backbone consumes: Linear(10->64, W-quant, pre-ReLU act-quant) -> ReLU -> block act-quant (zp = 0) -> max
# over the pillar's points -> scatter.
# --------------------------------------------------------------------------------------------------
from ..pillar_modules import PFNLayer, PillarVFE, PointPillar, augment_pillars  # noqa: E402
from .quant_layer import UniformAffineQuantizer  # noqa: E402
import torch.nn.functional as F  # noqa: E402


class QuantPFNLayer(BaseQuantBlock):
    def __init__(self, pfn_layer: PFNLayer, weight_quant_params={}, act_quant_params={}):
        super().__init__()
        self.last_vfe = pfn_layer.last_vfe
        self.use_norm = pfn_layer.use_norm
        self.part = pfn_layer.part
        self.linear = QuantModule(pfn_layer.linear, weight_quant_params, act_quant_params)
        if self.use_norm:
            self.linear.norm_function = pfn_layer.norm        # StraightThrough once BN is folded
        self.act_quantizer = UniformAffineQuantizer(**act_quant_params)

    def forward(self, inputs):
        x = F.relu(self.linear(inputs))
        if self.use_act_quant:
            x = self.act_quantizer(x)
        x_max = torch.max(x, dim=1, keepdim=True)[0]
        if self.last_vfe:
            return x_max
        return torch.cat([x, x_max.repeat(1, inputs.shape[1], 1)], dim=2)


class QuantPillarVFE(nn.Module):
    def __init__(self, pillar_vfe: PillarVFE, weight_quant_params={}, act_quant_params={}):
        super().__init__()
        self.with_distance = pillar_vfe.with_distance
        self.use_absolute_xyz = pillar_vfe.use_absolute_xyz
        self.voxel_x, self.voxel_y, self.voxel_z = pillar_vfe.voxel_x, pillar_vfe.voxel_y, pillar_vfe.voxel_z
        self.x_offset, self.y_offset, self.z_offset = pillar_vfe.x_offset, pillar_vfe.y_offset, pillar_vfe.z_offset
        self.pfn_layers = nn.ModuleList(QuantPFNLayer(l, weight_quant_params, act_quant_params)
                                        for l in pillar_vfe.pfn_layers)

    def forward(self, batch_dict):
        feats = augment_pillars(batch_dict["voxel_features"], batch_dict["voxel_num_points"],
                                batch_dict["voxel_coords"], (self.voxel_x, self.voxel_y, self.voxel_z),
                                (self.x_offset, self.y_offset, self.z_offset), self.use_absolute_xyz,
                                self.with_distance)
        for pfn in self.pfn_layers:
            feats = pfn(feats)
        batch_dict["pillar_features"] = feats.squeeze()
        return batch_dict


class QuantPointPillar(nn.Module):
    def __init__(self, point_pillar: PointPillar, weight_quant_params={}, act_quant_params={}):
        super().__init__()
        self.pillar_vfe = QuantPillarVFE(point_pillar.pillar_vfe, weight_quant_params, act_quant_params)
        self.scatter = point_pillar.scatter

    def forward(self, data_dict, modality_name):
        inp = data_dict[f"inputs_{modality_name}"]
        batch_dict = {"voxel_features": inp["voxel_features"], "voxel_coords": inp["voxel_coords"],
                      "voxel_num_points": inp["voxel_num_points"]}
        return self.scatter(self.pillar_vfe(batch_dict))["spatial_features"]

    def bev_delta(self) -> float:
        """Scale of the uint8 grid the BEV map lies on = the last PFN block's post-ReLU act quantizer."""
        q = self.pillar_vfe.pfn_layers[-1].act_quantizer
        d = q.delta
        return float(d.detach().reshape(-1)[0].item()) if isinstance(d, torch.Tensor) else float(d)


opencood_specials[PointPillar] = QuantPointPillar
