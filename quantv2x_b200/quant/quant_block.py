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

    def _fingerprint(self):
        """Identity + in-place version of every tensor an exported engine was built from (weights, biases, quantizer
        scales / zero-points, bit widths): re-calibration, bitwidth_refactor, load_state_dict or an optimizer step
        change it.  No device synchronisation."""
        fp = []
        for m in self.quant_modules():
            for t in (m.weight, m.bias, m.weight_quantizer.delta, m.weight_quantizer.zero_point,
                      m.act_quantizer.delta, m.act_quantizer.zero_point):
                fp.append((id(t), t._version) if isinstance(t, torch.Tensor) else repr(t))
            fp.append((m.weight_quantizer.n_bits, m.act_quantizer.n_bits, type(m.weight_quantizer).__name__))
        return tuple(fp)

    def attach_engine(self, engine):
        self._engine = engine
        self._engine_fp = None if engine is None else self._fingerprint()

    def _check_engine(self):
        if self._engine is None:
            raise RuntimeError(
                f"{type(self).__name__}: quantized inference requested but no libqv2x engine is attached; call "
                "quantv2x_b200.export.attach_engines(model) after calibration (there is no CPU fallback)")
        if getattr(self, "_engine_fp", None) is not None and self._engine_fp != self._fingerprint():
            self._engine = None
            raise RuntimeError(
                f"{type(self).__name__}: parameters or quantizer state changed after the libqv2x engine was built "
                "(re-calibration / bitwidth_refactor / load_state_dict); the stale engine was dropped -- call "
                "attach_engines(model) again")

    def _run_engine(self, x):
        self._check_engine()
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

    def forward_float(self, x):
        for layer in self.layers:
            x = layer(x)
        return x

    def forward(self, x):
        if self.engine_ready():
            return self._run_engine(x)
        return self.forward_float(x)


# --------------------------------------------------------------------------------------------------
# Pyramid-fusion backbone: reference quant_block.py:100-134 (QuantBottleneck), :337-385
# (QuantResNetModified), :462-549 (QuantPyramidFusion).  The torch bodies below are the calibration path; after
# calibration ``export_params`` hands the integer parameters to quantv2x_b200.pyramid.PyramidBackboneEngine and
# ``attach_engine`` makes ``forward_collab`` run there.
# --------------------------------------------------------------------------------------------------
from ..pyramid_modules import (BasicBlock, BasicBlockStages, Bottleneck, PyramidFusion, ResNetBEVBackbone,  # noqa: E402
                               ResNeXtStages, weighted_fuse_torch)


def _conv_params(qm: QuantModule):
    w_int, w_delta, w_zp = qm.integer_weight()
    bias = None if qm.bias is None else qm.bias.detach().float().cpu().numpy()
    return dict(w_int=w_int, w_delta=w_delta, w_zp=w_zp, bias=bias, w_bits=int(qm.weight_quantizer.n_bits))


def _act_delta(q) -> float:
    if float(q.zero_point) != 0.0:
        raise ValueError("the integer path needs activation zero-points of 0 (post-ReLU quantizers)")
    return float(q.delta)


class QuantBasicBlock(BaseQuantBlock):
    """Mirror of the reference QuantBasicBlock (quant_block.py:68-97): conv1 with ReLU + quantizer, conv2 (and the
    downsample conv) without; shortcut, ReLU and the block's quantizer follow."""

    def __init__(self, basic_block: BasicBlock, weight_quant_params={}, act_quant_params={}):
        super().__init__()
        from .quant_layer import UniformAffineQuantizer

        self.conv1 = QuantModule(basic_block.conv1, weight_quant_params, act_quant_params)
        self.conv1.norm_function, self.conv1.activation_function = basic_block.bn1, basic_block.relu
        self.conv2 = QuantModule(basic_block.conv2, weight_quant_params, act_quant_params, disable_act_quant=True)
        self.conv2.norm_function = basic_block.bn2
        self.downsample = None
        if basic_block.downsample is not None:
            self.downsample = QuantModule(basic_block.downsample[0], weight_quant_params, act_quant_params,
                                          disable_act_quant=True)
            self.downsample.norm_function = basic_block.downsample[1]
        self.activation_function = basic_block.relu
        self.act_quantizer = UniformAffineQuantizer(**act_quant_params)
        self.stride = basic_block.stride

    def forward(self, x):
        residual = x if self.downsample is None else self.downsample(x)
        out = self.activation_function(self.conv2(self.conv1(x)) + residual)
        if self.use_act_quant:
            out = self.act_quantizer(out)
        return out

    def export_params(self) -> dict:
        """The dict quantv2x_b200.pyramid.BasicBlockEngine takes."""
        p = dict(stride=int(self.stride), out_delta=_act_delta(self.act_quantizer),
                 conv1=_conv_params(self.conv1), conv2=_conv_params(self.conv2))
        p["conv1"]["act_delta"] = _act_delta(self.conv1.act_quantizer)
        if self.downsample is not None:
            p["down"] = _conv_params(self.downsample)
        return p


class QuantBottleneck(BaseQuantBlock):
    """conv1 / conv2 with their own ReLU + quantizer, conv3 (and the downsample conv) without: the shortcut is added
    first, then the block's ReLU and quantizer."""

    def __init__(self, bottleneck: Bottleneck, weight_quant_params={}, act_quant_params={}):
        super().__init__()
        from .quant_layer import UniformAffineQuantizer

        self.conv1 = QuantModule(bottleneck.conv1, weight_quant_params, act_quant_params)
        self.conv1.norm_function, self.conv1.activation_function = bottleneck.bn1, bottleneck.relu
        self.conv2 = QuantModule(bottleneck.conv2, weight_quant_params, act_quant_params)
        self.conv2.norm_function, self.conv2.activation_function = bottleneck.bn2, bottleneck.relu
        self.conv3 = QuantModule(bottleneck.conv3, weight_quant_params, act_quant_params, disable_act_quant=True)
        self.conv3.norm_function = bottleneck.bn3
        self.downsample = None
        if bottleneck.downsample is not None:
            self.downsample = QuantModule(bottleneck.downsample[0], weight_quant_params, act_quant_params,
                                          disable_act_quant=True)
            self.downsample.norm_function = bottleneck.downsample[1]
        self.activation_function = bottleneck.relu
        self.act_quantizer = UniformAffineQuantizer(**act_quant_params)
        self.stride = bottleneck.stride
        self.groups = bottleneck.conv2.groups

    def forward(self, x):
        residual = x if self.downsample is None else self.downsample(x)
        out = self.conv3(self.conv2(self.conv1(x)))
        out = self.activation_function(out + residual)
        if self.use_act_quant:
            out = self.act_quantizer(out)
        return out

    def export_params(self) -> dict:
        """The dict quantv2x_b200.pyramid.BottleneckEngine takes."""
        p = dict(stride=int(self.stride), groups=int(self.groups), out_delta=_act_delta(self.act_quantizer))
        for n in ("conv1", "conv2", "conv3"):
            p[n] = _conv_params(getattr(self, n))
        p["conv1"]["act_delta"] = _act_delta(self.conv1.act_quantizer)
        p["conv2"]["act_delta"] = _act_delta(self.conv2.act_quantizer)
        if self.downsample is not None:
            p["down"] = _conv_params(self.downsample)
        return p


class QuantResNeXtStages(BaseQuantBlock):
    def __init__(self, stages: ResNeXtStages, weight_quant_params={}, act_quant_params={}):
        super().__init__()
        self.layernum = stages.layernum
        for i in range(self.layernum):
            setattr(self, f"layer{i}", nn.Sequential(*[QuantBottleneck(b, weight_quant_params, act_quant_params)
                                                        for b in getattr(stages, f"layer{i}")]))

    def forward(self, x):
        feats = []
        for i in range(self.layernum):
            x = getattr(self, f"layer{i}")(x)
            feats.append(x)
        return feats


class QuantResNetBEVBackbone(BaseQuantBlock):
    """Mirror of the reference QuantResNetBEVBackbone (quant_block.py:398-460) for the agent-side backbone of the pyramid
    models: BasicBlock stages wrapped as QuantBasicBlocks.  Calibrate with the torch body; ``export_params()`` ->
    quantv2x_b200.pyramid.ResNetBackboneEngine; with an engine attached, quantized forwards run on libqv2x."""

    def __init__(self, backbone: ResNetBEVBackbone, weight_quant_params={}, act_quant_params={}):
        super().__init__()
        self.model_cfg = backbone.model_cfg
        self.num_levels = backbone.num_levels
        self.num_bev_features = backbone.num_bev_features
        self.resnet = nn.Module()
        self.resnet.layernum = backbone.resnet.layernum
        for i in range(backbone.resnet.layernum):
            setattr(self.resnet, f"layer{i}", nn.Sequential(*[QuantBasicBlock(b, weight_quant_params, act_quant_params)
                                                               for b in getattr(backbone.resnet, f"layer{i}")]))
        self.deblocks = nn.ModuleList()

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        super().set_quant_state(weight_quant, act_quant)
        for m in self.modules():
            if isinstance(m, BaseQuantBlock) and m is not self:
                m.use_weight_quant, m.use_act_quant = weight_quant, act_quant

    def blocks(self):
        return [b for i in range(self.resnet.layernum) for b in getattr(self.resnet, f"layer{i}")]

    def get_multiscale_feature(self, spatial_features):
        feats, x = [], spatial_features
        for i in range(self.resnet.layernum):
            x = getattr(self.resnet, f"layer{i}")(x)
            feats.append(x)
        return feats

    def forward_float(self, spatial_features):
        x = self.get_multiscale_feature(spatial_features)
        return torch.cat(x, dim=1) if len(x) > 1 else x[0]

    def engine_ready(self) -> bool:
        if not (self.use_weight_quant and self.use_act_quant):
            return False
        return all(m.weight_quantizer.inited and (m.disable_act_quant or m.act_quantizer.inited)
                   for m in self.quant_modules()) and all(b.act_quantizer.inited for b in self.blocks())

    def forward(self, spatial_features):
        if self.engine_ready() and spatial_features.is_cuda:
            return self._run_engine(spatial_features)
        return self.forward_float(spatial_features)

    def export_params(self) -> list:
        """One dict per block, in execution order (see quantv2x_b200.pyramid.BasicBlockEngine)."""
        if self.resnet.layernum != 1:
            raise NotImplementedError("multi-stage agent backbones concatenate several scales; the engine covers the "
                                      "single-stage configuration of the pyramid yamls")
        return [b.export_params() for b in self.blocks()]

    def out_delta(self) -> float:
        return _act_delta(self.blocks()[-1].act_quantizer)


class QuantPyramidFusion(BaseQuantBlock):
    def __init__(self, pyramid_fusion: PyramidFusion, weight_quant_params={}, act_quant_params={}):
        super().__init__()
        self.model_cfg = pyramid_fusion.model_cfg
        self.stage = pyramid_fusion.stage
        self.align_corners = pyramid_fusion.align_corners
        self.num_levels = pyramid_fusion.num_levels
        self.num_bev_features = pyramid_fusion.num_bev_features
        self.resnet = QuantResNeXtStages(pyramid_fusion.resnet, weight_quant_params, act_quant_params)
        self.deblocks = nn.ModuleList()
        for deblock in pyramid_fusion.deblocks:
            qm = QuantModule(deblock[0], weight_quant_params, act_quant_params)
            qm.norm_function, qm.activation_function = deblock[1], deblock[2]
            self.deblocks.append(nn.Sequential(qm))
        for i in range(self.num_levels):
            setattr(self, f"single_head_{i}", QuantModule(getattr(pyramid_fusion, f"single_head_{i}"),
                                                          weight_quant_params, act_quant_params))

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        super().set_quant_state(weight_quant, act_quant)
        for m in self.modules():                      # the nested blocks own quantizers too
            if isinstance(m, BaseQuantBlock) and m is not self:
                m.use_weight_quant, m.use_act_quant = weight_quant, act_quant

    def get_multiscale_feature(self, spatial_features):
        return self.resnet(spatial_features)

    def decode_multiscale_feature(self, x):
        ups = [self.deblocks[i](x[i]) if len(self.deblocks) > 0 else x[i] for i in range(self.num_levels)]
        return torch.cat(ups, dim=1) if len(ups) > 1 else ups[0]

    def forward_single(self, spatial_features):
        feats = self.get_multiscale_feature(spatial_features)
        occ = [getattr(self, f"single_head_{i}")(feats[i]) for i in range(self.num_levels)]
        return self.decode_multiscale_feature(feats), occ

    def forward_collab(self, spatial_features, record_len, affine_matrix, agent_modality_list=None, cam_crop_info=None):
        if cam_crop_info:
            raise NotImplementedError("the camera crop mask is not built (LiDAR agents only)")
        if self.engine_ready():
            return self._run_engine_collab(spatial_features, record_len, affine_matrix)
        return self.forward_collab_float(spatial_features, record_len, affine_matrix)

    def forward_collab_float(self, spatial_features, record_len, affine_matrix):
        """The reference's fake-quant torch body (calibration, and the float side of parity tests)."""
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

    # ------------------------------------------------------------------ engine hand-off
    def layer_nums(self):
        return [len(getattr(self.resnet, f"layer{i}")) for i in range(self.num_levels)]

    def export_params(self) -> dict:
        """The dict quantv2x_b200.pyramid.PyramidBackboneEngine takes (every quantizer must be calibrated)."""
        P = {}
        for li in range(self.num_levels):
            for bi, blk in enumerate(getattr(self.resnet, f"layer{li}")):
                P[f"l{li}.b{bi}"] = blk.export_params()
            head = getattr(self, f"single_head_{li}")
            P[f"head{li}"] = _conv_params(head)
            if not head.disable_act_quant:
                # the reference wraps single_head_i with an ACTIVE output quantizer (quant_block.py:474-478 builds
                # QuantModule(single_head_i, wq, aq)): the logits reach sigmoid / weighted_fuse fake-quantized
                aq = head.act_quantizer
                P[f"head{li}"].update(act_delta=float(aq.delta), act_zp=float(aq.zero_point), act_bits=int(aq.n_bits))
            if len(self.deblocks) > 0:
                qm = self.deblocks[li][0]
                up = _conv_params(qm)
                up.update(act_delta=_act_delta(qm.act_quantizer), stride=int(qm.fwd_kwargs["stride"][0]))
                P[f"up{li}"] = up
        return P

    def _run_engine_collab(self, x, record_len, affine_matrix):
        """forward_collab on libqv2x: NCHW FP32 in, the reference's NCHW FP32 tensors out (drop-in boundary)."""
        if self._engine is None:
            raise RuntimeError("QuantPyramidFusion: quantized inference requested but no libqv2x engine is attached; "
                               "call attach_engine(PyramidBackboneEngine(self.export_params(), self.layer_nums())) "
                               "after calibration (there is no CPU fallback)")
        from .. import engine as E

        eng, outs, occs, start = self._engine, [], None, 0
        for b, n in enumerate(int(v) for v in record_len):
            xb = E.nchw_to_nhwc_f32(x[start:start + n].contiguous().float())
            taps = {}
            fused = eng.forward_collab(xb, affine_matrix[b][0, :n].to(x.device, torch.float32).contiguous(), taps=taps)
            cat = eng.decode_multiscale_feature(fused)                               # uint8 [1, H, W, sum cout]
            parts, base = [], 0
            for d in eng.deblocks:
                parts.append(E.dequant_nhwc_u8_to_nchw_f32(cat[..., base:base + d.cout].contiguous(), d.delta))
                base += d.cout
            outs.append(torch.cat(parts, dim=1))
            occ_b = [taps[f"l{i}.occ"].unsqueeze(1) for i in range(self.num_levels)]
            occs = occ_b if occs is None else [torch.cat([a, c]) for a, c in zip(occs, occ_b)]
            start += n
        return torch.cat(outs), occs


opencood_specials = {
    BaseBEVBackbone: QuantBaseBEVBackbone,
    DownsampleConv: QuantDownsampleConv,
    PyramidFusion: QuantPyramidFusion,
    ResNetBEVBackbone: QuantResNetBEVBackbone,
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
        # Above `part` pillars the reference feeds the Linear in slices of `part` rows (quant_block.py:611-618).  It
        # matters for calibration: an un-initialised output quantizer of the Linear updates its scale on every call
        # (running min / max), so the calibrated delta depends on the slicing; the mirror slices the same way.
        m = inputs.shape[0]
        if m > self.part:
            pieces = [self.linear(inputs[lo:lo + self.part]) for lo in range(0, (m // self.part + 1) * self.part,
                                                                            self.part)]
            x = F.relu(torch.cat(pieces, dim=0))
        else:
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
