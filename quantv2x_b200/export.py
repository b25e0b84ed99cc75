"""Export of a calibrated quantized model to libqv2x engines.

The reference never persists PTQ results: every run re-calibrates and re-fake-quantizes
the weights on every forward (quant_layer.py:393).  Here the integer parameters are extracted ONCE from the
calibrated ``QuantModule``s -- integer weight grid, per-channel weight delta / zero-point, folded bias,
activation deltas -- and handed to the C ABI, which packs them for the tcgen05 kernels.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import engine as E
from .quant.quant_block import QuantBaseBEVBackbone, QuantDownsampleConv
from .quant.quant_layer import QuantModule, StraightThrough


def _scalar(v) -> float:
    return float(v.detach().reshape(-1)[0].item()) if isinstance(v, torch.Tensor) else float(v)


def qlayer_from_module(qm: QuantModule, in_delta, extra_pad: int = 0) -> E.QLayer:
    """Build the GPU layer for one calibrated QuantModule (conv or transposed conv).

    in_delta: activation scale(s) of the tensor(s) feeding the module (1 value, or 3 for the concat input).
    extra_pad: zero padding applied by an explicit nn.ZeroPad2d in front of the conv (the first conv of every
    backbone stage: ZeroPad2d(1) + padding 0, reference base_bev_backbone.py:41-46)."""
    if not isinstance(qm.norm_function, StraightThrough):
        raise ValueError("BatchNorm must be folded before export (QuantModel(..., is_fusing=True))")
    if not (qm.weight_quantizer.inited and qm.act_quantizer.inited):
        raise ValueError("quantizers are not calibrated")
    if qm.disable_act_quant:
        raise ValueError("layer has no output quantizer; it does not belong to the integer path")
    act = qm.activation_function
    relu = isinstance(act, nn.ReLU)
    if not relu and not isinstance(act, (StraightThrough, nn.Identity)):
        # QuantModel also fuses nn.ReLU6 (quant_model.py:50); the integer epilogue implements ReLU or nothing
        raise NotImplementedError(f"activation {type(act).__name__} has no integer epilogue (ReLU or identity only)")
    out_zp = _scalar(qm.act_quantizer.zero_point)
    if out_zp != 0.0:
        raise ValueError(f"activation zero-point {out_zp} != 0: the integer path needs post-ReLU activations")
    w_int, w_delta, w_zp = qm.integer_weight()
    bias = None if qm.bias is None else qm.bias.detach().cpu().numpy()
    kw = qm.fwd_kwargs
    if qm.fwd_func is torch.nn.functional.conv2d:
        k = w_int.shape[2]
        assert w_int.shape[2] == w_int.shape[3] and kw["groups"] == 1 and tuple(kw["dilation"]) == (1, 1)
        stride, pad = kw["stride"][0], kw["padding"][0] + extra_pad
        kind = 0
    elif qm.fwd_func is torch.nn.functional.conv_transpose2d:
        k = w_int.shape[2]
        stride, pad, kind = kw["stride"][0], 0, 1
        assert k == stride and tuple(kw["padding"]) == (0, 0) and extra_pad == 0
    else:
        raise ValueError("only Conv2d / ConvTranspose2d modules run on the integer path")
    return E.QLayer(kind=kind, w_int=w_int, w_delta=w_delta, w_zp=w_zp, bias=bias, ksize=k, stride=stride, pad=pad,
                    w_bits=qm.weight_quantizer.n_bits, relu=relu, in_delta=in_delta,
                    out_delta=_scalar(qm.act_quantizer.delta), out_zp=out_zp, out_bits=qm.act_quantizer.n_bits)


class BlockEngine:
    """A Plan plus the scales needed to enter / leave it from FP32 NCHW tensors (module-boundary drop-in)."""

    def __init__(self, plan: E.Plan, in_deltas, in_group_channels, out_deltas, out_group_channels):
        self.plan = plan
        self.in_deltas, self.in_group_channels = list(in_deltas), list(in_group_channels)
        self.out_deltas, self.out_group_channels = list(out_deltas), list(out_group_channels)

    def forward_u8(self, x_u8, out=None, slot=0, rowsum_in=None):
        return self.plan.forward(x_u8, out=out, slot=slot, rowsum_in=rowsum_in)

    def forward_nchw(self, x: torch.Tensor) -> torch.Tensor:
        """FP32 NCHW (values on the producers' grids) -> FP32 NCHW de-quantized output of the block."""
        if not x.is_cuda:
            raise RuntimeError("quantized inference runs on the GPU library only (no CPU fallback)")
        n, c, h, w = x.shape
        xq = torch.empty((n, h, w, c), dtype=torch.uint8, device=x.device)
        base = 0
        for d, cg in zip(self.in_deltas, self.in_group_channels):
            E.quantize_nchw_to_nhwc_u8(x[:, base:base + cg].contiguous().float(), d, out=xq, out_cbase=base)
            base += cg
        y = self.plan.forward(xq)
        outs, base = [], 0
        for d, cg in zip(self.out_deltas, self.out_group_channels):
            outs.append(E.dequant_nhwc_u8_to_nchw_f32(y[..., base:base + cg].contiguous(), d))
            base += cg
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)


def _backbone_steps(qb: QuantBaseBEVBackbone, in_delta: float, first_buf: int, in_buf: int):
    """Steps of blocks + deblocks.  Returns (steps, buf_channels dict, concat_buf, deltas of the concat groups, next_buf)."""
    steps, chans = [], {}
    nb = first_buf
    deltas, widths = [], []
    x_buf, x_delta = in_buf, in_delta
    cat_buf = None
    ups = []
    for i, blk in enumerate(qb.blocks):
        mods = list(blk)
        assert isinstance(mods[0], nn.ZeroPad2d), "backbone stages start with ZeroPad2d"
        pad = mods[0].padding
        assert len(set(pad)) == 1, "ZeroPad2d must pad all sides equally"
        ping = [nb, nb + 1]
        nb += 2
        for j, qm in enumerate(mods[1:]):
            layer = qlayer_from_module(qm, [x_delta], extra_pad=pad[0] if j == 0 else 0)
            ob = ping[j % 2]
            chans[ob] = layer.cout
            steps.append((layer, x_buf, 0, ob, 0))
            x_buf, x_delta = ob, _scalar(qm.act_quantizer.delta)
        ups.append((x_buf, x_delta))
    if len(qb.deblocks) > 0:
        assert len(qb.deblocks) == qb.num_levels, "an extra deblock after the concat is not supported"
        cat_buf = nb
        nb += 1
        cbase = 0
        for i, de in enumerate(qb.deblocks):
            qm = de[0]
            src_buf, src_delta = ups[i]
            layer = qlayer_from_module(qm, [src_delta])
            steps.append((layer, src_buf, 0, cat_buf, cbase))
            deltas.append(_scalar(qm.act_quantizer.delta))
            widths.append(layer.cout)
            cbase += layer.cout
        chans[cat_buf] = cbase
    else:
        raise ValueError("backbones without deblocks are not supported")
    return steps, chans, cat_buf, deltas, widths, nb


def _shrinker_steps(qs: QuantDownsampleConv, in_deltas, in_buf: int, first_buf: int):
    steps, chans = [], {}
    nb = first_buf
    x_buf, x_deltas = in_buf, list(in_deltas)
    for dc in qs.layers:
        for qm in dc.double_conv:
            layer = qlayer_from_module(qm, x_deltas)
            chans[nb] = layer.cout
            steps.append((layer, x_buf, 0, nb, 0))
            x_buf, x_deltas = nb, [_scalar(qm.act_quantizer.delta)]
            nb += 1
    return steps, chans, x_buf, x_deltas[0], nb


def _make_plan(steps, chans, in_channels, out_buf):
    """Renumber buffers so that the output is the last id, as qv2x_plan expects."""
    ids = sorted(chans)
    order = [b for b in ids if b != out_buf] + [out_buf]
    remap = {0: 0}
    for new, old in enumerate(order, start=1):
        remap[old] = new
    buf_channels = [in_channels] + [chans[old] for old in order]
    return E.Plan([(l, remap[ib], ic, remap[ob], oc) for (l, ib, ic, ob, oc) in steps], buf_channels)


def build_modality_engines(backbone: QuantBaseBEVBackbone, shrinker: QuantDownsampleConv, bev_delta: float):
    """Returns dict(fused=BlockEngine over backbone+shrinker, backbone=..., shrinker=...)."""
    in_channels = backbone.blocks[0][1].weight.shape[1]
    b_steps, b_ch, cat_buf, cat_deltas, cat_widths, nb = _backbone_steps(backbone, bev_delta, 1, 0)
    s_steps, s_ch, out_buf, out_delta, nb = _shrinker_steps(shrinker, cat_deltas, cat_buf, nb)
    fused = _make_plan(b_steps + s_steps, {**b_ch, **s_ch}, in_channels, out_buf)
    engines = {"fused": BlockEngine(fused, [bev_delta], [in_channels], [out_delta], [s_ch[out_buf]])}
    engines["backbone"] = BlockEngine(_make_plan(b_steps, b_ch, in_channels, cat_buf), [bev_delta], [in_channels],
                                      cat_deltas, cat_widths)
    # the shrinker alone: its input buffer 0 is the concat tensor
    s_alone = [(l, 0 if ib == cat_buf else ib, ic, ob, oc) for (l, ib, ic, ob, oc) in s_steps]
    engines["shrinker"] = BlockEngine(_make_plan(s_alone, s_ch, sum(cat_widths), out_buf), cat_deltas, cat_widths,
                                      [out_delta], [s_ch[out_buf]])
    engines["out_delta"] = out_delta
    return engines


# ---------------------------------------------------------------------------------------------------
# model level
# ---------------------------------------------------------------------------------------------------
def pillar_spec(enc) -> dict:
    """Plain-numpy description of a calibrated QuantPointPillar (reference quant_block.py:589-741): the
    fake-quantized Linear weights (BN folded), both activation quantizers and the grid geometry."""
    vfe = enc.pillar_vfe
    if len(vfe.pfn_layers) != 1 or not vfe.pfn_layers[0].last_vfe or not vfe.use_absolute_xyz or vfe.with_distance:
        raise NotImplementedError("the pillar kernel covers the single-layer PFN with absolute xyz and no distance "
                                  "feature (the configuration of every yaml of the hot path)")
    pfn = vfe.pfn_layers[0]
    lin = pfn.linear
    with torch.no_grad():
        w_hat = lin.weight_quantizer(lin.weight) if lin.use_weight_quant else lin.org_weight
    pre = None
    if lin.use_act_quant and not lin.disable_act_quant:
        aq = lin.act_quantizer
        pre = (_scalar(aq.delta), _scalar(aq.zero_point), int(aq.n_bits))
    oq = pfn.act_quantizer
    out_q = (_scalar(oq.delta), _scalar(oq.zero_point), int(oq.n_bits))
    if not pfn.use_act_quant:
        raise NotImplementedError("the BEV map must lie on the PFN block's activation grid (use_act_quant)")
    return dict(w_hat=w_hat.detach().cpu().numpy().astype(np.float32).copy(),
                bias=None if lin.bias is None else lin.bias.detach().cpu().numpy().astype(np.float32).copy(),
                nx=int(enc.scatter.nx), ny=int(enc.scatter.ny),
                voxel_size=(float(vfe.voxel_x), float(vfe.voxel_y), float(vfe.voxel_z)),
                offset=(float(vfe.x_offset), float(vfe.y_offset), float(vfe.z_offset)), pre_quant=pre, out_quant=out_q)


def build_pillar_engine(enc):
    """libqv2x engine of a calibrated QuantPointPillar."""
    from .engine import PillarEngine

    sp = pillar_spec(enc)
    return PillarEngine(sp["w_hat"], sp["bias"], nx=sp["nx"], ny=sp["ny"], voxel_size=sp["voxel_size"],
                        offset=sp["offset"], pre_quant=sp["pre_quant"], out_quant=sp["out_quant"])


def attach_engines(qmodel, bev_delta: float | None = None, device=None):
    """Build every libqv2x engine of a calibrated ``QuantModel(HeterBaselineCollabCodebookMC)`` and attach them:
    block-level engines to the quantized backbone / shrinker wrappers (module-boundary drop-in) and a
    frame-level ``CollabPipeline`` to the model (fast path used by encode_features / decode_features).

    bev_delta: scale of the uint8 BEV grid; read from the quantized PointPillar encoder when omitted."""
    from .pipeline import CollabPipeline, heads_from_quant_modules
    from .quant.quant_block import QuantPointPillar

    model = qmodel.model
    device = device or torch.device("cuda", torch.cuda.current_device())
    if getattr(model, "shrink_flag", False):
        # reference heter_baseline_collab_codebook_mc.py:156-157 applies shrink_conv between fusion and the heads;
        # the ego stage here goes fuse -> heads, so such a config would silently give other predictions
        raise NotImplementedError("a top-level `shrink_header` (post-fusion shrink_conv) is not on the B200 path; "
                                  "the shipped V2X-Real configs do not use it")
    if model.fusion_method == "att":
        fd = model.args["att"]["feat_dim"]
        if int(fd) != int(model.channel):
            raise NotImplementedError(f"AttFusion feat_dim {fd} != feature channels {model.channel}: the fused kernel "
                                      "scales scores by sqrt(C) (reference fusion_in_one.py:14-45 uses sqrt(feat_dim))")
    for name in model.modality_name_list:
        enc = getattr(model, f"encoder_{name}")
        bb, sh = getattr(model, f"backbone_{name}"), getattr(model, f"shrinker_{name}")
        d_in = bev_delta
        if d_in is None:
            if not isinstance(enc, QuantPointPillar):
                raise ValueError("bev_delta is required when the encoder is not quantized")
            d_in = enc.bev_delta()
        engines = build_modality_engines(bb, sh, d_in)
        bb.attach_engine(engines["backbone"])
        sh.attach_engine(engines["shrinker"])
        rng = model.cav_range
        vs = model.args[name]["encoder_args"]["voxel_size"]
        H, W = int(round((rng[4] - rng[1]) / vs[1])), int(round((rng[3] - rng[0]) / vs[0]))
        pipe = CollabPipeline(engines["fused"], engines["out_delta"], model.codebook.engine(),
                              heads_from_quant_modules(model.cls_head, model.reg_head, model.dir_head),
                              model.fusion_method, (H, W), device)
        pipe.bev_delta = d_in
        model.codebook.set_input_scale(engines["out_delta"])   # reference call shape: codebook.encode(flattened)
        # pillar-level input: the PFN + scatter kernel, when the encoder is quantized, calibrated, and its output
        # grid is the one the backbone engine was built for
        if isinstance(enc, QuantPointPillar):
            oq = enc.pillar_vfe.pfn_layers[-1].act_quantizer
            if getattr(oq, "inited", False) and abs(enc.bev_delta() - d_in) <= 1e-6 * abs(d_in):
                pipe.pillar_engine = build_pillar_engine(enc)
        model._pipelines[name] = pipe
    return model


def export_spec(qmodel, bev_delta: float, modality: str = "m1") -> dict:
    """Plain-numpy description of the calibrated model (float weights + every quantizer's delta / zero-point).
    This is what a serialized PTQ engine would hold, and what the CPU oracle (oracle/frame_ref.py) consumes."""
    model = qmodel.model

    def layer(qm, extra_pad=0):
        _, w_delta, w_zp = qm.integer_weight()
        kw = qm.fwd_kwargs
        kind = 0 if qm.fwd_func is torch.nn.functional.conv2d else 1
        aq = qm.act_quantizer
        return dict(kind=kind, w=qm.weight.detach().cpu().numpy().copy(),
                    bias=None if qm.bias is None else qm.bias.detach().cpu().numpy().copy(),
                    w_bits=qm.weight_quantizer.n_bits, w_delta=w_delta, w_zp=w_zp, stride=kw["stride"][0],
                    pad=(kw["padding"][0] + extra_pad) if kind == 0 else 0,
                    relu=isinstance(qm.activation_function, nn.ReLU),
                    act_delta=None if qm.disable_act_quant else _scalar(aq.delta),
                    act_zp=_scalar(aq.zero_point), act_bits=aq.n_bits)

    bb, sh = getattr(model, f"backbone_{modality}"), getattr(model, f"shrinker_{modality}")
    spec = {"bev_delta": float(bev_delta), "fusion": model.fusion_method}
    spec["blocks"] = [[layer(qm, extra_pad=blk[0].padding[0] if j == 0 else 0) for j, qm in enumerate(list(blk)[1:])]
                      for blk in bb.blocks]
    spec["deblocks"] = [layer(de[0]) for de in bb.deblocks]
    spec["shrinker"] = [layer(qm) for dc in sh.layers for qm in dc.double_conv]
    cbs, heads = model.codebook.head_params()
    spec["codebook"] = dict(codebooks=cbs, heads=heads)
    ws, bs = [], []
    for qm in (model.cls_head, model.reg_head, model.dir_head):
        with torch.no_grad():
            w = qm.weight_quantizer(qm.weight) if qm.use_weight_quant else qm.org_weight
        ws.append(w.detach().reshape(w.shape[0], -1).cpu().numpy())
        bs.append(qm.bias.detach().cpu().numpy())
    spec["heads"] = dict(w=np.concatenate(ws, 0), b=np.concatenate(bs, 0))
    return spec
