"""ORACLE (test infrastructure, not product code): the collaborative forward of the quantized pyramid backbone,
QuantPyramidFusion.forward_collab (opencood/quant/quant_block.py:504-541), up to the fused per-level features:

  FP32 decoded features -> ResNeXt stages of QuantBottleneck blocks (quant_block.py:100-134; the very first conv
  reads the FP32 features, QuantModule quantizes outputs only: quant_layer.py:391-410) -> per level: single_head_i
  occupancy logits (quant_block.py:516-518) -> weighted_fuse (pyramid_fuse.py:17-62) -> decode_multiscale_feature
  (quant_block.py:441-458): one ConvTranspose2d (kernel = stride) + ReLU + quantizer per level on the FP32 fused map,
  concatenated along the channels.

Integer layers follow oracle/int_oracle.py; the FP32-input 1x1 conv follows the order of the product's FP32 GEMM
(bias first, then one fma per input channel in ascending order).  Pinned against the reference by
tests/golden/pyramid_backbone.npz (oracle/gen_golden_pyramid.py).
"""
from __future__ import annotations

import numpy as np

from . import fusion_oracle, int_oracle
from .int_oracle import f32, fma32


def dequant_weight(c):
    """(w_int - zp) * delta in FP32, as the reference's weight quantizer returns it; [cout, cin] for a 1x1 conv."""
    w = (c["w_int"].astype(f32) - np.asarray(c["w_zp"], f32).reshape(-1, 1, 1, 1)) * np.asarray(c["w_delta"], f32).reshape(-1, 1, 1, 1)
    return w.reshape(w.shape[0], -1)


def conv1x1_f32(x, w_hat, bias):
    """x float32 [..., cin] -> float32 [..., cout]: y = bias, then y = fma(x[k], w[o][k], y) for k ascending."""
    y = np.broadcast_to(np.asarray(bias, f32), x.shape[:-1] + (w_hat.shape[0],)).astype(f32)
    for k in range(x.shape[-1]):
        y = fma32(x[..., k:k + 1], w_hat[:, k], y)
    return y


def quantize(y, delta, bits=8):
    return np.clip(np.rint(np.asarray(y, f32) / f32(delta)), 0, 2 ** bits - 1).astype(np.uint8)


def first_block(x, p, q1_override=None):
    """The block that reads the FP32 features (identity shortcut: stage 0 keeps 64 channels at stride 1)."""
    assert "down" not in p and p["stride"] == 1
    c1, c2, c3 = p["conv1"], p["conv2"], p["conv3"]
    q1 = quantize(conv1x1_f32(x, dequant_weight(c1), c1["bias"]), c1["act_delta"])
    q1_used = q1 if q1_override is None else q1_override
    _, q2 = int_oracle.conv_oracle(q1_used, c2["w_int"], c2["w_delta"], c2["w_zp"], c2["bias"], c1["act_delta"],
                                   c2["act_delta"], stride=1, pad=1, groups=p["groups"])
    _, out = int_oracle.conv_oracle(q2, c3["w_int"], c3["w_delta"], c3["w_zp"], c3["bias"], c2["act_delta"],
                                    p["out_delta"], stride=1, pad=0, residual=x)
    return dict(q1=q1, q2=q2, out=out)


def backbone_collab(x, P, aff, layer_nums, q1_override=None):
    """x float32 [N, H, W, 64] (agent 0 = ego), P: 'l{i}.b{j}' block dicts + 'head{i}' conv dicts, aff [N, 2, 3].
    Returns per level: codes uint8 [N, h, w, C], delta, occ float32 [N, h, w], fused float32 [h, w, C]; and q1 of
    the first block (the only FP32-accumulated codes)."""
    levels = []
    cur, cur_delta, q1_first = None, None, None
    for li, nb in enumerate(layer_nums):
        for bi in range(nb):
            p = P[f"l{li}.b{bi}"]
            if cur is None:
                r = first_block(x, p, q1_override)
                q1_first = r["q1"]
            else:
                r = int_oracle.bottleneck_oracle(cur, cur_delta, p)
            cur, cur_delta = r["out"], f32(p["out_delta"])
        h = P[f"head{li}"]
        _, occ = int_oracle.conv_oracle(cur, h["w_int"], h["w_delta"], h["w_zp"], h["bias"], cur_delta, None,
                                        stride=1, pad=0, relu=False)
        if "act_delta" in h:     # the head's own output quantizer (active in the reference, quant_block.py:474-478)
            d, zp, qmax = f32(h["act_delta"]), f32(h["act_zp"]), f32(2 ** int(h["act_bits"]) - 1)
            occ = ((np.clip(np.rint(occ / d) + zp, f32(0), qmax) - zp) * d).astype(f32)
        fused = fusion_oracle.weighted_fusion(cur.astype(f32) * cur_delta, occ[..., 0], aff)
        levels.append(dict(codes=cur, delta=cur_delta, occ=occ[..., 0], fused=fused))
    return levels, q1_first


def deblock_f32(fused, up):
    """QuantModule(ConvTranspose2d(cin, cout, s, stride=s)) + ReLU + act quantizer on an FP32 map (quant_block.py:
    389-396).  fused float32 [h, w, cin]; up: w_int [cin, cout, s, s], w_delta / w_zp per cin (dim 0 of a transposed
    conv weight, quant_layer.py:325-335), bias [cout], act_delta, stride.  Returns codes uint8 [h*s, w*s, cout]:
    y = bias, then one fma per input channel in ascending order (the FP32 GEMM's order), ReLU + quantize."""
    s = int(up["stride"])
    w_hat = (up["w_int"].astype(f32) - np.asarray(up["w_zp"], f32).reshape(-1, 1, 1, 1)) * np.asarray(up["w_delta"], f32).reshape(-1, 1, 1, 1)
    cin, cout = w_hat.shape[:2]
    h, w, _ = fused.shape
    rows = w_hat.transpose(2, 3, 1, 0).reshape(s * s * cout, cin)          # row (dy*s + dx)*cout + co
    y = conv1x1_f32(np.asarray(fused, f32), rows, np.tile(np.asarray(up["bias"], f32), s * s))
    q = quantize(y, up["act_delta"])                                        # [h, w, s*s*cout]
    return q.reshape(h, w, s, s, cout).transpose(0, 2, 1, 3, 4).reshape(h * s, w * s, cout)


def decode_multiscale(fused_levels, P):
    """[fused_0, fused_1, ...] -> codes uint8 [H, W, sum cout] and the per-level scales."""
    parts = [deblock_f32(f, P[f"up{li}"]) for li, f in enumerate(fused_levels)]
    return np.concatenate(parts, axis=-1), [f32(P[f"up{li}"]["act_delta"]) for li in range(len(parts))]
