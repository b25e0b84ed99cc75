"""Pyramid-fusion building blocks on the libqv2x kernels.

* ``BottleneckEngine`` -- one calibrated ``QuantBottleneck`` (opencood/quant/quant_block.py:100-134) of the ResNeXt
  pyramid backbone (``ResNetModified(Bottleneck, groups=32, width_per_group=4)``, pyramid_fuse.py:69-77) as four int8
  tensor-core convs: 1x1 -> grouped 3x3 (dense block-diagonal GEMM) -> 1x1 with the shortcut added in its epilogue
  before the ReLU and the block's quantizer; the optional 1x1 strided downsample conv writes FP32 (it has no
  quantizer) and is that shortcut.  Activations stay uint8 NHWC; per-pixel code sums travel with them so no layer
  re-reads its input to correct for the weight zero-points.
* ``OccupancyHead`` -- ``single_head_i`` (1x1 conv to one channel, no quantizer; quant_block.py:474-478).
* ``weighted_fuse_level`` -- score-weighted fusion of one level from codes + occupancy logits.

What is not here yet: the first block of stage 0 reads the FP32 (off-grid) decoded features, and the stage / deblock
wiring of ``QuantPyramidFusion.forward_collab``; see DESIGN.md section 1.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine as E


def _layer(c, *, ksize, stride, pad, relu, in_delta, out_delta, groups=1):
    return E.QLayer(kind=0, w_int=c["w_int"], w_delta=c["w_delta"], w_zp=c["w_zp"], bias=c.get("bias"), ksize=ksize,
                    stride=stride, pad=pad, w_bits=int(c.get("w_bits", 8)), relu=relu, in_delta=in_delta,
                    out_delta=out_delta, groups=groups)


class BottleneckEngine:
    """params: ``conv1`` / ``conv2`` / ``conv3`` (/ ``down``) dicts with ``w_int`` (PyTorch layout, conv2 grouped:
    [width, width/groups, 3, 3]), ``w_delta``, ``w_zp``, ``bias``; ``conv1.act_delta``, ``conv2.act_delta``,
    ``out_delta`` (the block's act_quantizer); ``stride``; ``groups``.  ``in_delta``: scale of the input codes."""

    def __init__(self, params: dict, in_delta: float):
        p = params
        self.in_delta, self.out_delta = float(in_delta), float(p["out_delta"])
        self.stride = int(p["stride"])
        d1, d2 = float(p["conv1"]["act_delta"]), float(p["conv2"]["act_delta"])
        self.conv1 = _layer(p["conv1"], ksize=1, stride=1, pad=0, relu=True, in_delta=in_delta, out_delta=d1)
        self.conv2 = _layer(p["conv2"], ksize=3, stride=self.stride, pad=1, relu=True, in_delta=d1, out_delta=d2,
                            groups=int(p["groups"]))
        self.conv3 = _layer(p["conv3"], ksize=1, stride=1, pad=0, relu=True, in_delta=d2, out_delta=self.out_delta)
        self.down = None
        if "down" in p:
            # no ReLU, no quantizer: out_delta is unused on the FP32-output path
            self.down = _layer(p["down"], ksize=1, stride=self.stride, pad=0, relu=False, in_delta=in_delta,
                               out_delta=1.0)
        self.cin, self.cout = self.conv1.cin, self.conv3.cout

    def forward(self, x: torch.Tensor, rowsum: torch.Tensor | None = None, want_rowsum: bool = False,
                taps: dict | None = None):
        """x uint8 NHWC [n, H, W, cin] (scale in_delta) -> uint8 NHWC [n, H/stride, W/stride, cout] (scale out_delta).
        rowsum: per-pixel sums of x (int32 [n, H, W]) when the producer emitted them.  With want_rowsum the sums of
        the output are returned too.  taps (test hook) receives the intermediate tensors."""
        n, h, w, _ = x.shape
        dev = x.device
        if rowsum is None:
            rowsum = E.rowsum_u8(x, 0, self.cin)
        rs1 = torch.zeros((n, h, w), dtype=torch.int32, device=dev)
        q1 = self.conv1.forward(x, rowsum_in=[rowsum], rowsum_out=rs1)
        ho, wo = self.conv2.out_shape(h, w)
        rs2 = torch.zeros((n, ho, wo), dtype=torch.int32, device=dev)
        q2 = self.conv2.forward(q1, rowsum_in=[rs1], rowsum_out=rs2)
        rs_out = torch.zeros((n, ho, wo), dtype=torch.int32, device=dev) if want_rowsum else None
        if self.down is not None:
            res = torch.empty((n, ho, wo, self.cout), dtype=torch.float32, device=dev)
            self.down.forward(x, rowsum_in=[rowsum], out_f32=res)
            out = self.conv3.forward(q2, rowsum_in=[rs2], residual=res, rowsum_out=rs_out)
        else:
            res = None
            out = self.conv3.forward(q2, rowsum_in=[rs2], residual=x, res_delta=self.in_delta, rowsum_out=rs_out)
        if taps is not None:
            taps.update(q1=q1, q2=q2, res=res)
        return (out, rs_out) if want_rowsum else out


class OccupancyHead:
    """``single_head_i``: nn.Conv2d(C, 1, 1) wrapped in a QuantModule whose output feeds sigmoid directly
    (quant_block.py:474-478, 516-520).  The one real output channel is padded to the 64-column tile of the GEMM
    (columns 1..63 carry the weight zero-point, i.e. zero weights)."""

    def __init__(self, w_int, w_delta, w_zp, bias, in_delta, w_bits=8):
        w_int = np.asarray(w_int, np.uint8)
        c = w_int.shape[1]
        wp = np.full((64, c, 1, 1), int(np.asarray(w_zp).reshape(-1)[0]), np.uint8)
        wp[0] = w_int[0]
        dl = np.full(64, np.float32(np.asarray(w_delta).reshape(-1)[0]), np.float32)
        zp = np.full(64, np.float32(np.asarray(w_zp).reshape(-1)[0]), np.float32)
        b = np.zeros(64, np.float32)
        if bias is not None:
            b[0] = np.asarray(bias, np.float32).reshape(-1)[0]
        self.layer = E.QLayer(kind=0, w_int=wp, w_delta=dl, w_zp=zp, bias=b, ksize=1, stride=1, pad=0, w_bits=w_bits,
                              relu=False, in_delta=in_delta, out_delta=1.0)

    def forward(self, x: torch.Tensor, rowsum: torch.Tensor | None = None) -> torch.Tensor:
        """x uint8 NHWC [n, H, W, C] -> occupancy logits float32 [n, H, W]."""
        n, h, w, _ = x.shape
        buf = torch.empty((n, h, w, 64), dtype=torch.float32, device=x.device)
        self.layer.forward(x, rowsum_in=None if rowsum is None else [rowsum], out_f32=buf)
        return buf[..., 0].contiguous()


def weighted_fuse_level(codes: torch.Tensor, delta: float, occ: torch.Tensor, affine) -> torch.Tensor:
    """One level of QuantPyramidFusion.forward_collab (quant_block.py:516-539): codes uint8 NHWC [N, H, W, C] of the
    level's features (scale delta, agent 0 = ego), occ float32 [N, H, W] logits of single_head_i, affine [N, 2, 3]
    -> fused float32 [H, W, C]."""
    feat = E.dequantize_u8(codes, delta)
    return E.fuse_weighted(feat, occ, affine, score_is_logit=True)
