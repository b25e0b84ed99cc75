"""Pyramid-fusion building blocks on the libqv2x kernels.

* ``BottleneckEngine`` -- one calibrated ``QuantBottleneck`` (opencood/quant/quant_block.py:100-134) of the ResNeXt
  pyramid backbone (``ResNetModified(Bottleneck, groups=32, width_per_group=4)``, pyramid_fuse.py:69-77) as four int8
  tensor-core convs: 1x1 -> grouped 3x3 (dense block-diagonal GEMM) -> 1x1 with the shortcut added in its epilogue
  before the ReLU and the block's quantizer; the optional 1x1 strided downsample conv writes FP32 (it has no
  quantizer) and is that shortcut.  Activations stay uint8 NHWC; per-pixel code sums travel with them so no layer
  re-reads its input to correct for the weight zero-points.
* ``BasicBlockEngine`` -- ``QuantBasicBlock`` (two 3x3 convs; the agent-side ResNetBEVBackbone of the pyramid models).
* ``OccupancyHead`` -- ``single_head_i`` (1x1 conv to one channel; its output quantizer, active in the reference
  (quant_block.py:474-478), is applied to the FP32 logits).
* ``weighted_fuse_level`` -- score-weighted fusion of one level from codes + occupancy logits.

* ``FirstBottleneckEngine`` -- the block that reads the FP32 (off-grid) decoded features: its first 1x1 conv is an
  FP32 GEMM (QuantModule quantizes outputs only, quant_layer.py:391-410), the shortcut is the FP32 input itself.
* ``PyramidBackboneEngine`` -- ``QuantPyramidFusion.forward_collab`` (quant_block.py:504-541) up to the fused
  per-level features: ResNeXt stages over every agent's map, occupancy heads, per-level fusion.

* ``DeblockF32`` / ``PyramidBackboneEngine.decode_multiscale_feature`` -- the deblocks after the fusion
  (quant_block.py:441-458): transposed convs (kernel = stride) on the FP32 fused maps as FP32 GEMMs, quantized and
  concatenated into the uint8 [H, W, 384] input of the shrink conv (three scales, as the att/max path's concat).

* ``ResNetBackboneEngine`` -- the agent-side ``QuantResNetBEVBackbone`` (a chain of BasicBlockEngine).

The model driver that strings these together (pillars -> agent backbone -> codebook | decode -> pyramid -> shrink
conv -> heads) is quantv2x_b200.pyramid_model; see DESIGN.md section 3.8.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import engine as E


class _ZeroPool:
    """Per-pixel code sums are accumulated with atomic adds by the producing conv, so every such buffer starts from
    zero.  One memset kernel per buffer was 52 launches (0.12 ms) of a pyramid frame: inside begin() / end() the
    buffers are slices of ONE zeroed allocation whose size was learned on the first pass over the same shapes."""

    def __init__(self):
        self.sizes, self.key, self.buf, self.off, self.used = {}, None, None, 0, 0

    def begin(self, key, device):
        if self.key is not None:             # nested use (block engines driven inside a model pass): keep the outer
            return False
        self.key, self.off, self.used = key, 0, 0
        need = self.sizes.get(key, 0)
        self.buf = torch.zeros((need,), dtype=torch.int32, device=device) if need else None
        return True

    def take(self, shape, device):
        numel = int(np.prod(shape))
        pad = (numel + 63) // 64 * 64        # 256-byte aligned slices
        if self.key is not None:
            self.used += pad
            if self.buf is not None and self.off + pad <= self.buf.numel() and self.buf.device == device:
                t = self.buf[self.off:self.off + numel].view(shape)
                self.off += pad
                return t
        return torch.zeros(shape, dtype=torch.int32, device=device)

    def end(self):
        self.sizes[self.key] = max(self.sizes.get(self.key, 0), self.used)
        self.key, self.buf = None, None


_zeros = _ZeroPool()


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
        rs1 = _zeros.take((n, h, w), dev)
        q1 = self.conv1.forward(x, rowsum_in=[rowsum], rowsum_out=rs1)
        ho, wo = self.conv2.out_shape(h, w)
        rs2 = _zeros.take((n, ho, wo), dev)
        q2 = self.conv2.forward(q1, rowsum_in=[rs1], rowsum_out=rs2)
        rs_out = _zeros.take((n, ho, wo), dev) if want_rowsum else None
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


class BasicBlockEngine:
    """One calibrated ``QuantBasicBlock`` (quant_block.py:68-97; the agent-side ResNetBEVBackbone of the pyramid
    models) as two int8 3x3 convs, the second with the shortcut in its epilogue; the optional strided 1x1 downsample
    conv writes the FP32 shortcut.  params: ``conv1`` / ``conv2`` (/ ``down``), ``conv1.act_delta``, ``out_delta``,
    ``stride``."""

    def __init__(self, params: dict, in_delta: float):
        p = params
        self.in_delta, self.out_delta = float(in_delta), float(p["out_delta"])
        self.stride = int(p["stride"])
        d1 = float(p["conv1"]["act_delta"])
        self.conv1 = _layer(p["conv1"], ksize=3, stride=self.stride, pad=1, relu=True, in_delta=in_delta, out_delta=d1)
        self.conv2 = _layer(p["conv2"], ksize=3, stride=1, pad=1, relu=True, in_delta=d1, out_delta=self.out_delta)
        self.down = None
        if "down" in p:
            self.down = _layer(p["down"], ksize=1, stride=self.stride, pad=0, relu=False, in_delta=in_delta,
                               out_delta=1.0)
        self.cin, self.cout = self.conv1.cin, self.conv2.cout

    def forward(self, x: torch.Tensor, rowsum: torch.Tensor | None = None, want_rowsum: bool = False,
                taps: dict | None = None):
        """x uint8 NHWC [n, H, W, cin] (scale in_delta) -> uint8 NHWC [n, Ho, Wo, cout] (scale out_delta)."""
        n, h, w, _ = x.shape
        dev = x.device
        if rowsum is None:
            rowsum = E.rowsum_u8(x, 0, self.cin)
        ho, wo = self.conv1.out_shape(h, w)
        rs1 = _zeros.take((n, ho, wo), dev)
        q1 = self.conv1.forward(x, rowsum_in=[rowsum], rowsum_out=rs1)
        rs_out = _zeros.take((n, ho, wo), dev) if want_rowsum else None
        if self.down is not None:
            res = torch.empty((n, ho, wo, self.cout), dtype=torch.float32, device=dev)
            self.down.forward(x, rowsum_in=[rowsum], out_f32=res)
            out = self.conv2.forward(q1, rowsum_in=[rs1], residual=res, rowsum_out=rs_out)
        else:
            res = None
            out = self.conv2.forward(q1, rowsum_in=[rs1], residual=x, res_delta=self.in_delta, rowsum_out=rs_out)
        if taps is not None:
            taps.update(q1=q1, res=res)
        return (out, rs_out) if want_rowsum else out


class ResNetBackboneEngine:
    """The agent-side ``QuantResNetBEVBackbone`` (single stage of BasicBlocks) on libqv2x: uint8 BEV codes in (scale
    ``in_delta`` = the PointPillar block quantizer), uint8 feature codes out (scale ``out_delta`` = the last block's
    quantizer) -- the codebook encoder's input.  params: ``QuantResNetBEVBackbone.export_params()``."""

    def __init__(self, block_params: list, in_delta: float):
        self.blocks, d = [], float(in_delta)
        for p in block_params:
            self.blocks.append(BasicBlockEngine(p, d))
            d = float(p["out_delta"])
        self.in_delta, self.out_delta = float(in_delta), d
        self.cin, self.cout = self.blocks[0].cin, self.blocks[-1].cout

    def forward_u8(self, x: torch.Tensor, rowsum: torch.Tensor | None = None) -> torch.Tensor:
        owner = _zeros.begin(("resnet",) + tuple(x.shape), x.device)
        try:
            rs = rowsum
            for i, b in enumerate(self.blocks):
                last = i == len(self.blocks) - 1
                if last:
                    x = b.forward(x, rowsum=rs)
                else:
                    x, rs = b.forward(x, rowsum=rs, want_rowsum=True)
            return x
        finally:
            if owner:
                _zeros.end()

    def forward_nchw(self, x: torch.Tensor) -> torch.Tensor:
        """Module-boundary drop-in: FP32 NCHW on the input grid -> FP32 NCHW de-quantized output."""
        if not x.is_cuda:
            raise RuntimeError("quantized inference runs on the GPU library only (no CPU fallback)")
        n, c, h, w = x.shape
        xq = E.quantize_nchw_to_nhwc_u8(x.contiguous().float(), self.in_delta)
        y = self.forward_u8(xq)
        return E.dequantize_u8(y, self.out_delta).permute(0, 3, 1, 2).contiguous()


class FirstBottleneckEngine:
    """The first block of stage 0 (64 -> 64 channels, stride 1, identity shortcut) on FP32 input features.  conv1 runs
    on the FP32 GEMM kernel of the detection heads (bias first, one fma per input channel in ascending order), in
    column chunks of 64, and its output is quantized by the NCHW -> NHWC converter (the clamp at 0 is the ReLU)."""

    def __init__(self, params: dict):
        p = params
        assert "down" not in p and int(p["stride"]) == 1
        c1 = p["conv1"]
        self.d1, d2 = float(c1["act_delta"]), float(p["conv2"]["act_delta"])
        self.out_delta = float(p["out_delta"])
        w_hat = ((np.asarray(c1["w_int"], np.float32) - np.asarray(c1["w_zp"], np.float32).reshape(-1, 1, 1, 1))
                 * np.asarray(c1["w_delta"], np.float32).reshape(-1, 1, 1, 1)).reshape(c1["w_int"].shape[0], -1)
        bias = np.zeros(w_hat.shape[0], np.float32) if c1.get("bias") is None else np.asarray(c1["bias"], np.float32)
        self.width, self.cin = w_hat.shape
        assert self.width % 64 == 0
        self.conv1 = E.HeadsEngine(w_hat, bias)            # one launch: 72-column output chunks are grid.y
        self._w_hat, self._bias = w_hat, bias
        self.fold = None                                   # attach_decode_fold: decode . conv1 . quantizer from codes
        self.conv2 = _layer(p["conv2"], ksize=3, stride=1, pad=1, relu=True, in_delta=self.d1, out_delta=d2,
                            groups=int(p["groups"]))
        self.conv3 = _layer(p["conv3"], ksize=1, stride=1, pad=0, relu=True, in_delta=d2, out_delta=self.out_delta)
        self.cout = self.conv3.cout
        assert self.cout == self.cin, "identity shortcut"

    def attach_decode_fold(self, codebook) -> bool:
        """When x is the decode of code planes of `codebook` (the pyramid model's ego stage), conv1 and its quantizer
        fold over the codeword tables: q1 comes straight from the codes (qv2x_decode_linear), without the FP32 GEMM
        on the decoded features and the transposing quantizer pass.  Returns whether the fold is supported."""
        self.fold = None
        if codebook.channel == self.cin and E.DecodeLinearEngine.supported(codebook, self.width):
            self.fold = E.DecodeLinearEngine(codebook, self._w_hat, self._bias, self.d1)
        return self.fold is not None

    def forward(self, x: torch.Tensor, want_rowsum: bool = False, taps: dict | None = None, q1_override=None,
                codes: torch.Tensor | None = None):
        """x float32 NHWC [n, H, W, cin] -> uint8 NHWC [n, H, W, cout] (scale out_delta).  codes: the uint8 code
        planes [levels, m, n*H*W] x was decoded from (x is then only read as the shortcut of conv3)."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[-1] == self.cin
        n, h, w, _ = x.shape
        dev = x.device
        if codes is not None and self.fold is not None and q1_override is None:
            q1, rs1 = self.fold.forward(codes, n * h * w)
            q1 = q1.view(n, h, w, self.width)
            rs1 = rs1.view(n, h, w)
            if taps is not None:
                taps["q1_first"] = q1
        else:
            planar = torch.empty((self.width, n * h * w), dtype=torch.float32, device=dev)
            self.conv1.forward(x, out=planar)
            q1 = E.quantize_nchw_to_nhwc_u8(planar.view(1, self.width, n * h, w), self.d1).view(n, h, w, self.width)
            if taps is not None:
                taps["q1_first"] = q1
            if q1_override is not None:            # test hook: teacher-force the only FP32-accumulated codes
                q1 = q1_override
            rs1 = E.rowsum_u8(q1, 0, self.width)
        rs2 = _zeros.take((n, h, w), dev)
        q2 = self.conv2.forward(q1, rowsum_in=[rs1], rowsum_out=rs2)
        rs_out = _zeros.take((n, h, w), dev) if want_rowsum else None
        out = self.conv3.forward(q2, rowsum_in=[rs2], residual=x, rowsum_out=rs_out)
        return (out, rs_out) if want_rowsum else out


class OccupancyHead:
    """``single_head_i``: nn.Conv2d(C, 1, 1) wrapped in a QuantModule whose output feeds sigmoid directly
    (quant_block.py:474-478, 516-520).  The one real output channel is padded to the 64-column tile of the GEMM
    (columns 1..63 carry the weight zero-point, i.e. zero weights)."""

    def __init__(self, w_int, w_delta, w_zp, bias, in_delta, w_bits=8, act=None):
        # act = (delta, zero_point, n_bits) of the head's own output quantizer, or None when it is disabled
        self.act = None if act is None else (float(act[0]), float(act[1]), int(act[2]))
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
        occ = buf[..., 0].contiguous()
        if self.act is not None:
            # UniformAffineQuantizer.forward (quant_layer.py:132-148) on the one-channel logit map: true division,
            # round half to even, clamp, de-quantize
            d, zp, bits = self.act
            occ = (torch.clamp(torch.round(occ / d) + zp, 0, float(2 ** bits - 1)) - zp) * d
        return occ


def weighted_fuse_level(codes: torch.Tensor, delta: float, occ: torch.Tensor, affine) -> torch.Tensor:
    """One level of QuantPyramidFusion.forward_collab (quant_block.py:516-539): codes uint8 NHWC [N, H, W, C] of the
    level's features (scale delta, agent 0 = ego), occ float32 [N, H, W] logits of single_head_i, affine [N, 2, 3]
    -> fused float32 [H, W, C]."""
    if codes.shape[-1] % 4 == 0 and codes.is_contiguous():
        return E.fuse_weighted_u8(codes, delta, occ, affine, score_is_logit=True)     # de-quantizes on load
    feat = E.dequantize_u8(codes, delta)
    return E.fuse_weighted(feat, occ, affine, score_is_logit=True)


class DeblockF32:
    """QuantModule(ConvTranspose2d(cin, cout, s, stride=s)) + ReLU + act quantizer on an FP32 input (the fused map of
    a level is off every quantization grid).  Output pixel (y*s + dy, x*s + dx) is a 1x1 conv of input pixel (y, x)
    with the weight slice [:, :, dy, dx]: one FP32 GEMM with s*s*cout columns (the heads kernel, 72-column output
    chunks as grid.y), then the quantizing converter and a pixel shuffle."""

    def __init__(self, up: dict):
        self.s = int(up["stride"])
        self.delta = float(up["act_delta"])
        w_hat = ((np.asarray(up["w_int"], np.float32) - np.asarray(up["w_zp"], np.float32).reshape(-1, 1, 1, 1))
                 * np.asarray(up["w_delta"], np.float32).reshape(-1, 1, 1, 1))         # [cin, cout, s, s]
        self.cin, self.cout = w_hat.shape[:2]
        rows = np.ascontiguousarray(w_hat.transpose(2, 3, 1, 0).reshape(self.s * self.s * self.cout, self.cin))
        b = np.zeros(self.cout, np.float32) if up.get("bias") is None else np.asarray(up["bias"], np.float32)
        bias = np.tile(b, self.s * self.s)
        self.gemm = E.HeadsEngine(rows, bias)              # one launch for all s*s sub-positions

    def forward(self, fused: torch.Tensor, out: torch.Tensor, out_cbase: int = 0) -> torch.Tensor:
        """fused float32 [h, w, cin] -> codes written to out[:, :, out_cbase : out_cbase + cout] (uint8 [h*s, w*s, C])."""
        assert fused.is_cuda and fused.dtype == torch.float32 and fused.is_contiguous() and fused.shape[-1] == self.cin
        h, w, _ = fused.shape
        s, n_col = self.s, self.s * self.s * self.cout
        if (self.cout % 8 == 0 and out_cbase % 8 == 0 and out.shape[-1] % 8 == 0 and out.is_contiguous()
                and os.environ.get("QV2X_DEBLOCK_CHAIN", "0") != "1"):
            # quantizer + pixel shuffle in the GEMM's epilogue: the same codes as the three-step path below
            return E.heads_forward_deconv_u8(self.gemm, fused, s, self.delta, out, out_cbase)
        planar = torch.empty((n_col, h * w), dtype=torch.float32, device=fused.device)
        self.gemm.forward(fused, out=planar)
        q = E.quantize_nchw_to_nhwc_u8(planar.view(1, n_col, h, w), self.delta)        # [1, h, w, s*s*cout]
        q = q.view(h, w, s, s, self.cout).permute(0, 2, 1, 3, 4).reshape(h * s, w * s, self.cout)
        out[:, :, out_cbase:out_cbase + self.cout].copy_(q)
        return out


class PyramidBackboneEngine:
    """``QuantPyramidFusion.forward_collab`` (quant_block.py:504-541) up to the fused per-level features.

    params: ``'l{i}.b{j}'`` -> bottleneck dicts (see BottleneckEngine) and ``'head{i}'`` -> dict(w_int, w_delta, w_zp,
    bias) of ``single_head_i``; layer_nums: blocks per stage."""

    def __init__(self, params: dict, layer_nums):
        self.stages, self.heads, self.deltas = [], [], []
        delta = None
        for li, nb in enumerate(layer_nums):
            blocks = []
            for bi in range(nb):
                p = params[f"l{li}.b{bi}"]
                blk = FirstBottleneckEngine(p) if delta is None else BottleneckEngine(p, delta)
                delta = blk.out_delta
                blocks.append(blk)
            self.stages.append(blocks)
            self.deltas.append(delta)
            h = params[f"head{li}"]
            act = (h["act_delta"], h["act_zp"], h["act_bits"]) if "act_delta" in h else None
            self.heads.append(OccupancyHead(h["w_int"], h["w_delta"], h["w_zp"], h.get("bias"), delta, act=act))
        self.deblocks = [DeblockF32(params[f"up{li}"]) for li in range(len(layer_nums)) if f"up{li}" in params]
        self.up_deltas = [d.delta for d in self.deblocks]

    def attach_decode_fold(self, codebook) -> bool:
        """See FirstBottleneckEngine.attach_decode_fold (the first block of stage 0)."""
        ok = False
        for blk in self.stages[0]:
            if isinstance(blk, FirstBottleneckEngine):
                ok = blk.attach_decode_fold(codebook)
        return ok

    def forward_collab(self, x: torch.Tensor, affine, taps: dict | None = None, q1_override=None,
                       codes: torch.Tensor | None = None):
        """x float32 NHWC [N, H, W, 64] decoded features of the N agents in range (agent 0 = ego); affine [N, 2, 3]
        = normalize_pairwise_tfm(...)[b][0, :N].  Returns the fused float32 [h_i, w_i, C_i] map of every level."""
        if not (isinstance(affine, torch.Tensor) and affine.is_cuda):
            affine = torch.as_tensor(np.asarray(affine, dtype=np.float32)).to(x.device)
        owner = _zeros.begin(("collab",) + tuple(x.shape), x.device)
        try:
            return self._forward_collab(x, affine, taps, q1_override, codes)
        finally:
            if owner:
                _zeros.end()

    def _forward_collab(self, x, affine, taps, q1_override, codes):
        fused = []
        cur, rs = x, None
        for li, blocks in enumerate(self.stages):
            for blk in blocks:
                if isinstance(blk, FirstBottleneckEngine):
                    cur, rs = blk.forward(cur, want_rowsum=True, taps=taps, q1_override=q1_override, codes=codes)
                else:
                    cur, rs = blk.forward(cur, rowsum=rs, want_rowsum=True)
            occ = self.heads[li].forward(cur, rowsum=rs)
            fused.append(weighted_fuse_level(cur, self.deltas[li], occ, affine))
            if taps is not None:
                taps[f"l{li}.codes"], taps[f"l{li}.occ"] = cur, occ
        return fused

    def decode_multiscale_feature(self, fused):
        """The deblocks + channel concat of QuantResNetBEVBackbone.decode_multiscale_feature (quant_block.py:441-458):
        fused level maps -> uint8 codes [1, H, W, sum cout] (level i's channels carry scale up_deltas[i])."""
        assert len(self.deblocks) == len(fused)
        h, w = fused[0].shape[0] * self.deblocks[0].s, fused[0].shape[1] * self.deblocks[0].s
        cat = torch.empty((h, w, sum(d.cout for d in self.deblocks)), dtype=torch.uint8, device=fused[0].device)
        base = 0
        for d, f in zip(self.deblocks, fused):
            d.forward(f, cat, base)
            base += d.cout
        return cat.unsqueeze(0)
