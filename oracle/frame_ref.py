"""ORACLE / CPU BASELINE (test infrastructure): the reference's own CPU arithmetic for one cooperative frame,
restated in PyTorch FP32 from a plain-numpy model spec (no import of the product package, no import of
/root/reference -- this file travels to the GPU box, the reference does not).

Follows, operation for operation, what the reference executes on the path:
  QuantModule.forward          opencood/quant/quant_layer.py:391-410   fake-quant W (every call) -> F.conv2d /
                                                                       F.conv_transpose2d -> ReLU -> fake-quant act
  QuantBaseBEVBackbone.forward opencood/quant/quant_block.py:280-303   blocks, deblocks, torch.cat
  QuantDownsampleConv.forward  opencood/quant/quant_block.py:583-586
  UMGMQuantizer.encode/decode  opencood/models/sub_modules/codebook.py:330-343 (FP32 linears, einsum distance, argmin)
  warp_affine_simple           opencood/models/sub_modules/torch_transformation_utils.py:323-332
  MaxFusion / AttFusion        opencood/models/fuse_modules/fusion_in_one.py:87-151
  heads                        1x1 convs with fake-quant weights, FP32 activations (quant_model.py:129-136)

Used as (a) the tolerance-tier reference of the GPU end-to-end test and (b) bench.py's cpu_baseline /
`--impl reference` arm (kind "port": the Python reference cannot be compiled or shipped).
Pinned against the real reference by tests/golden/ (oracle/gen_golden.py).

Model spec (all numpy):
  spec["bev_delta"]                 float
  spec["blocks"]   = [[layer, ...], ...]      layer = dict(kind, w, bias, w_bits, w_delta, w_zp, stride, pad,
  spec["deblocks"] = [layer, ...]                          act_delta, act_zp, act_bits, relu)
  spec["shrinker"] = [layer, ...]
  spec["codebook"] = dict(codebooks=[...], heads=[{name: (W, b) | None}, ...])
  spec["heads"]    = dict(w=[Cout, C] de-quantized weights, b=[Cout])
  spec["fusion"]   = "max" | "att"
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _fq(x, delta, zp, bits):
    q = torch.clamp(torch.round(x / delta) + zp, 0, 2 ** bits - 1)
    return (q - zp) * delta


def quant_layer(x, L):
    """One QuantModule.forward in fake-quant FP32 (weights re-quantized on every call, as the reference does)."""
    w = torch.from_numpy(L["w"])
    shape = [-1] + [1] * (w.dim() - 1)
    d = torch.from_numpy(np.asarray(L["w_delta"], np.float32)).reshape(shape)
    z = torch.from_numpy(np.asarray(L["w_zp"], np.float32)).reshape(shape)
    w_hat = _fq(w, d, z, L["w_bits"])
    b = None if L["bias"] is None else torch.from_numpy(L["bias"])
    if L["kind"] == 0:
        y = F.conv2d(x, w_hat, b, stride=L["stride"], padding=L["pad"])
    else:
        y = F.conv_transpose2d(x, w_hat, b, stride=L["stride"])
    if L["relu"]:
        y = torch.relu(y)
    if L.get("act_delta") is not None:
        y = _fq(y, L["act_delta"], L["act_zp"], L["act_bits"])
    return y


def backbone_shrinker(spec, x):
    """x: float32 NCHW BEV (on the uint8 grid bev_delta).  Returns float32 NCHW shrinker output."""
    ups = []
    for blk, de in zip(spec["blocks"], spec["deblocks"]):
        for L in blk:
            x = quant_layer(x, L)
        ups.append(quant_layer(x, de))
    x = torch.cat(ups, dim=1)
    for L in spec["shrinker"]:
        x = quant_layer(x, L)
    return x


def _lin(wb, x):
    return F.linear(x, torch.from_numpy(wb[0]), torch.from_numpy(wb[1]))


def codebook_encode(cb, x):
    """FP32 restatement of UMGMQuantizer.encode.  x [n, C] -> list of LongTensor [n, m]."""
    codes = []
    levels = len(cb["codebooks"])
    for l in range(levels):
        book = torch.from_numpy(cb["codebooks"][l])
        m, k, d = book.shape
        h = cb["heads"][l]
        z = _lin(h["latentStageEncoder"], x)
        q = _lin(h["quantizationHead"], z).reshape(-1, m, d)
        x2 = (q ** 2).sum(2, keepdim=True)
        c2 = (book ** 2).sum(-1)
        inter = torch.einsum("nmd,mkd->nmk", q, book)
        code = (x2 + c2 - 2 * inter).argmin(-1)
        codes.append(code)
        if h["latentHead"] is not None:
            ix = torch.arange(m).expand_as(code)
            x = _lin(h["latentHead"], z) - book[ix, code].reshape(code.shape[0], -1)
    return codes


def codebook_decode(cb, codes):
    former = None
    for l in reversed(range(len(codes))):
        book = torch.from_numpy(cb["codebooks"][l])
        m = book.shape[0]
        h = cb["heads"][l]
        code = codes[l]
        ix = torch.arange(m).expand_as(code)
        q = _lin(h["dequantizationHead"], book[ix, code].reshape(code.shape[0], -1))
        xhat = q if (h["sideHead"] is None or former is None) else q + _lin(h["sideHead"], former)
        former = _lin(h["restoreHead"], xhat)
    return former


def warp_affine_simple(src, M, dsize):
    grid = F.affine_grid(M, [src.shape[0], src.shape[1], dsize[0], dsize[1]], align_corners=False).to(src)
    return F.grid_sample(src, grid, align_corners=False)


def fuse(feat, aff, mode):
    """feat [N, C, H, W] float32, aff [N, 2, 3] -> [C, H, W]."""
    N, C, H, W = feat.shape
    x = warp_affine_simple(feat, aff, (H, W))
    if mode == "max":
        return torch.max(x, dim=0)[0]
    xx = x.view(N, C, -1).permute(2, 0, 1)
    score = torch.bmm(xx, xx.transpose(1, 2)) / np.sqrt(C)
    ctx = torch.bmm(F.softmax(score, -1), xx)
    return ctx.permute(1, 2, 0).view(N, C, H, W)[0]


@torch.no_grad()
def agent_forward(spec, bev_u8):
    """bev_u8 uint8 [n, H, W, C] -> (feature float32 [n, C, h, w], codes list of [n*h*w, m])."""
    x = torch.from_numpy(bev_u8.astype(np.float32) * np.float32(spec["bev_delta"])).permute(0, 3, 1, 2).contiguous()
    feat = backbone_shrinker(spec, x)
    n, C, h, w = feat.shape
    flat = feat.permute(0, 2, 3, 1).contiguous().view(-1, C)
    return feat, codebook_encode(spec["codebook"], flat)


@torch.no_grad()
def ego_forward(spec, codes, n, h, w, aff):
    """codes list of [n*h*w, m]; aff [n, 2, 3] -> (fused [C, h, w], preds [Cout, h, w])."""
    dec = codebook_decode(spec["codebook"], codes)
    C = dec.shape[1]
    feat = dec.view(n, h, w, C).permute(0, 3, 1, 2).contiguous()
    fused = fuse(feat, torch.from_numpy(np.asarray(aff, np.float32)), spec["fusion"])
    hw = torch.from_numpy(spec["heads"]["w"])
    preds = F.conv2d(fused.unsqueeze(0), hw.view(hw.shape[0], hw.shape[1], 1, 1), torch.from_numpy(spec["heads"]["b"]))[0]
    return fused, preds


@torch.no_grad()
def pillar_forward(ps, voxel_features, voxel_coords, voxel_num_points, batch):
    """The quantized PointPillars front end as the reference runs it, in torch FP32 (QuantPillarVFE.forward
    quant_block.py:666-716, QuantPFNLayer.forward :611-630, PointPillarScatter.forward point_pillar_scatter.py:19-75):
    decoration, Linear with fake-quant weights, its own activation quantizer, ReLU, the block quantizer, max over the
    points, scatter into a dense map.  ps = quantv2x_b200.export.pillar_spec(...).  Returns uint8 codes [batch, ny, nx, 64]."""
    vf = torch.from_numpy(np.asarray(voxel_features, np.float32))
    vc = torch.from_numpy(np.asarray(voxel_coords).astype(np.int64))
    n = torch.from_numpy(np.asarray(voxel_num_points).astype(np.int64))
    mean = vf[:, :, :3].sum(1, keepdim=True) / n.view(-1, 1, 1).float()
    vs, off = ps["voxel_size"], ps["offset"]
    centre = torch.stack([vc[:, 3].float() * vs[0] + off[0], vc[:, 2].float() * vs[1] + off[1],
                          vc[:, 1].float() * vs[2] + off[2]], 1).unsqueeze(1)
    f = torch.cat([vf, vf[:, :, :3] - mean, vf[:, :, :3] - centre], -1)
    f = f * (torch.arange(vf.shape[1]).view(1, -1) < n.view(-1, 1)).unsqueeze(-1).float()
    y = F.linear(f, torch.from_numpy(ps["w_hat"]), None if ps["bias"] is None else torch.from_numpy(ps["bias"]))
    if ps["pre_quant"] is not None:
        y = _fq(y, *ps["pre_quant"])
    y = torch.relu(y)
    y = _fq(y, *ps["out_quant"])
    y = y.max(1).values                                                # [M, 64]
    d, z, _ = ps["out_quant"]
    bev = torch.zeros((batch, ps["ny"], ps["nx"], y.shape[1]), dtype=torch.float32)
    bev[vc[:, 0], vc[:, 2], vc[:, 3]] = y
    return torch.round(bev / d + z).clamp(0, 255).to(torch.uint8).numpy()


@torch.no_grad()
def frame_from_pillars(spec, ps, voxel_features, voxel_coords, voxel_num_points, n, aff):
    """One cooperative frame of n agents from pillars to head maps: the work bench.py's GPU arm does per step."""
    bev = pillar_forward(ps, voxel_features, voxel_coords, voxel_num_points, n)
    return frame_forward(spec, bev, aff)


@torch.no_grad()
def frame_forward(spec, bev_u8, aff):
    feat, codes = agent_forward(spec, bev_u8)
    n, _, h, w = feat.shape
    fused, preds = ego_forward(spec, codes, n, h, w, aff)
    return dict(feat=feat, codes=codes, fused=fused, preds=preds)
