"""``MaxFusion`` / ``AttFusion`` / ``weighted_fuse`` -- mirrors of opencood/models/fuse_modules/fusion_in_one.py:87-151 with the
same call signature ``fusion_net(x[sum(N), C, H, W], record_len[B], affine_matrix[B, L, L, 2, 3]) -> [B, C, H, W]``.
The warp + fusion runs as one libqv2x kernel per frame (no torch compute, no CPU fallback)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import engine as E


def regroup(x, record_len):
    cum = torch.cumsum(record_len, dim=0)
    return torch.tensor_split(x, cum[:-1].cpu())


class _FusionBase(nn.Module):
    mode = "max"

    def forward(self, x, record_len, affine_matrix):
        if not x.is_cuda:
            raise RuntimeError("fusion runs on the GPU library only (no CPU fallback)")
        outs = []
        for b, xb in enumerate(regroup(x, record_len)):
            n = xb.shape[0]
            feat = E.nchw_to_nhwc_f32(xb.contiguous().float())
            aff = affine_matrix[b][0, :n].to(device=x.device, dtype=torch.float32).contiguous()
            fused = E.fuse(feat, aff, self.mode)                       # [H, W, C]
            outs.append(E.nhwc_to_nchw_f32(fused.unsqueeze(0))[0])
        return torch.stack(outs)


class MaxFusion(_FusionBase):
    mode = "max"


class AttFusion(_FusionBase):
    mode = "att"

    def __init__(self, feature_dims=None):
        super().__init__()
        self.feature_dims = feature_dims


def weighted_fuse(x, score, record_len, affine_matrix, align_corners=False, score_is_logit=False):
    """Mirror of ``weighted_fuse`` (opencood/models/fuse_modules/pyramid_fuse.py:17-62), the per-level fusion of the
    pyramid model: x [sum(N), C, H, W], score [sum(N), 1, H, W], record_len [B], affine_matrix [B, L, L, 2, 3]
    -> [B, C, H, W].  With ``score_is_logit`` the occupancy logits of ``single_head_i`` are passed instead and the
    kernel applies ``sigmoid + 1e-4`` itself (QuantPyramidFusion.forward_collab, quant_block.py:516-520)."""
    if align_corners:
        raise NotImplementedError("the warp kernel implements align_corners=False (the default; no shipped "
                                  "config sets it)")
    if not x.is_cuda:
        raise RuntimeError("fusion runs on the GPU library only (no CPU fallback)")
    outs = []
    for b, (xb, sb) in enumerate(zip(regroup(x, record_len), regroup(score, record_len))):
        n = xb.shape[0]
        feat = E.nchw_to_nhwc_f32(xb.contiguous().float())
        aff = affine_matrix[b][0, :n].to(device=x.device, dtype=torch.float32).contiguous()
        fused = E.fuse_weighted(feat, sb.reshape(n, *sb.shape[-2:]).contiguous().float(), aff, score_is_logit)
        outs.append(E.nhwc_to_nchw_f32(fused.unsqueeze(0))[0])
    return torch.stack(outs)
