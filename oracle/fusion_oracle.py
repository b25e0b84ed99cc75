"""ORACLE (test infrastructure): ego-side warp + fusion + heads restated in float64 numpy.

Follows
  * normalize_pairwise_tfm      opencood/utils/transformation_utils.py:68-92
  * warp_affine_simple          opencood/models/sub_modules/torch_transformation_utils.py:323-332
                                (F.affine_grid + F.grid_sample, bilinear, zeros padding, align_corners=False)
  * MaxFusion.forward           opencood/models/fuse_modules/fusion_in_one.py:87-124
  * AttFusion.forward           opencood/models/fuse_modules/fusion_in_one.py:126-151 (+ ScaledDotProductAttention :14-45)
  * weighted_fuse               opencood/models/fuse_modules/pyramid_fuse.py:17-62, with the score preparation of
                                QuantPyramidFusion.forward_collab (opencood/quant/quant_block.py:516-539)
Pinned against the reference's torch implementation by tests/golden/fusion_*.npz.
"""
from __future__ import annotations

import numpy as np


def normalize_pairwise_tfm(pairwise_t_matrix, H, W, discrete_ratio, downsample_rate=1):
    """[B, L, L, 4, 4] poses -> [B, L, L, 2, 3] normalized affine (H, W in metres on this path)."""
    t = np.asarray(pairwise_t_matrix, dtype=np.float64)
    aff = t[:, :, :, [0, 1], :][:, :, :, :, [0, 1, 3]].copy()
    aff[..., 0, 1] = aff[..., 0, 1] * H / W
    aff[..., 1, 0] = aff[..., 1, 0] * W / H
    aff[..., 0, 2] = aff[..., 0, 2] / (downsample_rate * discrete_ratio * W) * 2
    aff[..., 1, 2] = aff[..., 1, 2] / (downsample_rate * discrete_ratio * H) * 2
    return aff


def warp(feat, aff):
    """feat [N, H, W, C] float, aff [N, 2, 3] -> warped [N, H, W, C] (float64)."""
    feat = np.asarray(feat, np.float64)
    N, H, W, C = feat.shape
    j = np.arange(W)
    i = np.arange(H)
    xn = (2.0 * j + 1.0) / W - 1.0
    yn = (2.0 * i + 1.0) / H - 1.0
    out = np.zeros_like(feat)
    for a in range(N):
        M = np.asarray(aff[a], np.float64)
        xs = M[0, 0] * xn[None, :] + M[0, 1] * yn[:, None] + M[0, 2]
        ys = M[1, 0] * xn[None, :] + M[1, 1] * yn[:, None] + M[1, 2]
        ix = ((xs + 1.0) * W - 1.0) / 2.0
        iy = ((ys + 1.0) * H - 1.0) / 2.0
        x0 = np.floor(ix).astype(np.int64)
        y0 = np.floor(iy).astype(np.int64)
        tx = ix - x0
        ty = iy - y0
        acc = np.zeros((H, W, C))
        for dy in (0, 1):
            for dx in (0, 1):
                xx, yy = x0 + dx, y0 + dy
                w = (tx if dx else 1.0 - tx) * (ty if dy else 1.0 - ty)
                valid = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
                xc, yc = np.clip(xx, 0, W - 1), np.clip(yy, 0, H - 1)
                acc += np.where(valid, w, 0.0)[..., None] * feat[a][yc, xc]
        out[a] = acc
    return out


def max_fusion(feat, aff):
    return warp(feat, aff).max(axis=0)


def att_fusion(feat, aff):
    x = warp(feat, aff)                                   # [N, H, W, C]
    C = x.shape[-1]
    score = (x[0][None] * x).sum(-1) / np.sqrt(C)         # [N, H, W]  ego query against every agent
    score = score - score.max(axis=0, keepdims=True)
    p = np.exp(score)
    p = p / p.sum(axis=0, keepdims=True)
    return (p[..., None] * x).sum(axis=0)


def weighted_fusion(feat, score, aff, score_is_logit=True):
    """feat [N, H, W, C], score [N, H, W] (occupancy logits, or scores), aff [N, 2, 3] -> [H, W, C]."""
    s = np.asarray(score, np.float64)
    if score_is_logit:
        s = 1.0 / (1.0 + np.exp(-s)) + np.float64(np.float32(1e-4))
    x = warp(feat, aff)                                   # [N, H, W, C]
    ws = warp(s[..., None], aff)[..., 0]                  # [N, H, W]
    ws = np.where(ws == 0.0, -np.inf, ws)
    mx = ws.max(axis=0, keepdims=True)
    with np.errstate(invalid="ignore"):
        p = np.exp(ws - mx)
        p = p / p.sum(axis=0, keepdims=True)
    p = np.where(np.isnan(p), 0.0, p)                     # every agent excluded -> 0
    return (p[..., None] * x).sum(axis=0)


def heads(fused, w, b):
    """fused [H, W, C], w [Cout, C], b [Cout] -> [Cout, H, W]."""
    y = np.asarray(fused, np.float64) @ np.asarray(w, np.float64).T + np.asarray(b, np.float64)
    return y.transpose(2, 0, 1)
