"""Deterministic synthetic weights and inputs: there is no network for checkpoints or
datasets, so benches and tests use seeded He-normal weights that keep activations alive through all 24
layers, and V2X-Real-like sparse BEV / pillar inputs.  numpy's PCG64 stream is stable across machines,
which lets the golden-vector generator (run against the reference) and the tests rebuild identical models."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


@torch.no_grad()
def seeded_init(model: nn.Module, seed: int, skip=("codebook",), head_cls_bias: float | None = -4.0):
    """He-normal(fan_in) weights, U(-0.1, 0.1) biases for every Conv2d / ConvTranspose2d / Linear outside
    `skip`; BatchNorm left at its constructor state.  Modules are visited in sorted-name order."""
    rng = np.random.default_rng(seed)
    for name, m in sorted(model.named_modules(), key=lambda kv: kv[0]):
        if any(name == s or name.startswith(s + ".") or ("." + s + ".") in ("." + name + ".") for s in skip):
            continue
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
            w = m.weight
            fan_in = w.shape[0] if isinstance(m, nn.ConvTranspose2d) else int(np.prod(w.shape[1:]))
            w.copy_(torch.from_numpy(rng.normal(0.0, np.sqrt(2.0 / fan_in), size=tuple(w.shape)).astype(np.float32)))
            if m.bias is not None:
                m.bias.copy_(torch.from_numpy(rng.uniform(-0.1, 0.1, size=tuple(m.bias.shape)).astype(np.float32)))
                if head_cls_bias is not None and name.rsplit(".", 1)[-1] == "cls_head":
                    m.bias.fill_(head_cls_bias)
    return model


def synthetic_bev(seed: int, n_agents: int, H: int = 200, W: int = 704, C: int = 64, pillars: int = 6000):
    """uint8 NHWC [n, H, W, C] BEV codes: `pillars` random occupied cells per agent, ~50 % zeros inside."""
    rng = np.random.default_rng(seed)
    x = np.zeros((n_agents, H * W, C), dtype=np.uint8)
    for a in range(n_agents):
        cells = rng.permutation(H * W)[:min(pillars, H * W)]
        v = rng.integers(0, 256, size=(cells.size, C), dtype=np.uint8)
        v[rng.random((cells.size, C)) > 0.5] = 0
        x[a, cells] = v
    return x.reshape(n_agents, H, W, C)


def synthetic_poses(n_agents: int, max_cav: int = 5):
    """pairwise_t_matrix [1, L, L, 4, 4]: ego->agent j translated (3j, -1.5j) m and yawed 5j degrees."""
    L = max(n_agents, max_cav)
    t = np.tile(np.eye(4), (1, L, L, 1, 1))
    for j in range(1, n_agents):
        th = np.deg2rad(5.0 * j)
        t[0, 0, j, :2, :2] = [[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]
        t[0, 0, j, 0, 3] = 3.0 * j
        t[0, 0, j, 1, 3] = -1.5 * j
    return t


@torch.no_grad()
def seeded_init_codebook(codebook: nn.Module, seed: int):
    """Constructor-style init of a UMGMQuantizer (reference codebook.py:290-324) from a numpy stream, so the
    reference module and its mirror hold identical parameters: Linear U(+-1/sqrt(fan_in)), codebooks
    N(0, sqrt(2 / (5 d))).  Parameters are visited in sorted-name order."""
    rng = np.random.default_rng(seed)
    for name, p in sorted(codebook.named_parameters(), key=lambda kv: kv[0]):
        if name.endswith("_codebook"):
            d = p.shape[-1]
            p.copy_(torch.from_numpy(rng.normal(0.0, np.sqrt(2.0 / (5.0 * d)), size=tuple(p.shape)).astype(np.float32)))
        elif name.endswith(".weight") or name.endswith(".bias"):
            fan_in = p.shape[-1] if p.dim() == 2 else None
            if fan_in is None:                      # bias: same bound as its weight (square layers here)
                fan_in = p.shape[0]
            b = 1.0 / np.sqrt(fan_in)
            p.copy_(torch.from_numpy(rng.uniform(-b, b, size=tuple(p.shape)).astype(np.float32)))
    if hasattr(codebook, "reset_engine"):
        codebook.reset_engine()
    return codebook


def synthetic_pillars(seed: int, n_agents: int, lidar_range, voxel_size, pillars: int = 6000, max_points: int = 32):
    """Pillar-level inputs in the reference's dict schema: voxel_features [M, P, 4],
    voxel_coords [M, 4] = (agent, z, y, x), voxel_num_points [M]."""
    rng = np.random.default_rng(seed)
    nx = int(round((lidar_range[3] - lidar_range[0]) / voxel_size[0]))
    ny = int(round((lidar_range[4] - lidar_range[1]) / voxel_size[1]))
    feats, coords, nums = [], [], []
    for a in range(n_agents):
        M = min(pillars, nx * ny)
        cells = rng.permutation(nx * ny)[:M]
        yy, xx = cells // nx, cells % nx
        npts = rng.integers(1, max_points + 1, size=M)
        f = np.zeros((M, max_points, 4), np.float32)
        f[..., 0] = (xx[:, None] + rng.random((M, max_points))) * voxel_size[0] + lidar_range[0]
        f[..., 1] = (yy[:, None] + rng.random((M, max_points))) * voxel_size[1] + lidar_range[1]
        f[..., 2] = rng.uniform(lidar_range[2], lidar_range[5], size=(M, max_points))
        f[..., 3] = rng.random((M, max_points))
        f *= (np.arange(max_points)[None, :] < npts[:, None])[..., None]
        feats.append(f)
        coords.append(np.stack([np.full(M, a), np.zeros(M, np.int64), yy, xx], 1))
        nums.append(npts)
    return (np.concatenate(feats).astype(np.float32), np.concatenate(coords).astype(np.int32),
            np.concatenate(nums).astype(np.int32))
