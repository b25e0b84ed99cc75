"""CPU oracle of the quantized PointPillars front end (TEST INFRASTRUCTURE ONLY -- never imported by the product path).

Restates, in numpy, what the reference computes from pillars to the BEV map and what csrc/pillar.cu evaluates:

* QuantPillarVFE.forward  (opencood/quant/quant_block.py:666-716): point decoration -- offsets to the pillar's point
  mean and to the cell centre, padded points zeroed;
* QuantPFNLayer.forward   (quant_block.py:611-630) with its QuantModule Linear (quant_layer.py:391-410, BN folded by
  fold_bn.py:161-175): fake-quant weights, the Linear's own activation quantizer BEFORE the ReLU, ReLU, the block's
  activation quantizer, max over the 32 points;
* PointPillarScatter.forward (opencood/models/sub_modules/point_pillar_scatter.py:19-75).

Normative FP32 order (the kernel follows it bit for bit; the reference's torch kernels use their own summation order,
so the comparison against the reference's golden BEV codes is a tolerance check):
    sum_xyz : xor-butterfly over the 32 point slots (16, 8, 4, 2, 1); mean = sum / float(num_points)
    y[o]    : bias[o], then y = fma(f[c], w[o][c], y) for c = 0..9
    both quantizers use true division and round-half-even; the max over points is taken on y (the quantizers and the
    ReLU are monotone non-decreasing, so this equals the reference's max over the quantized values).
Pinned by tests/golden/e2e_*.npz (voxel_* inputs and bev_codes produced by the unmodified reference).
"""
import numpy as np

from .int_oracle import fma32

f32 = np.float32


def _butterfly_sum(v: np.ndarray) -> np.ndarray:
    """v float32 [M, 32] -> [M]: the sum every lane holds after xor-shuffle reductions with offsets 16, 8, 4, 2, 1."""
    s = v.astype(f32).copy()
    lanes = np.arange(32)
    for off in (16, 8, 4, 2, 1):
        s = (s + s[:, lanes ^ off]).astype(f32)
    return s[:, 0]


def decorate(voxel_features, voxel_coords, voxel_num_points, voxel_size, offset):
    """[M, 32, 4] -> [M, 32, 10] float32 (quant_block.py:676-709)."""
    vf = np.asarray(voxel_features, dtype=f32)
    n = np.asarray(voxel_num_points).astype(np.int64)
    fn = n.astype(f32)
    mean = np.stack([(_butterfly_sum(vf[:, :, i]) / fn).astype(f32) for i in range(3)], axis=1)      # [M, 3]
    c = np.asarray(voxel_coords)
    centre = np.stack([(c[:, 3].astype(f32) * f32(voxel_size[0]) + f32(offset[0])).astype(f32),
                       (c[:, 2].astype(f32) * f32(voxel_size[1]) + f32(offset[1])).astype(f32),
                       (c[:, 1].astype(f32) * f32(voxel_size[2]) + f32(offset[2])).astype(f32)], axis=1)
    f = np.concatenate([vf, (vf[:, :, :3] - mean[:, None, :]).astype(f32),
                        (vf[:, :, :3] - centre[:, None, :]).astype(f32)], axis=2)
    live = np.arange(32)[None, :] < n[:, None]
    return np.where(live[:, :, None], f, f32(0)).astype(f32)


def pfn_codes(feats, w_hat, bias, pre_quant, out_quant):
    """feats float32 [M, 32, 10] -> uint8 [M, 64] pillar codes."""
    w = np.asarray(w_hat, dtype=f32)
    b = np.zeros(w.shape[0], f32) if bias is None else np.asarray(bias, dtype=f32)
    y = np.broadcast_to(b[None, None, :], feats.shape[:2] + (w.shape[0],)).astype(f32)
    for c in range(feats.shape[2]):
        y = fma32(feats[:, :, c:c + 1], w[None, None, :, c], y)
    y = y.max(axis=1)                                               # [M, 64]
    if pre_quant is not None:
        d1, z1, b1 = f32(pre_quant[0]), f32(pre_quant[1]), int(pre_quant[2])
        t = np.clip(np.rint(y / d1) + z1, f32(0), f32(2 ** b1 - 1)).astype(f32)
        y = ((t - z1) * d1).astype(f32)
    y = np.maximum(y, f32(0))
    d2, z2, b2 = f32(out_quant[0]), f32(out_quant[1]), int(out_quant[2])
    q = np.clip(np.rint(y / d2) + z2, f32(0), f32(2 ** b2 - 1))
    return q.astype(np.uint8)


def pillar_bev(spec: dict, voxel_features, voxel_coords, voxel_num_points, batch: int) -> np.ndarray:
    """spec = quantv2x_b200.export.pillar_spec(...).  Returns uint8 BEV codes [batch, ny, nx, 64] (empty cells 0)."""
    feats = decorate(voxel_features, voxel_coords, voxel_num_points, spec["voxel_size"], spec["offset"])
    codes = pfn_codes(feats, spec["w_hat"], spec["bias"], spec["pre_quant"], spec["out_quant"])
    bev = np.zeros((batch, spec["ny"], spec["nx"], codes.shape[1]), np.uint8)
    c = np.asarray(voxel_coords)
    bev[c[:, 0], c[:, 2], c[:, 3]] = codes
    return bev
