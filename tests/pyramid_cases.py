"""Seeded ResNeXt bottleneck blocks of the pyramid backbone, shared by oracle/gen_golden_pyramid.py
(which runs the reference's QuantBottleneck on them) and the parity tests.  Shapes follow ResNetModified(Bottleneck,
groups=32, width_per_group=4) with Bottleneck.expansion = 1 (pyramid_fuse.py:69-77): width = 2 * planes."""
import numpy as np

GROUPS = 32
# name, inplanes, planes, stride, H, W
BLOCK_CASES = [
    ("bneck_identity", 128, 128, 1, 10, 12),       # a block inside stage 1: identity shortcut on the input grid
    ("bneck_down_s2", 64, 128, 2, 12, 16),         # first block of stage 1: strided 3x3 + 1x1 stride-2 downsample
]
IN_DELTA = np.float32(0.047)
# BasicBlock (two 3x3 convs) of the agent-side ResNetBEVBackbone: name, inplanes, planes, stride, H, W
BASIC_CASES = [
    ("basic_identity", 64, 64, 1, 10, 12),
    ("basic_down_s2", 64, 128, 2, 11, 16),         # odd height: (H - 1) // 2 + 1 output rows on both branches
]


def basic_tensors(idx):
    name, inplanes, planes, stride, H, W = BASIC_CASES[idx]
    rng = np.random.default_rng(700 + idx)

    def conv(cout, cin, k, gain=1.0):
        w = rng.normal(0, gain * np.sqrt(2.0 / (cin * k * k)), size=(cout, cin, k, k)).astype(np.float32)
        return w, rng.uniform(-0.2, 0.2, size=cout).astype(np.float32)

    t = {"conv1": conv(planes, inplanes, 3), "conv2": conv(planes, planes, 3, 0.5)}
    if stride != 1 or inplanes != planes:
        t["down"] = conv(planes, inplanes, 1)
    q_in = rng.integers(0, 256, size=(2, inplanes, H, W)).astype(np.uint8)
    q_in[rng.random(q_in.shape) > 0.6] = 0
    return t, q_in


def block_tensors(idx):
    """Float weights / biases of the block's convs (PyTorch layouts) and the input codes, from one numpy stream."""
    name, inplanes, planes, stride, H, W = BLOCK_CASES[idx]
    width = 2 * planes
    rng = np.random.default_rng(500 + idx)

    def conv(cout, cin_g, k):
        w = rng.normal(0, np.sqrt(2.0 / (cin_g * k * k)), size=(cout, cin_g, k, k)).astype(np.float32)
        b = rng.uniform(-0.2, 0.2, size=cout).astype(np.float32)
        return w, b

    t = {"conv1": conv(width, inplanes, 1), "conv2": conv(width, width // GROUPS, 3), "conv3": conv(planes, width, 1)}
    if stride != 1 or inplanes != planes:
        t["down"] = conv(planes, inplanes, 1)
    q_in = rng.integers(0, 256, size=(2, inplanes, H, W)).astype(np.uint8)
    q_in[rng.random(q_in.shape) > 0.6] = 0
    return t, q_in


# ------------------------------------------------------------------------------------------ small pyramid backbone
# HEAL's pyramid backbone scaled down in depth and map size, same widths / strides / grouping
# (hypes_yaml/.../pyramid configs: layer_nums [3, 5, 8], layer_strides [1, 2, 2], num_filters [64, 128, 256]).
PYRAMID_CFG = dict(layer_nums=[2, 2, 1], layer_strides=[1, 2, 2], num_filters=[64, 128, 256],
                   upsample_strides=[1, 2, 4], num_upsample_filter=[128, 128, 128], inplanes=64, resnext=True,
                   stage="collab")
PYRAMID_AGENTS, PYRAMID_H, PYRAMID_W = 3, 24, 32


def pyramid_tensors(seed=900):
    """Float weights of every conv of the ResNeXt pyramid backbone + occupancy heads, and the FP32 input features
    [N, 64, H, W] (what the codebook decoder hands to the backbone: off the quantization grid, any sign)."""
    rng = np.random.default_rng(seed)

    def conv(cout, cin_g, k, gain=1.0):
        w = rng.normal(0, gain * np.sqrt(2.0 / (cin_g * k * k)), size=(cout, cin_g, k, k)).astype(np.float32)
        b = rng.uniform(-0.2, 0.2, size=cout).astype(np.float32)
        return w, b

    t = {}
    inpl = PYRAMID_CFG["inplanes"]
    for li, (nb, stride, planes) in enumerate(zip(PYRAMID_CFG["layer_nums"], PYRAMID_CFG["layer_strides"],
                                                  PYRAMID_CFG["num_filters"])):
        width = 2 * planes
        for bi in range(nb):
            s = stride if bi == 0 else 1
            blk = {"conv1": conv(width, inpl, 1), "conv2": conv(width, width // GROUPS, 3),
                   "conv3": conv(planes, width, 1, 0.5), "stride": s}
            if bi == 0 and (s != 1 or inpl != planes):
                blk["down"] = conv(planes, inpl, 1)
            t[f"l{li}.b{bi}"] = blk
            inpl = planes
        t[f"head{li}"] = conv(1, planes, 1)
    x = (rng.standard_normal((PYRAMID_AGENTS, 64, PYRAMID_H, PYRAMID_W)) * 1.5).astype(np.float32)
    x[rng.random(x.shape) > 0.7] = 0
    # deblocks (ConvTranspose2d, kernel = stride, BN folded): weight [cin, cout, s, s]
    for li, (cin, cout, s) in enumerate(zip(PYRAMID_CFG["num_filters"], PYRAMID_CFG["num_upsample_filter"],
                                            PYRAMID_CFG["upsample_strides"])):
        w = rng.normal(0, np.sqrt(2.0 / cin), size=(cin, cout, s, s)).astype(np.float32)
        t[f"up{li}"] = (w, rng.uniform(-0.2, 0.2, size=cout).astype(np.float32))
    return t, x
