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
