"""quantv2x_b200: B200-native (sm_100a) fast path for QuantV2X's fully quantized intermediate-fusion
inference.  CUDA kernels + C ABI live in csrc/ (built into libqv2x.so); the Python modules mirror the
reference's opencood.quant / codebook / fusion interfaces for this one path."""

__version__ = "0.1.0"
