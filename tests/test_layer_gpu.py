"""GPU parity of the tcgen05 quantized conv / transposed-conv layers against the integer oracle:
int32 accumulators and uint8 outputs must be bit-exact (north_star check #1)."""
import zlib

import numpy as np
import pytest
import torch

from oracle import int_oracle
from tests.layer_cases import make_conv, make_deconv, make_input

pytestmark = pytest.mark.gpu

CONV_CASES = [
    # name, n, H, W, cin, cout, k, stride, pad, w_bits, groups
    ("s0_64_64", 2, 20, 36, 64, 64, 3, 1, 1, 8, 1),
    ("s1_first_64_128_s2", 1, 24, 40, 64, 128, 3, 2, 1, 8, 1),
    ("s1_128_128", 2, 10, 44, 128, 128, 3, 1, 1, 8, 1),
    ("s2_first_128_256_s2", 1, 50, 176, 128, 256, 3, 2, 1, 8, 1),
    ("s2_256_256", 1, 25, 88, 256, 256, 3, 1, 1, 8, 1),
    ("shrink0_cat384_256", 1, 12, 40, 384, 256, 3, 1, 1, 8, 3),
    ("w4_128_128", 1, 16, 32, 128, 128, 3, 1, 1, 4, 1),
    ("w4_cat384_256", 1, 8, 32, 384, 256, 3, 1, 1, 4, 3),
    ("head_1x1_256_64", 1, 9, 33, 256, 64, 1, 1, 0, 8, 1),
    ("odd_size_64_64", 3, 7, 13, 64, 64, 3, 1, 1, 8, 1),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_bit_exact(cuda_device, case):
    from quantv2x_b200.engine import QLayer

    name, n, H, W, cin, cout, k, stride, pad, w_bits, groups = case
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    p = make_conv(rng, cin, cout, k, w_bits, groups)
    x = make_input(rng, n, H, W, cin)
    acc_ref, q_ref = int_oracle.conv_oracle(x, p["w_int"], p["w_delta"], p["w_zp"], p["bias"], p["in_delta"],
                                            p["out_delta"], 0.0, stride=stride, pad=pad, relu=True)
    layer = QLayer(kind=0, w_int=p["w_int"], w_delta=p["w_delta"], w_zp=p["w_zp"], bias=p["bias"], ksize=k,
                   stride=stride, pad=pad, w_bits=w_bits, relu=True, in_delta=p["in_delta"], out_delta=p["out_delta"])
    xd = torch.from_numpy(x).to(cuda_device)
    ho, wo = layer.out_shape(H, W)
    acc = torch.zeros((groups, n * ho * wo, cout), dtype=torch.int32, device=cuda_device)
    rs_out = torch.zeros((n, ho, wo), dtype=torch.int32, device=cuda_device)
    y = layer.forward(xd, acc_dump=acc, rowsum_out=rs_out)
    torch.cuda.synchronize()
    acc = acc.cpu().numpy().reshape(groups, n, ho, wo, cout)
    assert np.array_equal(acc.astype(np.int64), acc_ref), "int32 accumulators differ"
    yq = y.cpu().numpy()
    assert q_ref.std() > 5, "degenerate test vector"
    assert np.array_equal(yq, q_ref), f"uint8 outputs differ at {np.argwhere(yq != q_ref)[:5]}"
    assert np.array_equal(rs_out.cpu().numpy(), int_oracle.rowsum_oracle(q_ref))


DECONV_CASES = [
    ("de0_64_128_s1", 1, 12, 40, 64, 128, 1),
    ("de1_128_128_s2", 2, 10, 24, 128, 128, 2),
    ("de2_256_128_s4", 1, 7, 22, 256, 128, 4),
]


@pytest.mark.parametrize("case", DECONV_CASES, ids=[c[0] for c in DECONV_CASES])
def test_deconv_bit_exact(cuda_device, case):
    from quantv2x_b200.engine import QLayer

    name, n, H, W, cin, cout, s = case
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    p = make_deconv(rng, cin, cout, s)
    x = make_input(rng, n, H, W, cin)
    acc_ref, q_ref = int_oracle.deconv_oracle(x, p["w_int"], p["w_delta"], p["w_zp"], p["bias"], p["in_delta"][0],
                                              p["out_delta"], 0.0, stride=s, relu=True)
    layer = QLayer(kind=1, w_int=p["w_int"], w_delta=p["w_delta"], w_zp=p["w_zp"], bias=p["bias"], ksize=s, stride=s,
                   pad=0, w_bits=8, relu=True, in_delta=p["in_delta"], out_delta=p["out_delta"])
    xd = torch.from_numpy(x).to(cuda_device)
    acc = torch.zeros((3, n * H * W, s * s * cout), dtype=torch.int32, device=cuda_device)
    # write into a channel slice of a wider concat buffer, as the backbone does
    out = torch.zeros((n, H * s, W * s, 384), dtype=torch.uint8, device=cuda_device)
    rs_out = torch.zeros((n, H * s, W * s), dtype=torch.int32, device=cuda_device)
    layer.forward(xd, out=out, out_cbase=128, acc_dump=acc, rowsum_out=rs_out)
    torch.cuda.synchronize()
    assert np.array_equal(acc.cpu().numpy().astype(np.int64), acc_ref), "digit accumulators differ"
    o = out.cpu().numpy()
    assert q_ref.std() > 5
    assert np.array_equal(o[..., 128:256], q_ref)
    assert not o[..., :128].any() and not o[..., 256:].any(), "wrote outside the channel slice"
    assert np.array_equal(rs_out.cpu().numpy(), int_oracle.rowsum_oracle(q_ref))


def test_conv_full_size_crop(cuda_device):
    """Shrinker conv at the real 100x352 map: compare rows 40..59 with the oracle run on the halo crop."""
    from quantv2x_b200.engine import QLayer

    rng = np.random.default_rng(7)
    p = make_conv(rng, 256, 256, 3)
    x = make_input(rng, 1, 100, 352, 256)
    layer = QLayer(kind=0, w_int=p["w_int"], w_delta=p["w_delta"], w_zp=p["w_zp"], bias=p["bias"], ksize=3, stride=1,
                   pad=1, w_bits=8, relu=True, in_delta=p["in_delta"], out_delta=p["out_delta"])
    y = layer.forward(torch.from_numpy(x).to(cuda_device)).cpu().numpy()
    _, q_ref = int_oracle.conv_oracle(x[:, 39:61], p["w_int"], p["w_delta"], p["w_zp"], p["bias"], p["in_delta"],
                                      p["out_delta"], 0.0, stride=1, pad=1, relu=True)
    assert np.array_equal(y[:, 40:60], q_ref[:, 1:21])
    _, q_top = int_oracle.conv_oracle(x[:, :9], p["w_int"], p["w_delta"], p["w_zp"], p["bias"], p["in_delta"],
                                      p["out_delta"], 0.0, stride=1, pad=1, relu=True)
    assert np.array_equal(y[:, :8], q_top[:, :8])


def test_adaround_quantizer_through_the_engine(cuda_device):
    """AdaRound-style weight quantizer + nn.Parameter activation scale -> integer_weight() -> libqv2x layer: bit-exact
    against the integer oracle on the exported grid (which tests/test_golden_cpu.py pins to the reference)."""
    from quantv2x_b200.export import qlayer_from_module
    from tests.adaround_case import build

    qm, g = build()
    layer = qlayer_from_module(qm, [float(g["in_delta"])])
    w_int, w_delta, w_zp = qm.integer_weight()
    xq = np.ascontiguousarray(g["xq"].transpose(0, 2, 3, 1))
    acc_ref, q_ref = int_oracle.conv_oracle(xq, w_int, w_delta, w_zp, g["bias"], g["in_delta"],
                                            float(qm.act_quantizer.delta))
    n, H, W, _ = xq.shape
    acc = torch.zeros((1, n * H * W, q_ref.shape[-1]), dtype=torch.int32, device=cuda_device)
    y = layer.forward(torch.from_numpy(xq).to(cuda_device), acc_dump=acc)
    torch.cuda.synchronize()
    assert np.array_equal(acc.cpu().numpy().reshape(acc_ref.shape).astype(np.int64), acc_ref)
    assert np.array_equal(y.cpu().numpy(), q_ref)
    d = np.abs(q_ref.astype(np.int64) - g["out_codes"].transpose(0, 2, 3, 1).astype(np.int64))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
