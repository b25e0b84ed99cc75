"""Whole quantized backbone + shrinker through the C++ plan vs the integer oracle chained layer by layer
(bit-exact), and the block-level drop-in wrappers vs the torch fake-quant path (tolerance tier)."""
import numpy as np
import pytest
import torch

from oracle import int_oracle

pytestmark = pytest.mark.gpu

BACKBONE_CFG = dict(layer_nums=[1, 2, 2], layer_strides=[2, 2, 2], num_filters=[64, 128, 256],
                    upsample_strides=[1, 2, 4], num_upsample_filter=[128, 128, 128])
SHRINK_CFG = dict(kernal_size=[3], stride=[1], padding=[1], dim=[256], input_dim=384)


def build_calibrated(seed, w_bits, H, W):
    import torch.nn as nn

    from quantv2x_b200.bev_modules import BaseBEVBackbone, DownsampleConv
    from quantv2x_b200.quant import QuantModel, set_act_quantize_params, set_weight_quantize_params
    from quantv2x_b200.synthetic import seeded_init

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone_m1 = BaseBEVBackbone(BACKBONE_CFG, 64)
            self.shrinker_m1 = DownsampleConv(SHRINK_CFG)

        def forward(self, x):
            return self.shrinker_m1(self.backbone_m1(x))

    m = M().eval()
    seeded_init(m, seed)
    q = QuantModel(m, dict(n_bits=w_bits, channel_wise=True, scale_method="minmax"),
                   dict(n_bits=8, channel_wise=False, scale_method="minmax", leaf_param=True, prob=1.0)).eval()
    set_weight_quantize_params(q)
    rng = np.random.default_rng(seed)
    bev_delta = np.float32(0.05)
    xq = rng.integers(0, 256, size=(2, H, W, 64), dtype=np.uint8)
    xq[rng.random((2, H, W, 1)).repeat(64, 3) > 0.3] = 0          # sparse pillars
    x = torch.from_numpy((xq.astype(np.float32) * bev_delta).transpose(0, 3, 1, 2).copy())
    set_act_quantize_params(q, [x])
    return q, xq, x, bev_delta


def oracle_chain(q, xq, bev_delta):
    """Layer-by-layer integer oracle over the calibrated modules."""
    bb, sh = q.model.backbone_m1, q.model.shrinker_m1

    def params(qm):
        w_int, d, z = qm.integer_weight()
        b = None if qm.bias is None else qm.bias.detach().numpy()
        return w_int, d, z, b, float(qm.act_quantizer.delta)

    ups, x, dx = [], xq, bev_delta
    for blk in bb.blocks:
        for j, qm in enumerate(list(blk)[1:]):
            w, d, z, b, od = params(qm)
            _, x = int_oracle.conv_oracle(x, w, d, z, b, dx, od, stride=2 if j == 0 else 1, pad=1)
            dx = od
        ups.append((x, dx))
    cat, cat_d = [], []
    for (u, du), de in zip(ups, bb.deblocks):
        w, d, z, b, od = params(de[0])
        _, y = int_oracle.deconv_oracle(u, w, d, z, b, du, od, stride=de[0].fwd_kwargs["stride"][0])
        cat.append(y)
        cat_d.append(od)
    x = np.concatenate(cat, axis=-1)
    dxs = cat_d
    for dc in sh.layers:
        for qm in dc.double_conv:
            w, d, z, b, od = params(qm)
            _, x = int_oracle.conv_oracle(x, w, d, z, b, dxs, od, stride=1, pad=1)
            dxs = [od]
    return np.concatenate(cat, axis=-1), cat_d, x, dxs[0]


@pytest.mark.parametrize("w_bits", [8, 4])
def test_plan_bit_exact_and_dropin(cuda_device, w_bits):
    from quantv2x_b200.export import build_modality_engines

    q, xq, x, bev_delta = build_calibrated(5, w_bits, 24, 40)
    eng = build_modality_engines(q.model.backbone_m1, q.model.shrinker_m1, float(bev_delta))
    cat_ref, cat_d, out_ref, out_d = oracle_chain(q, xq, bev_delta)
    xd = torch.from_numpy(xq).to(cuda_device)
    cat = eng["backbone"].forward_u8(xd).cpu().numpy()
    assert np.array_equal(cat, cat_ref), "backbone concat output differs from the chained integer oracle"
    out = eng["fused"].forward_u8(xd).cpu().numpy()
    assert out_ref.std() > 3, "degenerate activations"
    assert np.array_equal(out, out_ref), "fused backbone+shrinker output differs from the chained integer oracle"
    assert abs(eng["out_delta"] - out_d) == 0

    # tolerance tier: de-quantized output vs the torch fake-quant path (the reference's arithmetic)
    with torch.no_grad():
        q.model.backbone_m1._engine = None
        bb_engine_ready = q.model.backbone_m1.engine_ready()
        assert bb_engine_ready
        # float path: temporarily mark the blocks as "not ready" by calling their float bodies directly
        feats = q.model.backbone_m1.decode_multiscale_feature(q.model.backbone_m1.get_multiscale_feature(x))
        y = feats
        for layer in q.model.shrinker_m1.layers:
            y = layer(y)
    y_ref = y.numpy().transpose(0, 2, 3, 1)
    # LSB flips of the FP32 path (accumulation order) cascade through the layers, so end to end only the
    # magnitude is bounded (a few LSB; SURVEY 8c tier 2); per-layer parity is exact (test_layer_gpu).
    lsb = np.abs(np.rint(y_ref / out_d) - out.astype(np.float64))
    assert lsb.max() <= 4 and lsb.mean() < 0.3, (lsb.max(), lsb.mean())

    # block-level drop-in: FP32 NCHW in / out through the attached engines
    q.model.backbone_m1.attach_engine(eng["backbone"])
    q.model.shrinker_m1.attach_engine(eng["shrinker"])
    q.cuda()
    with torch.no_grad():
        y_gpu = q(x.to(cuda_device)).cpu().numpy().transpose(0, 2, 3, 1)
    assert np.array_equal(np.rint(y_gpu / np.float32(out_d)).astype(np.uint8), out)


def test_engine_required(cuda_device):
    """Quantized inference without an attached engine must fail loudly, not fall back to torch."""
    q, xq, x, bev_delta = build_calibrated(6, 8, 16, 24)
    with pytest.raises(RuntimeError, match="no libqv2x engine"):
        q(x)
