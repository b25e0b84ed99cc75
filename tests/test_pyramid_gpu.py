"""GPU parity of the pyramid-fusion building blocks: ResNeXt bottleneck blocks (grouped 3x3 as a
block-diagonal int8 GEMM, shortcut + block quantizer fused into the last conv's epilogue, FP32 downsample conv),
occupancy head and the per-level score-weighted fusion.  Integer outputs are bit-exact against oracle/int_oracle.py,
which tests/test_golden_cpu.py pins to the reference's QuantBottleneck."""
import os

import numpy as np
import pytest
import torch

from oracle import fusion_oracle as fo
from oracle import int_oracle
from tests.pyramid_cases import BLOCK_CASES, IN_DELTA
from tests.test_golden_cpu import GOLD, block_params

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("idx", [0, 1], ids=[c[0] for c in BLOCK_CASES])
def test_bottleneck_block_bit_exact(cuda_device, idx):
    from quantv2x_b200.pyramid import BottleneckEngine

    g = np.load(os.path.join(GOLD, "pyramid_blocks.npz"))
    name, p, x = block_params(g, idx)
    ref = int_oracle.bottleneck_oracle(x, IN_DELTA, p)
    eng = BottleneckEngine(p, IN_DELTA)
    taps = {}
    xd = torch.from_numpy(np.ascontiguousarray(x)).to(cuda_device)
    out, rs = eng.forward(xd, want_rowsum=True, taps=taps)
    assert np.array_equal(taps["q1"].cpu().numpy(), ref["q1"])
    assert np.array_equal(taps["q2"].cpu().numpy(), ref["q2"])
    if "down" in p:
        assert np.array_equal(taps["res"].cpu().numpy(), ref["res"])          # FP32 output: same fma, bit-exact
    assert np.array_equal(out.cpu().numpy(), ref["out"])
    assert np.array_equal(rs.cpu().numpy(), ref["out"].astype(np.int64).sum(-1))
    # and against the reference block itself: <= 1 LSB on < 1 % of the codes (free-running over three convs)
    d = np.abs(out.cpu().numpy().astype(np.int64) - g[f"{name}.out.codes"].transpose(0, 2, 3, 1).astype(np.int64))
    assert d.max() <= 1 and (d > 0).mean() < 1e-2


def test_bottleneck_chain_and_larger_map(cuda_device):
    """Two chained blocks (down then identity) on a map that spans several tiles, W4 weights on the grouped conv."""
    from quantv2x_b200.pyramid import BottleneckEngine

    rng = np.random.default_rng(11)
    n, H, W = 2, 50, 88

    def qconv(cout, cin_g, k, bits=8):
        w = rng.normal(0, np.sqrt(2.0 / (cin_g * k * k)), size=(cout, cin_g, k, k)).astype(np.float32)
        d, z = int_oracle.weight_qparams_minmax(w, bits)
        return dict(w_int=int_oracle.weight_int_grid(w, d, z, bits), w_delta=d, w_zp=z, w_bits=bits,
                    bias=rng.uniform(-0.2, 0.2, size=cout).astype(np.float32))

    def block(inpl, planes, stride, down):
        width = 2 * planes
        p = dict(stride=stride, groups=32, out_delta=np.float32(0.05), conv1=qconv(width, inpl, 1),
                 conv2=qconv(width, width // 32, 3, 4), conv3=qconv(planes, width, 1))
        p["conv1"]["act_delta"], p["conv2"]["act_delta"] = np.float32(0.03), np.float32(0.04)
        if down:
            p["down"] = qconv(planes, inpl, 1)
        return p

    pa, pb = block(64, 128, 2, True), block(128, 128, 1, False)
    x = rng.integers(0, 256, size=(n, H, W, 64)).astype(np.uint8)
    x[rng.random(x.shape) > 0.5] = 0
    ra = int_oracle.bottleneck_oracle(x, IN_DELTA, pa)
    rb = int_oracle.bottleneck_oracle(ra["out"], pa["out_delta"], pb)
    ea, eb = BottleneckEngine(pa, IN_DELTA), BottleneckEngine(pb, float(pa["out_delta"]))
    ya, rsa = ea.forward(torch.from_numpy(x).to(cuda_device), want_rowsum=True)
    yb = eb.forward(ya, rowsum=rsa)
    assert ra["out"].std() > 3 and rb["out"].std() > 3
    assert np.array_equal(ya.cpu().numpy(), ra["out"])
    assert np.array_equal(yb.cpu().numpy(), rb["out"])


def test_occupancy_head_and_level_fusion(cuda_device):
    """single_head_i (1x1 conv to one channel, FP32 logits) bit-exact against the oracle's fma, then one level of
    forward_collab from codes: dequantize -> warp features and scores -> softmax over agents."""
    from quantv2x_b200.pyramid import OccupancyHead, weighted_fuse_level
    from tests.test_fusion_gpu import make_affines

    rng = np.random.default_rng(5)
    n, H, W, C = 3, 20, 36, 128
    delta = np.float32(0.021)
    codes = rng.integers(0, 256, size=(n, H, W, C)).astype(np.uint8)
    codes[rng.random(codes.shape) > 0.5] = 0
    w = rng.normal(0, 0.1, size=(1, C, 1, 1)).astype(np.float32)
    wd, wz = int_oracle.weight_qparams_minmax(w, 8)
    w_int = int_oracle.weight_int_grid(w, wd, wz, 8)
    bias = np.array([-0.3], np.float32)
    _, occ_ref = int_oracle.conv_oracle(codes, w_int, wd, wz, bias, delta, None, stride=1, pad=0, relu=False)
    head = OccupancyHead(w_int, wd, wz, bias, float(delta))
    cd = torch.from_numpy(codes).to(cuda_device)
    occ = head.forward(cd)
    assert np.array_equal(occ.cpu().numpy(), occ_ref[..., 0])
    aff = make_affines(n, rng)
    fused = weighted_fuse_level(cd, float(delta), occ, aff).cpu().numpy()
    ref = fo.weighted_fusion(codes.astype(np.float32) * delta, occ_ref[..., 0], aff)
    np.testing.assert_allclose(fused, ref, atol=1e-4 * np.abs(ref).max(), rtol=1e-4)


def test_pyramid_backbone_collab(cuda_device):
    """QuantPyramidFusion.forward_collab up to the fused level features on the small seeded backbone: the first
    conv (FP32 GEMM on the decoded features) within 1 LSB of the oracle on < 0.1 % of the codes; with those codes
    teacher-forced, every level's codes and occupancy logits bit-exact and the fused maps within 1e-4; free-running,
    within the drift of the reference fixture (tests/golden/pyramid_backbone.npz)."""
    from oracle import pyramid_oracle
    from quantv2x_b200.pyramid import PyramidBackboneEngine
    from tests.pyramid_cases import PYRAMID_AGENTS, PYRAMID_CFG
    from tests.test_golden_cpu import pyramid_params

    g = np.load(os.path.join(GOLD, "pyramid_backbone.npz"))
    P, x = pyramid_params(g)
    aff = g["affine"][0, 0, :PYRAMID_AGENTS]
    nums = PYRAMID_CFG["layer_nums"]
    eng = PyramidBackboneEngine(P, nums)
    xd = torch.from_numpy(x).to(cuda_device)
    taps = {}
    fused = eng.forward_collab(xd, aff, taps=taps)
    q1_gpu = taps["q1_first"].cpu().numpy()
    levels, q1_ref = pyramid_oracle.backbone_collab(x, P, aff, nums, q1_override=q1_gpu)
    d = np.abs(q1_gpu.astype(np.int64) - q1_ref.astype(np.int64))
    assert q1_ref.std() > 3 and d.max() <= 1 and (d > 0).mean() < 1e-3, (d.max(), (d > 0).mean())
    for li, lv in enumerate(levels):
        assert np.array_equal(taps[f"l{li}.codes"].cpu().numpy(), lv["codes"]), li
        assert np.array_equal(taps[f"l{li}.occ"].cpu().numpy(), lv["occ"]), li
        np.testing.assert_allclose(fused[li].cpu().numpy(), lv["fused"], atol=1e-4 * np.abs(lv["fused"]).max(),
                                   rtol=1e-4)
        # against the reference itself (free-running): same bounds as the CPU oracle test
        ref = g[f"l{li}.codes"].transpose(0, 2, 3, 1)
        dr = np.abs(taps[f"l{li}.codes"].cpu().numpy().astype(np.int64) - ref.astype(np.int64))
        assert dr.max() <= 2 and (dr > 0).mean() < 3e-2 and (dr > 1).mean() < 1e-3
        fr = g[f"l{li}.fused"].transpose(1, 2, 0)
        np.testing.assert_allclose(fused[li].cpu().numpy(), fr, atol=2e-2 * np.abs(fr).max())
    # deblocks + concat (decode_multiscale_feature): the oracle teacher-forced with the GPU's fused maps -- FP32 GEMM
    # in the documented order -> within 1 LSB on < 0.1 % of the codes; and within the drift bounds of the reference
    cat = eng.decode_multiscale_feature(fused)[0].cpu().numpy()
    cat_ref, deltas = pyramid_oracle.decode_multiscale([f.cpu().numpy() for f in fused], P)
    assert cat.shape == cat_ref.shape and [float(v) for v in deltas] == eng.up_deltas
    dc = np.abs(cat.astype(np.int64) - cat_ref.astype(np.int64))
    assert cat_ref.std() > 3 and dc.max() <= 1 and (dc > 0).mean() < 1e-3, (dc.max(), (dc > 0).mean())
    gold = np.concatenate([g[f"up{li}.codes"] for li in range(3)]).transpose(1, 2, 0)
    dg = np.abs(cat.astype(np.int64) - gold.astype(np.int64))
    assert dg.max() <= 3 and (dg > 0).mean() < 5e-2 and (dg > 1).mean() < 2e-3


def test_quant_pyramid_fusion_module_on_engine(cuda_device):
    """The drop-in module: QuantPyramidFusion calibrated in torch, exported, engine attached -> forward_collab returns
    the reference's tensors (NCHW FP32 [1, 384, H, W] on the deblocks' grids + occupancy maps) within the free-running
    drift bounds of the fake-quant float body."""
    from quantv2x_b200.pyramid import PyramidBackboneEngine
    from tests.pyramid_cases import PYRAMID_AGENTS, pyramid_tensors
    from tests.test_golden_cpu import build_pyramid_mirror

    q, g, final, occ = build_pyramid_mirror()
    q.attach_engine(PyramidBackboneEngine(q.export_params(), q.layer_nums()))
    _, x = pyramid_tensors()
    out, occ_g = q.forward_collab(torch.from_numpy(x).to(cuda_device), torch.tensor([PYRAMID_AGENTS]),
                                  torch.from_numpy(g["affine"]).to(cuda_device))
    assert tuple(out.shape) == tuple(final.shape) and out.dtype == torch.float32
    for li in range(3):
        d = float(q.deblocks[li][0].act_quantizer.delta)
        a = torch.round(out[0, 128 * li:128 * (li + 1)].cpu() / d).numpy().astype(np.int64)
        b = torch.round(final[0, 128 * li:128 * (li + 1)] / d).numpy().astype(np.int64)
        dd = np.abs(a - b)
        assert dd.max() <= 3 and (dd > 0).mean() < 5e-2 and (dd > 1).mean() < 2e-3, (li, dd.max(), (dd > 0).mean())
        assert tuple(occ_g[li].shape) == tuple(occ[li].shape)
        # the head's own quantizer is active (as in the reference): both maps lie on its grid, whole steps apart at most
        hd = float(getattr(q, f"single_head_{li}").act_quantizer.delta)
        steps = np.abs(occ_g[li].cpu().numpy() - occ[li].numpy()) / hd
        assert steps.max() <= 2.001 and (steps > 0.5).mean() < 0.12, (li, steps.max(), (steps > 0.5).mean())


@pytest.mark.parametrize("idx", [0, 1], ids=["basic_identity", "basic_down_s2"])
def test_basic_block_bit_exact(cuda_device, idx):
    """QuantBasicBlock on the engine (two 3x3 int8 convs, shortcut in the second conv's epilogue, FP32 strided 1x1
    downsample): every tensor bit-exact against the integer oracle, and within 1 LSB on < 1 % of the reference block."""
    from quantv2x_b200.pyramid import BasicBlockEngine
    from tests.test_golden_cpu import basic_params

    g = np.load(os.path.join(GOLD, "basic_blocks.npz"))
    name, p, x = basic_params(g, idx)
    ref = int_oracle.basicblock_oracle(x, IN_DELTA, p)
    eng = BasicBlockEngine(p, IN_DELTA)
    taps = {}
    out, rs = eng.forward(torch.from_numpy(np.ascontiguousarray(x)).to(cuda_device), want_rowsum=True, taps=taps)
    assert np.array_equal(taps["q1"].cpu().numpy(), ref["q1"])
    if "down" in p:
        assert np.array_equal(taps["res"].cpu().numpy(), ref["res"])
    assert np.array_equal(out.cpu().numpy(), ref["out"])
    assert np.array_equal(rs.cpu().numpy(), ref["out"].astype(np.int64).sum(-1))
    d = np.abs(out.cpu().numpy().astype(np.int64) - g[f"{name}.out.codes"].transpose(0, 2, 3, 1).astype(np.int64))
    assert d.max() <= 1 and (d > 0).mean() < 1e-2


@pytest.mark.parametrize("case", [(64, 128, 1, 20, 44), (128, 128, 2, 25, 44), (256, 128, 4, 13, 22), (64, 16, 2, 5, 7)],
                         ids=lambda c: f"cin{c[0]}_c{c[1]}_s{c[2]}_{c[3]}x{c[4]}")
def test_deblock_f32_quantizing_epilogue(cuda_device, case, monkeypatch):
    """DeblockF32 with the quantizer + pixel shuffle in the GEMM's epilogue (qv2x_heads_forward_deconv_u8) writes the
    same codes, bit for bit, as GEMM -> planar FP32 -> quantizing converter -> permuted copy, and leaves the other
    channels of the concat buffer untouched."""
    from quantv2x_b200.pyramid import DeblockF32

    cin, cout, s, h, w = case
    rng = np.random.default_rng(cin + s)
    up = dict(stride=s, act_delta=0.043, w_int=rng.integers(0, 256, size=(cin, cout, s, s)).astype(np.uint8),
              w_zp=rng.integers(100, 156, size=cin).astype(np.float32),
              w_delta=(rng.random(cin) * 0.004 + 0.001).astype(np.float32), bias=rng.normal(size=cout).astype(np.float32))
    d = DeblockF32(up)
    x = torch.from_numpy((rng.standard_normal((h, w, cin)) * 1.5).astype(np.float32)).to(cuda_device)
    ctot, cbase = 2 * cout + 8, 8
    a = torch.full((h * s, w * s, ctot), 7, dtype=torch.uint8, device=cuda_device)
    b = a.clone()
    d.forward(x, a, cbase)
    monkeypatch.setenv("QV2X_DEBLOCK_CHAIN", "1")
    d.forward(x, b, cbase)
    assert torch.equal(a, b)
    assert int(a[..., cbase:cbase + cout].max()) > 7 and bool((a[..., :cbase] == 7).all())
