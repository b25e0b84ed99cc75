"""Boundary checks that need no GPU: the documented reference-side binding matches the C ABI, descriptors carry their
size, the compressor keeps the reference's call shape, stale engines and unsupported configs fail loudly."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _integration_stub():
    """The python block of INTEGRATION.md section B, up to (not including) make_layer(): what a maintainer copies."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n# opencood/quant/qv2x_binding.py\n(.*?)```", text, re.S).group(1)
    head = block.split("def _ck(rc):")[0]
    head = head.replace('_L = ctypes.CDLL("libqv2x.so")', "").replace("_L.qv2x_last_error.restype = ctypes.c_char_p", "")
    ns = {}
    exec(head, ns)
    return ns


def test_layer_desc_of_the_integration_doc_matches_the_library():
    from quantv2x_b200 import _lib

    ns = _integration_stub()
    doc, ours = ns["LayerDesc"], _lib.LayerDesc
    assert [f[0] for f in doc._fields_] == [f[0] for f in ours._fields_]
    assert ctypes.sizeof(doc) == ctypes.sizeof(ours)
    for name, _ in doc._fields_:
        assert getattr(doc, name).offset == getattr(ours, name).offset, name
    assert [f[0] for f in ns["PlanStep"]._fields_] == [f[0] for f in _lib.PlanStep._fields_]
    # the header declares the same fields in the same order
    hdr = open(os.path.join(ROOT, "include", "qv2x.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} qv2x_layer_desc;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [re.sub(r"\[.*\]", "", n.strip()) for decl in body.split(";") if decl.strip()
             for n in decl.strip().split(None, 1)[1].split(",")]
    assert names == [f[0] for f in ours._fields_]


def test_descriptor_with_wrong_struct_size_is_rejected():
    from quantv2x_b200 import _lib

    lib = _lib.lib()
    d = _lib.LayerDesc()
    assert d.struct_size == ctypes.sizeof(_lib.LayerDesc)
    d.kind, d.cin, d.cout, d.ksize, d.stride, d.pad, d.w_bits, d.relu, d.n_in_groups = 0, 64, 64, 3, 1, 1, 8, 1, 1
    d.out_delta, d.out_bits = 0.1, 8
    d.struct_size -= 4                      # a caller built against a header without the last field
    w = np.zeros((64, 64, 3, 3), np.uint8)
    f = np.ones(64, np.float32)
    h = ctypes.c_void_p()
    rc = lib.qv2x_layer_create(ctypes.byref(d), w.ctypes.data_as(ctypes.c_void_p), f.ctypes.data_as(ctypes.c_void_p),
                               f.ctypes.data_as(ctypes.c_void_p), None, ctypes.byref(h))
    assert rc != 0 and b"struct_size" in lib.qv2x_last_error()
    for cls in (_lib.CodebookDesc, _lib.PillarDesc, _lib.PostprocessDesc, _lib.LayerExtra):
        assert cls().struct_size == ctypes.sizeof(cls)


def test_codebook_grid_scale_is_recovered_from_on_grid_features():
    from quantv2x_b200.codebook import UMGMQuantizer

    rng = np.random.default_rng(3)
    for delta in (0.0371, 0.25, 1.7e-3):
        q = rng.integers(0, 256, size=(500, 64)).astype(np.float32)
        q[rng.random(q.shape) < 0.5] = 0
        q[0, 0] = 6.0            # smallest positive code is not 1: the scale must still come out right
        q[q == 1] = 0
        q[q == 2] = 0
        q[q == 3] = 0
        x = torch.from_numpy(q * np.float32(delta))
        d = UMGMQuantizer._recover_grid_scale(x)
        codes = torch.round(x / d)
        # any scale that reproduces the values on an integer grid <= 255 is a valid answer; the encoder only needs q
        assert float((codes * d - x).abs().max()) <= 1e-4 * delta and float(codes.max()) <= 255
        assert abs(d / delta - round(d / delta)) < 1e-3
    with pytest.raises(ValueError):
        UMGMQuantizer._recover_grid_scale(torch.rand(100, 8))          # not on a grid
    with pytest.raises(ValueError):
        UMGMQuantizer._recover_grid_scale(torch.tensor([[-1.0, 2.0]]))


def test_post_fusion_shrink_header_is_refused():
    from quantv2x_b200 import yaml_utils
    from quantv2x_b200.export import attach_engines
    from quantv2x_b200.quant import QuantModel

    hypes = yaml_utils.load_yaml(yaml_utils.default_config("att"))
    args = hypes["model"]["args"]
    args["lidar_range"] = [-12.8, -6.4, -3, 12.8, 6.4, 1]
    args["m1"]["encoder_args"]["lidar_range"] = [-12.8, -6.4, -3, 12.8, 6.4, 1]
    args["shrink_header"] = dict(kernal_size=[3], stride=[1], padding=[1], dim=[256], input_dim=256)
    model = yaml_utils.create_model(hypes).eval()
    assert model.shrink_flag
    q = QuantModel(model, dict(n_bits=8, channel_wise=True, scale_method="minmax"),
                   dict(n_bits=8, channel_wise=False, scale_method="minmax", leaf_param=True, prob=1.0))
    with pytest.raises(NotImplementedError, match="shrink_header"):
        attach_engines(q, bev_delta=0.05, device=torch.device("cpu"))


def test_stale_engine_is_dropped_when_parameters_change():
    from tests.test_plan_gpu import build_calibrated

    q, _, x, _ = build_calibrated(11, 8, 16, 24)
    bb = q.model.backbone_m1

    class Dummy:
        def forward_nchw(self, t):
            return t

    bb.attach_engine(Dummy())
    bb._check_engine()                                        # fresh: fine
    qm = bb.quant_modules()[3]
    with torch.no_grad():
        qm.act_quantizer.delta.mul_(2.0)                      # re-calibration in place
    with pytest.raises(RuntimeError, match="changed after the libqv2x engine was built"):
        bb._check_engine()
    with pytest.raises(RuntimeError, match="no libqv2x engine"):   # ... and the stale engine is gone
        bb._check_engine()
    bb.attach_engine(Dummy())
    qm.weight_quantizer.n_bits = 4                            # bitwidth_refactor
    with pytest.raises(RuntimeError, match="changed after"):
        bb._check_engine()


def test_wire_header_is_validated():
    from quantv2x_b200.serialize import pack_codes, unpack_codes

    rng = np.random.default_rng(0)
    codes = rng.integers(0, 128, size=(3, 1, 1000), dtype=np.uint8)
    msg = pack_codes(codes, 128)
    out, k = unpack_codes(msg)
    assert k == 128 and np.array_equal(out, codes)
    with pytest.raises(ValueError):
        pack_codes(codes, 512)                                # 9 bits do not fit a byte code
    bad = bytearray(msg)
    bad[6] = 3                                                # bits field no longer ceil(log2 k)
    with pytest.raises(ValueError, match="inconsistent header"):
        unpack_codes(bytes(bad))
    msg64 = pack_codes(np.minimum(codes, 63), 64)
    forged = bytearray(msg64)
    forged[24] = 0xff                                         # first code becomes 63 -> still valid ...
    unpack_codes(bytes(forged))
    # a message claiming k = 100 (7 bits) but carrying code 127
    m100 = bytearray(pack_codes(np.minimum(codes, 99), 100))
    m100[24] = 0xfe
    with pytest.raises(ValueError, match="out of range"):
        unpack_codes(bytes(m100))


def test_quant_pfn_layer_slices_like_the_reference():
    """Above `part` rows the Linear runs slice by slice (reference quant_block.py:611-618): an un-initialised output
    quantizer updates its scale on every call (running min / max), so the calibrated scale depends on the slicing."""
    import torch
    from quantv2x_b200.pillar_modules import PFNLayer
    from quantv2x_b200.quant.quant_block import QuantPFNLayer

    torch.manual_seed(0)
    pfn = PFNLayer(10, 64, use_norm=False, last_layer=True)
    pfn.part = 8
    wq = dict(n_bits=8, channel_wise=True, scale_method="minmax")
    aq = dict(n_bits=8, channel_wise=False, scale_method="minmax", leaf_param=True, prob=1.0)
    q = QuantPFNLayer(pfn, wq, aq)
    x = torch.randn(20, 32, 10)
    x[:16] *= 50.0                                  # the last slice (rows 16..19) has a much smaller range
    q.set_quant_state(True, True)
    q.linear.weight_quantizer.set_inited(False)
    q.linear.act_quantizer.set_inited(False)
    q.act_quantizer.set_inited(False)
    out = q(x)
    assert out.shape == (20, 1, 64)
    d_last = float(q.linear.act_quantizer.delta)
    q2 = QuantPFNLayer(pfn, wq, aq)
    q2.set_quant_state(True, True)
    for z in (q2.linear.weight_quantizer, q2.linear.act_quantizer, q2.act_quantizer):
        z.set_inited(False)
    for lo in (0, 8, 16):
        q2.linear(x[lo:lo + 8])
    assert abs(float(q2.linear.act_quantizer.delta) - d_last) <= 1e-7 * d_last
    q3 = QuantPFNLayer(pfn, wq, aq)
    q3.part = 10 ** 6
    q3.set_quant_state(True, True)
    for z in (q3.linear.weight_quantizer, q3.linear.act_quantizer, q3.act_quantizer):
        z.set_inited(False)
    q3(x)
    d_whole = float(q3.linear.act_quantizer.delta)                      # one call over the whole input
    assert abs(d_whole - d_last) > 1e-3 * d_whole


def test_yaml_anchor_config_matches_the_heads():
    """The post-processor reads postprocess.anchor_args.anchor_generator_config (reference
    voxel_postprocessor_3heads.py:28-40): one entry per class, consistent with the model's head widths."""
    import glob
    import os

    from quantv2x_b200 import yaml_utils
    from quantv2x_b200.postprocess import anchor_config_from_hypes

    here = os.path.dirname(os.path.abspath(yaml_utils.__file__))
    files = sorted(glob.glob(os.path.join(here, "hypes_yaml/v2x_real/Codebook/*/*.yaml")))
    assert len(files) == 3
    for fn in files:
        hy = yaml_utils.load_yaml(fn)
        cfg = anchor_config_from_hypes(hy["postprocess"])
        args = hy["model"]["args"]
        assert len(cfg) == args["num_class"] == 3, fn
        assert all(len(c["anchor_rotations"]) == args["anchor_number"] for c in cfg), fn
        assert {c["feature_map_stride"] for c in cfg} == {2}
