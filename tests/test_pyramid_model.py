"""The pyramid model driver (reference heter_pyramid_collab_codebook_mc_encdec.py:33-208) against the
fixture generated from the reference itself (oracle/gen_golden_pyramid_e2e.py): the mirror's torch body reproduces
the reference's calibration and outputs on CPU; the libqv2x path reproduces its feature codes, codebook codes and
head maps on the GPU up to the activation-rounding flips of the integer path (rates printed, thresholds explicit)."""
import numpy as np
import pytest
import torch

from tests.pyramid_model_case import build_and_calibrate, frame_dict, load_fixture


def test_mirror_calibration_reproduces_reference():
    g = load_fixture()
    qt, out = build_and_calibrate(g)
    qs = {name: m.act_quantizer for name, m in qt.model.named_modules() if hasattr(m, "act_quantizer")}
    ref_d = dict(zip(g["quantizer_names"].tolist(), g["quantizer_deltas"].tolist()))
    ref_z = dict(zip(g["quantizer_names"].tolist(), g["quantizer_zero_points"].tolist()))
    assert set(ref_d) <= set(qs) and len(ref_d) == 64
    for k, v in ref_d.items():
        assert abs(float(qs[k].delta) - v) <= 1e-6 * abs(v), k
        assert float(qs[k].zero_point) == ref_z[k], k
    feat_codes = torch.round(out["agent_feature"] / float(g["feat_delta"])).numpy().astype(np.int64)
    assert np.array_equal(feat_codes, g["feat_codes"].astype(np.int64))
    codes = np.stack([c.numpy() for c in out["codes"]])
    assert np.array_equal(codes, g["codes"].astype(np.int64))          # encode_float == reference encode
    np.testing.assert_allclose(out["preds_tensor"].numpy(), g["preds"], atol=1e-4 * np.abs(g["preds"]).max())
    for i, o in enumerate(out["occ_single_list"]):
        np.testing.assert_allclose(o.numpy(), g[f"occ{i}"], atol=1e-5 * max(1.0, np.abs(g[f"occ{i}"]).max()))


def test_driver_fails_loudly_without_engines():
    g = load_fixture()
    qt, _ = build_and_calibrate(g)
    with pytest.raises(RuntimeError, match="no libqv2x engines attached"):
        qt.model.encode_features(frame_dict(g))
    with pytest.raises(RuntimeError, match="no libqv2x engines attached"):
        qt.model.decode_features([torch.zeros((1536, 1), dtype=torch.long)] * 3,
                                 {"record_len": torch.tensor([3]), "feature_shape": (3, 64, 16, 32),
                                  "affine_matrix": torch.zeros(1, 5, 5, 2, 3)})


def test_yaml_creates_shipped_configuration():
    from quantv2x_b200 import yaml_utils
    import os

    here = os.path.dirname(os.path.abspath(yaml_utils.__file__))
    hy = yaml_utils.load_yaml(os.path.join(here, "hypes_yaml/v2x_real/Codebook/Pyramid/lidar_pyramid_stage3.yaml"))
    m = yaml_utils.create_model(hy)
    assert [n for n, _ in m.named_children()] == ["encoder_m1", "backbone_m1", "aligner_m1", "pyramid_backbone",
                                                   "shrink_conv", "cls_head", "reg_head", "dir_head", "codebook"]
    assert m.codebook._channel == 64 and m.codebook._m == 1 and m.codebook._k == [128] * 3
    assert len(m.backbone_m1.resnet.layer0) == 3 and m.backbone_m1.resnet.layer0[0].downsample is not None
    with pytest.raises(NotImplementedError):
        hy["model"]["args"]["m1"]["aligner_args"]["core_method"] = "convnext"
        yaml_utils.create_model(hy)


@pytest.mark.gpu
def test_pyramid_driver_vs_reference_fixture(cuda_device):
    from quantv2x_b200.pyramid_model import attach_pyramid_engines

    g = load_fixture()
    qt, _ = build_and_calibrate(g)
    model = attach_pyramid_engines(qt, device=cuda_device)
    data = frame_dict(g, cuda_device)
    eng = model._engines

    # ---- agent side: pillars -> BEV codes (exact) -> BasicBlock chain -> feature codes -> codebook codes
    pe = eng["pillar"]
    n = 3
    bev = pe.forward(data["inputs_m1"]["voxel_features"], data["inputs_m1"]["voxel_coords"],
                     data["inputs_m1"]["voxel_num_points"], n)
    assert np.array_equal(bev.cpu().numpy(), g["bev_codes"].transpose(0, 2, 3, 1))
    feat = eng["backbone"].forward_u8(bev).cpu().numpy().astype(np.int64)
    ref_feat = g["feat_codes"].transpose(0, 2, 3, 1).astype(np.int64)
    d = np.abs(feat - ref_feat)
    print(f"agent feature codes: exact {np.mean(d == 0):.5f}, max |diff| {d.max()}")
    assert d.max() <= 2 and np.mean(d == 0) > 0.98

    codes, mods, info = model.encode_features(data)
    assert mods == ["m1"] * n and info["feature_shape"] == (n, 64, 16, 32)
    codes_np = np.stack([c.cpu().numpy() for c in codes])
    agree = float((codes_np == g["codes"]).all(axis=(0, 2)).mean())
    print(f"codebook codes: rows equal to the reference's in every level {agree:.4f}")
    assert codes_np.shape == g["codes"].shape and agree > 0.90

    # ---- ego side from the REFERENCE's codes: isolates decode -> pyramid -> shrink -> heads
    ref_codes = [torch.from_numpy(g["codes"][l].astype(np.int64)).to(cuda_device) for l in range(3)]
    taps = {}
    out = model.decode_features(ref_codes, info, taps=taps)
    torch.cuda.synchronize()
    pyr = eng["pyramid"]
    # (a) the integer restatement (oracle/pyramid_oracle.py), teacher-forced with the GPU's first-conv codes, at the
    #     shipped depth [3, 5, 8]: every level's codes and occupancy logits bit-exact
    from oracle import pyramid_oracle

    P = qt.model.pyramid_backbone.export_params()
    aff = info["affine_matrix"][0][0, :n].numpy()
    levels, q1_ref = pyramid_oracle.backbone_collab(taps["decoded"].cpu().numpy(), P, aff, [3, 5, 8],
                                                    q1_override=taps["q1_first"].cpu().numpy())
    d1 = np.abs(taps["q1_first"].cpu().numpy().astype(np.int64) - q1_ref.astype(np.int64))
    assert d1.max() <= 1 and (d1 > 0).mean() < 1e-3
    for li, lv in enumerate(levels):
        assert np.array_equal(taps[f"l{li}.codes"].cpu().numpy(), lv["codes"]), li
        assert np.array_equal(taps[f"l{li}.occ"].cpu().numpy(), lv["occ"]), li
    # (b) against the reference's fake-quant FP32 body: rounding flips compound with depth (level 0 after 3
    #     bottlenecks, level 2 after 16); the rates are printed and bounded
    cat = taps["cat"].cpu().numpy().astype(np.int64)[0]
    base = 0
    for li, dblk in enumerate(pyr.deblocks):
        ref_c = np.rint(g["cat"][0, base:base + dblk.cout].astype(np.float64).transpose(1, 2, 0) / dblk.delta)
        dc = np.abs(cat[..., base:base + dblk.cout] - ref_c.astype(np.int64))
        print(f"deblock {li} codes (after {sum(len(s) for s in pyr.stages[:li + 1])} bottlenecks + fusion): exact "
              f"{np.mean(dc == 0):.4f}, within 1 {np.mean(dc <= 1):.4f}, max {dc.max()}, nonzero {np.mean(ref_c > 0):.3f}")
        base += dblk.cout
        assert np.mean(dc <= 1) > (0.999, 0.995, 0.95)[li] and np.mean(dc == 0) > (0.999, 0.85, 0.70)[li]
    sh = taps["shrink"].cpu().numpy().astype(np.int64)[0]
    ref_sh = g["shrink_codes"][0].transpose(1, 2, 0).astype(np.int64)
    ds = np.abs(sh - ref_sh)
    print(f"shrink conv codes: exact {np.mean(ds == 0):.4f}, within 1 {np.mean(ds <= 1):.4f}, max {ds.max()}")
    preds = out["preds_tensor"].cpu().numpy()
    scale = np.abs(g["preds"]).max()
    err = np.abs(preds - g["preds"])
    corr = np.corrcoef(preds.ravel(), g["preds"].ravel())[0, 1]
    print(f"head maps: max err {err.max() / scale:.4f} of range, mean {err.mean() / scale:.5f}, corr {corr:.6f}")
    assert preds.shape == g["preds"].shape
    assert np.mean(ds <= 1) > 0.93 and np.mean(ds == 0) > 0.68 and corr > 0.9995 and err.mean() / scale < 5e-3
    for i, o in enumerate(out["occ_single_list"]):
        ro = g[f"occ{i}"]
        assert tuple(o.shape) == ro.shape
        eo = np.abs(o.cpu().numpy() - ro)
        print(f"occupancy level {i}: mean err {eo.mean():.4f}, max {eo.max():.4f} (range {np.abs(ro).max():.2f})")

    # ---- the whole frame through the public call equals encode -> decode
    out2 = model(data)
    out3 = model.decode_features(codes, info)
    assert torch.equal(out2["preds_tensor"], out3["preds_tensor"])
    assert out2["cls_preds"].shape[1] == 18 and out2["reg_preds"].shape[1] == 42 and out2["dir_preds"].shape[1] == 12

    # ---- module-boundary drop-ins run on the library too
    x = torch.from_numpy(g["bev_codes"].astype(np.float32) * float(g["bev_delta"])).to(cuda_device)
    y = qt.model.backbone_m1(x)
    assert torch.equal(torch.round(y / eng["backbone"].out_delta).cpu().long().permute(0, 2, 3, 1),
                       torch.from_numpy(feat))
