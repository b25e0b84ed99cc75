"""PointPillars front end: the oracle restatement against the reference's golden BEV codes (CPU), and
the CUDA kernel bit-exact against the oracle + through the model-level API from pillars (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import pillar_oracle
from tests.test_golden_cpu import GOLD, build_small_mirror, calibrate_from_pillars


def _setup(fusion="att"):
    from quantv2x_b200.export import pillar_spec

    g = np.load(os.path.join(GOLD, f"e2e_{fusion}.npz"))
    qt = build_small_mirror(fusion)
    data, _ = calibrate_from_pillars(qt, g)
    return g, qt, data, pillar_spec(qt.model.encoder_m1)


def test_oracle_vs_reference_bev_codes():
    """The reference's own BEV codes (golden, produced by the unmodified QuantPointPillar) vs the restatement:
    identical up to isolated 1-LSB flips from the different FP32 summation orders."""
    g, qt, data, spec = _setup()
    assert abs(spec["out_quant"][0] - float(g["bev_delta"])) <= 1e-6 * float(g["bev_delta"])
    bev = pillar_oracle.pillar_bev(spec, g["voxel_features"], g["voxel_coords"], g["voxel_num_points"], 3)
    ref = g["bev_codes"].transpose(0, 2, 3, 1)
    assert bev.shape == ref.shape
    d = np.abs(bev.astype(np.int64) - ref.astype(np.int64))
    assert d.max() <= 1, f"max difference {d.max()} LSB"
    occupied = (ref.max(-1) > 0).sum() * ref.shape[-1]
    assert (d > 0).sum() <= 1e-3 * occupied, f"{(d > 0).sum()} of {occupied} occupied elements differ"
    # the torch mirror (used for calibration) agrees with the reference as well
    with torch.no_grad():
        mirror = qt.model.encoder_m1(data, "m1")
    mc = torch.round(mirror / float(g["bev_delta"])).numpy().astype(np.int64).transpose(0, 2, 3, 1)
    assert np.abs(mc - ref.astype(np.int64)).max() <= 1


def test_oracle_edge_cases():
    """Empty input, a pillar with a single point, a full pillar, duplicate-free scatter at the grid corners."""
    _, _, _, spec = _setup()
    ny, nx = spec["ny"], spec["nx"]
    assert pillar_oracle.pillar_bev(spec, np.zeros((0, 32, 4), np.float32), np.zeros((0, 4), np.int32),
                                    np.zeros((0,), np.int32), 2).max() == 0
    rng = np.random.default_rng(0)
    vf = np.zeros((3, 32, 4), np.float32)
    vf[0, :1] = rng.normal(size=(1, 4))
    vf[1] = rng.normal(size=(32, 4))
    vf[2, :7] = rng.normal(size=(7, 4))
    coords = np.array([[0, 0, 0, 0], [1, 0, ny - 1, nx - 1], [0, 0, ny - 1, 0]], np.int32)
    num = np.array([1, 32, 7], np.int32)
    bev = pillar_oracle.pillar_bev(spec, vf, coords, num, 2)
    assert bev.shape == (2, ny, nx, 64)
    mask = np.zeros((2, ny, nx), bool)
    mask[coords[:, 0], coords[:, 2], coords[:, 3]] = True
    assert bev[~mask].max() == 0


@pytest.mark.gpu
def test_kernel_bit_exact_vs_oracle(cuda_device):
    from quantv2x_b200.export import build_pillar_engine

    g, qt, _, spec = _setup()
    eng = build_pillar_engine(qt.model.encoder_m1)
    rng = np.random.default_rng(1)
    # (a) the golden pillars; (b) a large random set with ragged point counts and grid-corner cells
    cases = [(g["voxel_features"], g["voxel_coords"], g["voxel_num_points"], 3)]
    M, batch = 5000, 4
    num = rng.integers(1, 33, size=M).astype(np.int32)
    vf = (rng.normal(size=(M, 32, 4)) * np.array([8.0, 4.0, 1.0, 0.3])).astype(np.float32)
    vf[np.arange(32)[None, :] >= num[:, None]] = 0
    cells = rng.choice(batch * spec["ny"] * spec["nx"], size=M, replace=False)
    coords = np.stack([cells // (spec["ny"] * spec["nx"]), np.zeros(M, np.int64),
                       (cells // spec["nx"]) % spec["ny"], cells % spec["nx"]], 1).astype(np.int32)
    cases.append((vf, coords, num, batch))
    cases.append((np.zeros((0, 32, 4), np.float32), np.zeros((0, 4), np.int32), np.zeros((0,), np.int32), 1))
    for vf, coords, num, b in cases:
        ref = pillar_oracle.pillar_bev(spec, vf, coords, num, b)
        got = eng.forward(torch.from_numpy(vf).to(cuda_device), torch.from_numpy(coords).to(cuda_device),
                          torch.from_numpy(num).to(cuda_device), b).cpu().numpy()
        assert np.array_equal(got, ref), f"{(got != ref).sum()} BEV bytes differ from the oracle"


@pytest.mark.gpu
def test_model_api_from_pillars(cuda_device):
    """encode_features on the reference-style pillar dict == encode_features on the BEV codes the kernel produced;
    and those BEV codes are within 1 LSB of the reference's golden ones."""
    from quantv2x_b200.export import attach_engines

    g, qt, data, _ = _setup()
    attach_engines(qt, device=cuda_device)          # bev_delta from the calibrated encoder
    model = qt.model
    pipe = model._pipelines["m1"]
    assert getattr(pipe, "pillar_engine", None) is not None
    data_gpu = {"inputs_m1": {k: v.to(cuda_device) for k, v in data["inputs_m1"].items()},
                "agent_modality_list": ["m1"] * 3, "pairwise_t_matrix": data["pairwise_t_matrix"],
                "record_len": torch.tensor([3])}
    bev = pipe.bev_from_inputs(data_gpu)
    ref = torch.from_numpy(np.ascontiguousarray(g["bev_codes"].transpose(0, 2, 3, 1)))
    assert (bev.cpu().to(torch.int64) - ref.to(torch.int64)).abs().max() <= 1
    codes_a, _, info = model.encode_features(data_gpu)
    codes_b, _, _ = model.encode_features({**data_gpu, "inputs_m1": {"bev_u8": bev}})
    for a, b in zip(codes_a, codes_b):
        assert torch.equal(a, b)
    out = model.decode_features(codes_a, info)
    assert out["preds_tensor"].shape == g["preds"].shape


@pytest.mark.gpu
def test_saved_engine_reproduces_the_pipeline(cuda_device, tmp_path):
    """a pipeline saved with serialize.save_pipeline and rebuilt with load_pipeline (no float model,
    no yaml, no calibration) produces bit-identical BEV codes, features, code planes and head maps; the packed wire
    message round-trips the code planes."""
    from quantv2x_b200.collab_model import normalize_pairwise_tfm
    from quantv2x_b200.export import attach_engines
    from quantv2x_b200.serialize import load_pipeline, pack_codes, save_pipeline, unpack_codes

    g, qt, data, _ = _setup()
    attach_engines(qt, device=cuda_device)
    pipe = qt.model._pipelines["m1"]
    path = str(tmp_path / "engine.npz")
    save_pipeline(pipe, path)
    pipe2 = load_pipeline(path, cuda_device)
    assert pipe2.pillar_engine is not None and pipe2.fusion_mode == pipe.fusion_mode
    inp = {k: v.to(cuda_device) for k, v in data["inputs_m1"].items()}
    aff = normalize_pairwise_tfm(data["pairwise_t_matrix"], qt.model.H, qt.model.W, qt.model.fake_voxel_size)[
        0, 0, :3].contiguous().to(cuda_device)
    outs = []
    for p in (pipe, pipe2):
        bev = p.pillar_engine.forward(inp["voxel_features"], inp["voxel_coords"], inp["voxel_num_points"], 3)
        codes = p.encode_agents(bev)
        preds = p.decode_fuse_heads(codes, aff)
        outs.append((bev.clone(), p.encode_buffers(3)["feat"].clone(), codes.clone(), preds.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    k = pipe.codebook.k[0]
    msg = pack_codes(outs[0][2].cpu().numpy(), k)
    back, k2 = unpack_codes(msg)
    assert k2 == k and np.array_equal(back, outs[0][2].cpu().numpy())
    assert len(msg) < outs[0][2].numel()            # fewer bytes than one byte per code


@pytest.mark.gpu
def test_front_end_emits_the_row_sums_the_plan_needs(cuda_device):
    """qv2x_pillar_forward_rs: the per-cell code sums written next to the BEV map equal a scan of the map, and the plan
    fed with them (qv2x_plan_forward_rs) gives the same features as the plan that scans the map itself."""
    import bench
    from quantv2x_b200.export import attach_engines
    from quantv2x_b200.synthetic import synthetic_pillars

    q, bev_delta = bench.build_calibrated_model(cuda_device, "max", 8)
    attach_engines(q, bev_delta=bev_delta, device=cuda_device)
    pipe = q.model._pipelines["m1"]
    enc = q.hypes["model"]["args"]["m1"]["encoder_args"]
    pil = [torch.from_numpy(t).to(cuda_device) for t in synthetic_pillars(3, 2, enc["lidar_range"], enc["voxel_size"], 3000)]
    pe = pipe.pillar_engine
    rs = torch.full((2, pe.ny, pe.nx), -7, dtype=torch.int32, device=cuda_device)
    bev = pe.forward(*pil, 2, rowsum_out=rs)
    assert torch.equal(rs, bev.to(torch.int32).sum(-1).to(torch.int32))
    a = pipe.encode_agents(bev, slot=0).clone()
    f_a = pipe.encode_buffers(2, 0)["feat"].clone()
    b = pipe.encode_pillars(*pil, 2, slot=1)
    assert torch.equal(pipe.encode_buffers(2, 1)["feat"], f_a) and torch.equal(a, b)


@pytest.mark.gpu
def test_keep_zero_protocol_of_the_agent_stage(cuda_device):
    """encode_pillars scatters into an all-zero map and zeroes this frame's cells again after the plan has consumed it
    (qv2x_pillar_scatter / qv2x_pillar_clear): successive DIFFERENT frames through the same slot give the same codes as
    the clearing path, and the stage's buffers are all-zero between frames."""
    import bench
    from quantv2x_b200.export import attach_engines
    from quantv2x_b200.synthetic import synthetic_pillars

    q, bev_delta = bench.build_calibrated_model(cuda_device, "max", 8)
    attach_engines(q, bev_delta=bev_delta, device=cuda_device)
    pipe = q.model._pipelines["m1"]
    enc = q.hypes["model"]["args"]["m1"]["encoder_args"]
    pe = pipe.pillar_engine
    frames = [[torch.from_numpy(t).to(cuda_device) for t in
               synthetic_pillars(seed, 2, enc["lidar_range"], enc["voxel_size"], 2500)] for seed in (11, 12, 13)]
    # out-of-grid rows are dropped by the scatter and must be ignored by the clear as well
    frames[1][1][0, 2] = 10 ** 6
    ext = torch.empty((2, pe.ny, pe.nx, pe.cout), dtype=torch.uint8, device=cuda_device)
    for pil in frames:
        ref = pipe.encode_pillars(*pil, 2, slot=3, bev_out=ext).clone()          # full clear, caller's buffer
        got = pipe.encode_pillars(*pil, 2, slot=3)
        assert torch.equal(ref, got)
        buf = pipe._enc_buf[("pillar", 2, 3)]
        assert int(buf["bev"].max()) == 0 and int(buf["rowsum"].abs().max()) == 0
