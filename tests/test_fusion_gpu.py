"""GPU parity of warp + max/attention fusion, heads GEMM and the layout converters (tolerance tier:
floating-point path, atol = 1e-4 * max|x| against the float64 oracle; converters are bit-exact)."""
import numpy as np
import pytest
import torch

from oracle import fusion_oracle as fo

pytestmark = pytest.mark.gpu


def make_affines(n, rng, H=80.0, W=281.6):
    """Ego row of pairwise poses: identity for the ego, agent j translated (3j, -1.5j) m and yawed 5j degrees."""
    t = np.tile(np.eye(4), (1, n, n, 1, 1))
    for j in range(1, n):
        th = np.deg2rad(5.0 * j)
        t[0, 0, j, :2, :2] = [[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]
        t[0, 0, j, 0, 3] = 3.0 * j
        t[0, 0, j, 1, 3] = -1.5 * j
    return fo.normalize_pairwise_tfm(t, H, W, 1.0)[0, 0, :n]


@pytest.mark.parametrize("mode", ["max", "att"])
@pytest.mark.parametrize("shape", [(1, 12, 20, 64), (3, 25, 44, 256), (8, 10, 16, 256), (2, 9, 7, 128)])
def test_fuse(cuda_device, mode, shape):
    from quantv2x_b200.engine import fuse

    n, H, W, C = shape
    rng = np.random.default_rng(n * 1000 + H)
    feat = (rng.random((n, H, W, C)) * 4 * (rng.random((n, H, W, C)) > 0.4)).astype(np.float32)
    aff = make_affines(n, rng)
    ref = (fo.max_fusion if mode == "max" else fo.att_fusion)(feat, aff)
    out = fuse(torch.from_numpy(feat).to(cuda_device), aff, mode).cpu().numpy()
    np.testing.assert_allclose(out, ref, atol=1e-4 * np.abs(ref).max(), rtol=1e-4)


def test_weighted_fuse_vs_reference_golden(cuda_device):
    """Pyramid-level fusion through the module mirror of weighted_fuse, against the outputs of the
    reference itself (tests/golden/fusion_weighted.npz) at the three level shapes; tolerance 1e-4 * max|x|."""
    import os

    from quantv2x_b200.fusion_modules import weighted_fuse

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fusion_weighted.npz"))
    aff = torch.from_numpy(g["affine"]).to(cuda_device)
    for lvl in range(3):
        feat = torch.from_numpy(g[f"l{lvl}.feat"]).to(cuda_device)
        occ = torch.from_numpy(g[f"l{lvl}.occ"]).to(cuda_device)
        rl = torch.tensor([feat.shape[0]])
        ref = g[f"l{lvl}.fused"]
        out = weighted_fuse(feat, occ, rl, aff, False, score_is_logit=True)[0].cpu().numpy()
        np.testing.assert_allclose(out, ref, atol=1e-4 * np.abs(ref).max(), rtol=1e-4)
        out2 = weighted_fuse(feat, torch.sigmoid(occ) + 1e-4, rl, aff, False)[0].cpu().numpy()
        np.testing.assert_allclose(out2, ref, atol=1e-4 * np.abs(ref).max(), rtol=1e-4)


@pytest.mark.parametrize("shape", [(1, 12, 20, 64), (5, 25, 44, 128), (8, 10, 16, 256), (2, 9, 7, 512)])
def test_weighted_fuse_vs_oracle(cuda_device, shape):
    from quantv2x_b200.engine import fuse_weighted

    n, H, W, C = shape
    rng = np.random.default_rng(n * 77 + H)
    feat = rng.standard_normal((n, H, W, C)).astype(np.float32)
    occ = (rng.standard_normal((n, H, W)) * 3).astype(np.float32)
    aff = make_affines(n, rng)
    ref = fo.weighted_fusion(feat, occ, aff)
    out = fuse_weighted(torch.from_numpy(feat).to(cuda_device), torch.from_numpy(occ).to(cuda_device), aff)
    np.testing.assert_allclose(out.cpu().numpy(), ref, atol=1e-4 * np.abs(ref).max(), rtol=1e-4)
    # ready-made scores with an exact zero region (the camera crop mask case): those pixels drop the agent.
    # (Only for a rotated agent: with the identity warp the sample points sit exactly on pixel centres, where
    # "warped score == 0" depends on the last ulp of the sampling coordinate in any implementation.)
    if n > 1:
        score = (1.0 / (1.0 + np.exp(-occ.astype(np.float64))) + 1e-4).astype(np.float32)
        score[n - 1, : H // 2] = 0.0
        ref = fo.weighted_fusion(feat, score, aff, score_is_logit=False)
        out = fuse_weighted(torch.from_numpy(feat).to(cuda_device), torch.from_numpy(score).to(cuda_device), aff,
                            score_is_logit=False)
        np.testing.assert_allclose(out.cpu().numpy(), ref, atol=1e-4 * np.abs(ref).max(), rtol=1e-4)
    # every agent out of range -> zeros
    far = np.tile(np.array([[1.0, 0.0, 9.0], [0.0, 1.0, 0.0]], np.float32), (n, 1, 1))
    z = fuse_weighted(torch.from_numpy(feat).to(cuda_device), torch.from_numpy(occ).to(cuda_device), far)
    assert not z.any()


@pytest.mark.parametrize("shape", [(1, 12, 20, 64), (5, 25, 44, 128), (8, 10, 16, 256), (2, 9, 8, 512), (3, 7, 12, 72)])
def test_weighted_fuse_u8_equals_dequant_then_fuse(cuda_device, shape):
    """qv2x_fuse_weighted_u8 (the level fusion straight from the uint8 codes) is dequantize_u8 + fuse_weighted bit for
    bit: the codes are de-quantized on load by the same fl(delta * code)."""
    from quantv2x_b200.engine import dequantize_u8, fuse_weighted, fuse_weighted_u8

    n, H, W, C = shape
    rng = np.random.default_rng(n * 31 + C)
    codes = rng.integers(0, 256, size=(n, H, W, C), dtype=np.uint8)
    codes[rng.random(codes.shape) > 0.6] = 0
    occ = (rng.standard_normal((n, H, W)) * 3).astype(np.float32)
    aff = make_affines(n, rng)
    delta = 0.0371
    cd = torch.from_numpy(codes).to(cuda_device)
    od = torch.from_numpy(occ).to(cuda_device)
    a = fuse_weighted_u8(cd, delta, od, aff)
    b = fuse_weighted(dequantize_u8(cd, delta), od, aff)
    assert torch.equal(a, b)
    ref = fo.weighted_fusion(codes.astype(np.float64) * np.float64(np.float32(delta)), occ, aff)
    np.testing.assert_allclose(a.cpu().numpy(), ref, atol=1e-4 * np.abs(ref).max(), rtol=1e-4)


def test_heads(cuda_device):
    from quantv2x_b200.engine import HeadsEngine

    rng = np.random.default_rng(3)
    # cout > 72: wider FP32 GEMMs (the deblocks of the pyramid path) run as 72-column output chunks of one launch
    for pixels, cin, cout in [(1000, 256, 72), (35200, 256, 72), (77, 64, 20), (2200, 256, 2048), (8800, 128, 512),
                              (35200, 64, 128), (300, 64, 100)]:
        x = rng.normal(size=(pixels, cin)).astype(np.float32)
        w = rng.normal(size=(cout, cin)).astype(np.float32) / 16
        b = rng.normal(size=cout).astype(np.float32)
        ref = (x.astype(np.float64) @ w.astype(np.float64).T + b).T
        out = HeadsEngine(w, b).forward(torch.from_numpy(x).to(cuda_device)).cpu().numpy()
        np.testing.assert_allclose(out, ref, atol=1e-5 * np.abs(ref).max(), rtol=1e-5)


def test_converters_bit_exact(cuda_device):
    from quantv2x_b200 import engine

    rng = np.random.default_rng(4)
    x = (rng.random((2, 64, 13, 37)) * 30).astype(np.float32)
    x[rng.random(x.shape) > 0.5] = 0
    delta = np.float32(0.1137)
    q_ref = np.clip(np.rint(x / delta), 0, 255).astype(np.uint8).transpose(0, 2, 3, 1)
    xd = torch.from_numpy(x).to(cuda_device)
    q = engine.quantize_nchw_to_nhwc_u8(xd, float(delta))
    assert np.array_equal(q.cpu().numpy(), q_ref)
    back = engine.dequant_nhwc_u8_to_nchw_f32(q, float(delta)).cpu().numpy()
    assert np.array_equal(back, (q_ref.astype(np.float32) * delta).transpose(0, 3, 1, 2))
    nhwc = engine.nchw_to_nhwc_f32(xd)
    assert np.array_equal(nhwc.cpu().numpy(), x.transpose(0, 2, 3, 1))
    assert np.array_equal(engine.nhwc_to_nchw_f32(nhwc).cpu().numpy(), x)


def test_tiled_ego_stage_equals_whole(cuda_device):
    """decode_regions + fuse_tile + heads per tile, assembled, must equal the whole-frame ego stage bit for bit
    (this is what makes the multi-GPU result identical to the single-GPU one)."""
    from quantv2x_b200 import engine as E
    from quantv2x_b200.distributed import rank_tile, tile_grid
    from quantv2x_b200.pipeline import CollabPipeline
    from tests.codebook_cases import make_codebook_params

    rng = np.random.default_rng(0)
    n, ho, wo, C = 5, 20, 48, 256
    cbs, heads = make_codebook_params(3, C, 1, [128] * 3)
    cb = E.CodebookEngine(cbs, heads)
    hd = E.HeadsEngine(rng.normal(size=(72, C)).astype(np.float32) / 16, rng.normal(size=72).astype(np.float32))

    class _Plan:
        def out_shape(self, h, w):
            return ho, wo, C

    class _Fused:
        plan = _Plan()

    aff = make_affines(n, rng)
    aff[3, 0, 2] += 0.4           # push one agent partly out of view
    aff_d = torch.from_numpy(aff.astype(np.float32)).to(cuda_device)
    codes = torch.from_numpy(rng.integers(0, 128, size=(3, 1, n * ho * wo), dtype=np.uint8)).to(cuda_device)
    for mode in ("att", "max"):
        pipe = CollabPipeline(_Fused(), 0.1, cb, hd, mode, (2 * ho, 2 * wo), cuda_device)
        whole = pipe.decode_fuse_heads_chain(codes, aff_d).clone()
        if mode == "att":
            # the one-kernel ego stage (qv2x_ego_att) is the same function up to fp32 summation order
            assert pipe.ego_att is not None
            folded = pipe.decode_fuse_heads(codes, aff_d)
            tol = 1e-4 * float(whole.abs().max())
            assert float((folded - whole).abs().max()) <= tol
        for world in (2, 4, 8):
            gy, gx = tile_grid(world)
            th, tw = ho // gy, wo // gx
            out = torch.empty((world, 72, th * tw), dtype=torch.float32, device=cuda_device)
            for r in range(world):
                out[r] = pipe.decode_fuse_heads_tile(codes, aff_d, aff, rank_tile(r, world, ho, wo))
            full = out.view(gy, gx, 72, th, tw).permute(2, 0, 3, 1, 4).reshape(72, ho * wo)
            assert torch.equal(full, whole), f"{mode}, {world} tiles: tiled ego stage differs from the whole frame"


EGO_ATT_CASES = [
    # n, ho, wo, C, m, ks, cout, ego matrix perturbed
    (5, 20, 48, 256, 1, [128] * 3, 72, False),
    (8, 25, 44, 256, 1, [64] * 3, 72, False),
    (1, 12, 20, 64, 1, [64, 64], 20, False),
    (2, 9, 7, 128, 2, [64] * 3, 72, False),        # six code planes
    (3, 17, 33, 256, 1, [128] * 3, 72, True),      # general (non-identity) ego matrix: four-tap query
    (4, 100, 352, 256, 1, [128] * 3, 72, False),   # the benchmark's map
]


@pytest.mark.parametrize("case", EGO_ATT_CASES, ids=lambda c: f"n{c[0]}_{c[1]}x{c[2]}_c{c[3]}_m{c[4]}_k{c[5][0]}x{len(c[5])}"
                                                               f"_o{c[6]}{'_egowarp' if c[7] else ''}")
def test_ego_att_one_kernel(cuda_device, case):
    """qv2x_ego_att (codes -> head maps in one kernel over the folded tables) against the float64 restatement of
    decode -> warp -> attention fusion -> heads on the library's own fp32 decode tables, and against the three-kernel
    chain; tolerance tier, atol = 1e-4 * max|y|."""
    from quantv2x_b200 import engine as E
    from tests.codebook_cases import make_codebook_params

    n, ho, wo, C, m, ks, cout, egowarp = case
    rng = np.random.default_rng(n * 131 + ho)
    cbs, hds = make_codebook_params(7, C, m, ks)
    cb = E.CodebookEngine(cbs, hds)
    w = rng.normal(size=(cout, C)).astype(np.float32) / 16
    b = rng.normal(size=cout).astype(np.float32)
    hd = E.HeadsEngine(w, b)
    assert E.EgoAttEngine.supported(cb, hd)
    eng = E.EgoAttEngine(cb, hd)
    aff = make_affines(n, rng).astype(np.float32)
    if n > 2:
        aff[n - 1, 0, 2] += 0.4                       # one agent partly out of view
    if egowarp:
        th = np.deg2rad(2.0)
        aff[0] = np.array([[np.cos(th), -np.sin(th) * ho / wo, 0.01], [np.sin(th) * wo / ho, np.cos(th), -0.02]])
    hw = ho * wo
    nt = len(ks) * m
    kk = [k for k in ks for _ in range(m)]
    codes = np.stack([rng.integers(0, kk[i], size=n * hw, dtype=np.uint8) for i in range(nt)]).reshape(len(ks), m, -1)
    codes_d = torch.from_numpy(codes).to(cuda_device)
    aff_d = torch.from_numpy(aff).to(cuda_device)
    out = eng.forward(codes_d, aff_d, n, ho, wo).cpu().numpy()

    tab = cb.folded(5).astype(np.float64).reshape(-1, C)
    cst = cb.folded(4).astype(np.float64)
    feat = np.tile(cst, (n * hw, 1))
    base = 0
    for i in range(nt):
        feat += tab[base + codes.reshape(nt, -1)[i].astype(np.int64)]
        base += kk[i]
    feat = feat.reshape(n, ho, wo, C)
    ref = fo.heads(fo.att_fusion(feat, aff), w, b).reshape(cout, hw)
    np.testing.assert_allclose(out, ref, atol=1e-4 * np.abs(ref).max(), rtol=1e-4)

    chain = hd.forward(E.fuse(cb.decode(codes_d).view(n, ho, wo, C), aff_d, "att")).cpu().numpy()
    np.testing.assert_allclose(out, chain, atol=1e-4 * np.abs(ref).max(), rtol=1e-4)


def test_ego_att_unsupported_configurations(cuda_device):
    """A head table that does not fit in shared memory (k = 256 x 3 levels) is reported as unsupported and the
    pipeline keeps the three-kernel chain; so does max fusion."""
    from quantv2x_b200 import engine as E
    from quantv2x_b200.pipeline import CollabPipeline
    from tests.codebook_cases import make_codebook_params

    rng = np.random.default_rng(1)
    hd = E.HeadsEngine(rng.normal(size=(72, 256)).astype(np.float32) / 16, None)
    big = E.CodebookEngine(*make_codebook_params(3, 256, 1, [256] * 3))
    assert not E.EgoAttEngine.supported(big, hd)
    with pytest.raises(Exception, match="unsupported configuration"):
        E.EgoAttEngine(big, hd)
    small = E.CodebookEngine(*make_codebook_params(3, 256, 1, [128] * 3))

    class _Plan:
        def out_shape(self, h, w):
            return 10, 16, 256

    class _Fused:
        plan = _Plan()

    assert CollabPipeline(_Fused(), 0.1, big, hd, "att", (20, 32), cuda_device).ego_att is None
    assert CollabPipeline(_Fused(), 0.1, small, hd, "max", (20, 32), cuda_device).ego_att is None
    assert CollabPipeline(_Fused(), 0.1, small, hd, "att", (20, 32), cuda_device).ego_att is not None


def test_push_planes_and_heads_tile_addressing(cuda_device):
    """The two kernels of the peer-memory exchange, exercised on one GPU (the 'peers' are local buffers):
    qv2x_push_planes must place every plane at its agent-major offset in every destination, and
    qv2x_heads_forward_tile must store a tile's head maps at their place in the full map, bit-identical to the
    compact result."""
    from quantv2x_b200 import engine as E

    rng = np.random.default_rng(5)
    planes, rows_local, world, rank = 6, 35200, 4, 2
    codes = torch.from_numpy(rng.integers(0, 256, size=(3, 2, rows_local), dtype=np.uint8)).to(cuda_device)
    dsts = [torch.zeros((planes, world * rows_local), dtype=torch.uint8, device=cuda_device) for _ in range(3)]
    E.push_planes(codes, world * rows_local, rank * rows_local, [d.data_ptr() for d in dsts])
    torch.cuda.synchronize()
    for d in dsts:
        got = d.cpu().numpy()
        assert np.array_equal(got[:, rank * rows_local:(rank + 1) * rows_local], codes.cpu().numpy().reshape(planes, -1))
        assert got[:, :rank * rows_local].max() == 0 and got[:, (rank + 1) * rows_local:].max() == 0

    # all-to-all form: destination p receives the rows of frame p at this rank's agent offset
    per_rows, nfr = 704, 4
    batch = torch.from_numpy(rng.integers(0, 256, size=(3, 2, nfr * per_rows), dtype=np.uint8)).to(cuda_device)
    dsts = [torch.zeros((planes, nfr * per_rows), dtype=torch.uint8, device=cuda_device) for _ in range(nfr)]
    E.scatter_planes(batch, nfr * per_rows, rank * per_rows, [d.data_ptr() for d in dsts])
    torch.cuda.synchronize()
    b = batch.cpu().numpy().reshape(planes, nfr, per_rows)
    for f, d in enumerate(dsts):
        got = d.cpu().numpy()
        assert np.array_equal(got[:, rank * per_rows:(rank + 1) * per_rows], b[:, f])
        assert got[:, :rank * per_rows].max() == 0 and got[:, (rank + 1) * per_rows:].max() == 0

    ho, wo, C = 20, 48, 256
    hd = E.HeadsEngine(rng.normal(size=(72, C)).astype(np.float32) / 16, rng.normal(size=72).astype(np.float32))
    y0, y1, x0, x1 = 10, 20, 12, 24
    x = torch.from_numpy(rng.normal(size=((y1 - y0) * (x1 - x0), C)).astype(np.float32)).to(cuda_device)
    compact = hd.forward(x)                                            # [72, tile_pixels]
    full = torch.full((72, ho * wo), -7.0, dtype=torch.float32, device=cuda_device)
    E.heads_forward_tile(hd, x, full.data_ptr() + 4 * (y0 * wo + x0), x1 - x0, wo, ho * wo)
    torch.cuda.synchronize()
    f = full.view(72, ho, wo)
    assert torch.equal(f[:, y0:y1, x0:x1].reshape(72, -1), compact)
    mask = torch.ones((ho, wo), dtype=torch.bool, device=cuda_device)
    mask[y0:y1, x0:x1] = False
    assert bool((f[:, mask] == -7.0).all()), "stores outside the tile"
