"""CPU tests of the oracle itself (no GPU): internal consistency of the restatements."""
import numpy as np
import pytest

from oracle import codebook_oracle as co
from oracle import int_oracle
from tests.codebook_cases import make_codebook_params, make_features, oracle_params
from tests.layer_cases import make_conv, make_deconv, make_input


def test_fold_matches_sequential_fp64():
    cbs, heads = make_codebook_params(3, 64, 2, [32, 32, 32])
    p = oracle_params(cbs, heads)
    q = make_features(0, 600, 64)
    delta = np.float32(0.21)
    x = q.astype(np.float64) * np.float64(delta)
    c64 = co.encode_fp64(p, x)
    cfx = co.encode_fixed_point(co.fold_encode(p), q, delta)
    for a, b in zip(c64, cfx):
        assert (a != b).mean() < 2e-3
    d_seq = co.decode_fp64(p, c64)
    d_tab = co.decode_tables(co.fold_decode(p), c64)
    np.testing.assert_allclose(d_tab, d_seq, atol=1e-12)


def test_split_digits_roundtrip():
    rng = np.random.default_rng(0)
    m = rng.integers(-8355711, 8355712, size=10000)
    hi, mid, lo = int_oracle.split_digits(m)
    assert np.array_equal(hi * 65536 + mid * 256 + lo, m)
    assert lo.min() >= -128 and lo.max() <= 127 and mid.min() >= -128 and mid.max() <= 127


def test_conv_oracle_matches_fakequant_float():
    """Integer oracle vs the reference-style FP32 fake-quant evaluation: <= 1 LSB on a tiny fraction."""
    rng = np.random.default_rng(11)
    p = make_conv(rng, 64, 64, 3)
    x = make_input(rng, 1, 10, 14, 64)
    _, q = int_oracle.conv_oracle(x, p["w_int"], p["w_delta"], p["w_zp"], p["bias"], p["in_delta"], p["out_delta"])
    w_fake = (p["w_int"].astype(np.float32) - p["w_zp"].reshape(-1, 1, 1, 1)) * p["w_delta"].reshape(-1, 1, 1, 1)
    x_hat = (x.astype(np.float32) * p["in_delta"][0]).transpose(0, 3, 1, 2)
    y = int_oracle.fakequant_layer_float(x_hat, w_fake, p["w_delta"], p["w_zp"], p["bias"], p["out_delta"], 0.0)
    qf = np.rint(y / p["out_delta"]).astype(np.int64).transpose(0, 2, 3, 1)
    diff = np.abs(qf - q.astype(np.int64))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3


def test_deconv_oracle_matches_fakequant_float():
    rng = np.random.default_rng(12)
    p = make_deconv(rng, 64, 128, 2)
    x = make_input(rng, 1, 6, 10, 64)
    _, q = int_oracle.deconv_oracle(x, p["w_int"], p["w_delta"], p["w_zp"], p["bias"], p["in_delta"][0],
                                    p["out_delta"], stride=2)
    w_fake = (p["w_int"].astype(np.float32) - p["w_zp"].reshape(-1, 1, 1, 1)) * p["w_delta"].reshape(-1, 1, 1, 1)
    x_hat = (x.astype(np.float32) * p["in_delta"][0]).transpose(0, 3, 1, 2)
    y = int_oracle.fakequant_layer_float(x_hat, w_fake, p["w_delta"], p["w_zp"], p["bias"], p["out_delta"], 0.0,
                                         kind=1, stride=2)
    qf = np.rint(y / p["out_delta"]).astype(np.int64).transpose(0, 2, 3, 1)
    diff = np.abs(qf - q.astype(np.int64))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3


def test_grouped_conv_equals_block_diagonal_dense():
    """The identity behind qv2x_layer_desc.groups: a grouped conv is the dense conv of the block-diagonal weight whose
    off-group entries are the channel's zero-point (real weight 0) -- same int accumulators, same codes."""
    rng = np.random.default_rng(8)
    groups, cin, cout = 8, 64, 64
    w = rng.normal(0, 0.1, size=(cout, cin // groups, 3, 3)).astype(np.float32)
    d, z = int_oracle.weight_qparams_minmax(w, 8)
    wi = int_oracle.weight_int_grid(w, d, z, 8)
    dense = np.empty((cout, cin, 3, 3), np.uint8)
    dense[:] = z.astype(np.uint8).reshape(-1, 1, 1, 1)
    cg, og = cin // groups, cout // groups
    for co in range(cout):
        g0 = (co // og) * cg
        dense[co, g0:g0 + cg] = wi[co]
    x = rng.integers(0, 256, size=(1, 9, 11, cin)).astype(np.uint8)
    b = rng.uniform(-0.1, 0.1, size=cout).astype(np.float32)
    for stride in (1, 2):
        a1, q1 = int_oracle.conv_oracle(x, wi, d, z, b, 0.05, 0.07, stride=stride, pad=1, groups=groups)
        a2, q2 = int_oracle.conv_oracle(x, dense, d, z, b, 0.05, 0.07, stride=stride, pad=1)
        assert np.array_equal(a1, a2) and np.array_equal(q1, q2) and q1.std() > 3


def test_encode_filter_error_bound_holds():
    """The fp32 filter of the codebook encoder (csrc/codebook.cu, chunk_fast / step_end) accepts a row without the
    float64 pass when second - eps(second) > best + eps(best) with the per-score bound
    eps(s) = 2^-20 * (|s| + 2 (max|g0| + sum max|B|) + 2^30 max|d|); the constant covers the conversion error of the
    digit parts, which survives when they cancel in vf.  Restated here in numpy fp32 (same operations, same order) and
    checked against the float64 score of EVERY column on random and adversarial accumulators: digit cancellation
    (65536 hi ~ -(256 mid + lo)), huge and tiny scales, both digit-recombination variants."""
    rng = np.random.default_rng(2024)
    n, cols = 4000, 64
    f32, fma = np.float32, int_oracle.fma32
    for pack16 in (True, False):
        for case in range(4):
            hi = rng.integers(-2 ** 23, 2 ** 23, size=(n, cols))
            mid = rng.integers(-2 ** 22, 2 ** 22, size=(n, cols))
            lo = rng.integers(-2 ** 23, 2 ** 23, size=(n, cols))
            if case == 1:                        # the high digit cancels the low part almost exactly
                hi = rng.integers(-2 ** 14, 2 ** 14, size=(n, cols))
                t12 = -hi * 65536 + rng.integers(-300, 300, size=(n, cols))
                mid, lo = t12 >> 8, t12 & 255
            elif case == 2:                      # small accumulators, constants dominate
                hi, mid, lo = hi >> 20, mid >> 14, lo >> 10
            d = rng.normal(size=cols) * 10.0 ** rng.uniform(-12, -4)
            g0 = rng.normal(size=cols) * 10.0 ** rng.uniform(-3, 3)
            B = [rng.normal(size=(n, cols)) * 10.0 ** rng.uniform(-3, 3) for _ in range(2 if case != 3 else 0)]
            V = hi * 65536 + mid * 256 + lo                                     # exact (int64)
            exact = V.astype(np.float64) * d + g0
            for b in B:
                exact = exact + b
            hif = hi.astype(f32)
            if pack16:
                lof = (mid * 256 + lo).astype(f32)
            else:
                mf, l2 = mid.astype(f32), lo.astype(f32)
                lof = fma(mf, f32(256), l2)
            vf = fma(hif, f32(65536), lof)
            d32, g32 = d.astype(f32), g0.astype(f32)
            sc = fma(vf, d32, g32)
            for b in B:
                sc = sc + b.astype(f32)
            cabs = np.nextafter(f32((np.abs(g0).max() + sum(np.abs(b).max() for b in B)) * (1 + 1e-6)), f32(np.inf))
            k1 = fma(f32(2.0 ** 30), np.abs(d32).max(), f32(2) * cabs)                # 2 C + 2^30 max|d|
            eps = f32(2.0 ** -20) * (np.abs(sc) + k1).astype(f32)                     # per score
            err = np.abs(sc.astype(np.float64) - exact)
            assert (err <= eps.astype(np.float64)).all(), (pack16, case, float((err / eps).max()))
            assert (err / eps).max() < 0.7            # the margin the comment in the kernel claims (10/16)


def test_saturating_requant_fast_path_is_exact():
    """The 8-bit / ReLU epilogue (csrc/epilogue_requant.cuh, FAST8) replaces q = clip(rint(y / delta), 0, 255) by a
    multiply with fl(1/delta), a clamp, the 1.5 * 2^23 magic-number rounding and -- only for elements within 1e-4 of a
    rounding boundary -- the true division.  Restated in numpy fp32 and compared with the normative form over
    log-uniform scales, the whole output range and points placed right at the half-integer boundaries."""
    rng = np.random.default_rng(77)
    f32 = np.float32
    magic = f32(12582912.0)
    n_fallback = n_total = 0
    for _ in range(40):
        delta = f32(10.0 ** rng.uniform(-4, 1))
        rdelta = f32(1.0) / delta
        k = rng.integers(-40, 300, size=200000)
        frac = np.where(rng.random(k.size) < 0.5, 0.5 + rng.normal(0, 3e-5, k.size), rng.random(k.size))
        y = ((k + frac) * np.float64(delta)).astype(f32)
        y[:1000] = (rng.integers(0, 256, 1000) + 0.5).astype(f32) * delta          # as close to a tie as fp32 gets
        want = np.clip(np.rint(np.maximum(y, f32(0)) / delta), 0, 255).astype(np.int64)
        t = np.minimum(np.maximum(y * rdelta, f32(-0.25)), f32(255.25)).astype(f32)
        sft = (t + magic).astype(f32)
        r = (sft - magic).astype(f32)
        near = np.abs((t - r).astype(f32)) > f32(0.4999)
        fast = (sft.view(np.uint32) & 0xFF).astype(np.int64)
        exact = np.clip(np.rint(y / delta), 0, 255).astype(np.int64)
        got = np.where(near, exact, fast)
        assert np.array_equal(got, want), (float(delta), int((got != want).sum()))
        n_fallback += int(near.sum())
        n_total += y.size
    assert 0 < n_fallback < 0.6 * n_total          # both branches were exercised


def test_library_exports_all_symbols():
    """The C-ABI library loads without a GPU and exports every symbol include/qv2x.h declares."""
    import ctypes

    from quantv2x_b200 import _lib

    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = _lib.exported_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(handle, n), f"libqv2x.so does not export {n}"
    assert _lib.lib().qv2x_version() >= 100


def test_c_abi_error_convention_without_gpu():
    """Argument validation happens before any CUDA call: a bad descriptor returns QV2X_ERR_INVALID (-1), leaves the
    output handle untouched and sets a message -- the ctypes shim turns that into an exception (the reference's own
    error behaviour is a Python exception).  No compute call is made."""
    import ctypes
    from ctypes import byref, c_void_p

    from quantv2x_b200 import _lib

    L = _lib.lib()
    # pillar front end: only the 10 -> 64 / 32-point configuration exists
    d = _lib.PillarDesc()
    d.n_feat, d.cout, d.max_points, d.nx, d.ny = 9, 64, 32, 704, 200
    d.out_delta, d.out_bits, d.pre_bits, d.pre_delta = 0.05, 8, 8, 0.1
    w = (ctypes.c_float * 640)()
    h = c_void_p()
    rc = L.qv2x_pillar_create(byref(d), w, None, byref(h))
    assert rc == -1 and not h.value
    assert b"10 decorated features" in L.qv2x_last_error()
    with pytest.raises(_lib.Qv2xError):
        _lib.check(rc)
    # quantized layer: unsupported kernel size
    ld = _lib.LayerDesc()
    ld.kind, ld.cin, ld.cout, ld.ksize, ld.stride, ld.pad = 0, 64, 64, 5, 1, 2
    ld.w_bits, ld.relu, ld.n_in_groups, ld.out_delta, ld.out_bits = 8, 1, 1, 0.1, 8
    buf = (ctypes.c_uint8 * 16)()
    fl = (ctypes.c_float * 64)()
    rc = L.qv2x_layer_create(byref(ld), buf, fl, fl, fl, byref(h))
    assert rc == -1 and b"ksize" in L.qv2x_last_error()
    # grouped weights: channels must divide, conv only
    ld.ksize, ld.pad, ld.groups, ld.cin, ld.cout = 3, 1, 32, 80, 64
    rc = L.qv2x_layer_create(byref(ld), buf, fl, fl, fl, byref(h))
    assert rc == -1 and b"grouped" in L.qv2x_last_error()
    ld.kind, ld.ksize, ld.stride, ld.pad, ld.cin, ld.cout = 1, 2, 2, 0, 64, 64
    assert L.qv2x_layer_create(byref(ld), buf, fl, fl, fl, byref(h)) == -1 and b"grouped" in L.qv2x_last_error()
    # null arguments
    assert L.qv2x_fuse(1, 2, 4, 4, 256, None, None, None, None) == -1
    assert L.qv2x_push_planes(None, 3, 16, 16, 0, None, 1, None) == -1
    assert L.qv2x_scatter_planes(None, 3, 16, 16, 0, None, 1, None) == -1
    assert L.qv2x_fuse_weighted(2, 4, 4, 64, None, None, 1, None, None, None) == -1
    assert b"score" in L.qv2x_last_error()
    assert L.qv2x_dequant_u8(None, 16, 0.1, None, None) == -1
    assert L.qv2x_layer_forward_ex(None, 1, 4, 4, None, 64, 0, None, None, 64, 0, None, None, None, None) == -1


def test_weighted_fuse_module_host_logic():
    """The mirror of weighted_fuse refuses what the kernel does not implement instead of falling back."""
    import torch

    from quantv2x_b200.fusion_modules import weighted_fuse

    x, sc = torch.zeros(2, 8, 4, 4), torch.ones(2, 1, 4, 4)
    aff = torch.zeros(1, 2, 2, 2, 3)
    with pytest.raises(NotImplementedError):
        weighted_fuse(x, sc, torch.tensor([2]), aff, align_corners=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        weighted_fuse(x, sc, torch.tensor([2]), aff, False)


def test_wire_format_roundtrip():
    """pack_codes / unpack_codes: exact round trip for every codebook size, the packed size is
    ceil(log2 k) bits per code, ragged row counts and corrupt messages are handled."""
    from quantv2x_b200.serialize import pack_codes, unpack_codes

    rng = np.random.default_rng(0)
    for k, rows, m in [(128, 35200, 1), (64, 1001, 2), (256, 77, 1), (16, 5, 4), (128, 0, 1)]:
        codes = rng.integers(0, k, size=(3, m, rows), dtype=np.uint8)
        msg = pack_codes(codes, k)
        bits = int(np.ceil(np.log2(k)))
        assert len(msg) == 24 + (3 * m * rows * bits + 7) // 8
        back, k2 = unpack_codes(msg)
        assert k2 == k and back.dtype == np.uint8 and np.array_equal(back, codes)
    assert len(pack_codes(np.zeros((3, 1, 35200), np.uint8), 128)) == 24 + 92400        # 92.4 KB per agent
    with pytest.raises(ValueError):
        pack_codes(np.full((1, 1, 4), 200, np.uint8), 128)
    with pytest.raises(ValueError):
        unpack_codes(b"XXXX" + bytes(40))
    with pytest.raises(ValueError):
        unpack_codes(pack_codes(np.zeros((3, 1, 100), np.uint8), 128)[:-3])


def test_ego_fold_equals_chain():
    """The folded form of the attention-fusion ego stage (oracle/ego_fold_oracle.py = what qv2x_ego_att evaluates:
    Gram matrix of the decode tables for the scores, head table for the outputs) equals the operator-by-operator
    restatement decode -> warp -> attention fusion -> heads, including an agent partly out of view, a rotated ego
    matrix and two segments per level."""
    from oracle import codebook_oracle as co
    from oracle import ego_fold_oracle as ef
    from oracle import fusion_oracle as fo
    from tests.codebook_cases import make_codebook_params, oracle_params

    for n, H, W, C, m, ks, cout, egowarp in [(3, 9, 14, 64, 1, [32, 32, 32], 20, False),
                                             (4, 8, 11, 64, 2, [16, 16], 12, True)]:
        rng = np.random.default_rng(n * 7 + H)
        fd = co.fold_decode(oracle_params(*make_codebook_params(3, C, m, ks)))
        tables = [fd["tables"][l][s] for l in range(len(ks)) for s in range(m)]
        kk = [k for k in ks for _ in range(m)]
        hw = H * W
        codes = np.stack([rng.integers(0, kk[i], size=n * hw) for i in range(len(kk))])
        t = np.tile(np.eye(4), (1, n, n, 1, 1))
        for j in range(1, n):
            th = np.deg2rad(7.0 * j)
            t[0, 0, j, :2, :2] = [[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]
            t[0, 0, j, 0, 3], t[0, 0, j, 1, 3] = 2.0 * j, -1.0 * j
        aff = fo.normalize_pairwise_tfm(t, 8.0, 12.0, 1.0)[0, 0, :n]
        aff[n - 1, 0, 2] += 0.5
        if egowarp:
            aff[0] = np.array([[0.99, -0.05, 0.02], [0.06, 0.98, -0.03]])
        w = rng.normal(size=(cout, C)) / 8
        b = rng.normal(size=cout)
        feat = np.tile(fd["const"], (n * hw, 1))
        for i in range(len(kk)):
            feat += tables[i][codes[i]]
        ref = fo.heads(fo.att_fusion(feat.reshape(n, H, W, C), aff), w, b).reshape(cout, hw)
        got = ef.ego_att_folded(codes, tables, fd["const"], aff, w, b, H, W)
        np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())


def test_decode_linear_fold_equals_chain():
    """oracle/ego_fold_oracle.decode_linear_folded (what qv2x_decode_linear evaluates: folded table rows, then the
    activation quantizer) equals decode -> 1x1 conv -> ReLU -> quantizer evaluated operator by operator."""
    from oracle import codebook_oracle as co
    from oracle import ego_fold_oracle as ef
    from tests.codebook_cases import make_codebook_params, oracle_params

    rng = np.random.default_rng(5)
    C, m, ks, cout, rows = 64, 1, [32, 32, 32], 24, 500
    fd = co.fold_decode(oracle_params(*make_codebook_params(9, C, m, ks)))
    tables = [fd["tables"][l][s] for l in range(len(ks)) for s in range(m)]
    codes = np.stack([rng.integers(0, k, size=rows) for k in ks])
    w = rng.normal(size=(cout, C)) * 0.5
    b = rng.normal(size=cout) * 0.1
    feat = co.decode_tables(fd, [codes[l][:, None] for l in range(len(ks))])
    y = feat @ w.T + b
    delta = np.abs(y).max() / 200.0
    q_ref = np.clip(np.rint(np.maximum(y, 0.0) / delta), 0, 255)
    q, rs, yy = ef.decode_linear_folded(codes, tables, fd["const"], w, b, delta)
    np.testing.assert_allclose(yy, y, rtol=1e-10, atol=1e-10)
    t = y / delta
    near = np.abs(np.abs(t - np.floor(t)) - 0.5) < 1e-6
    assert np.array_equal(q[~near], q_ref[~near].astype(np.uint8)) and np.array_equal(rs, q.astype(np.int64).sum(1))
