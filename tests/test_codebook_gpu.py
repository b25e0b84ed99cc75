"""GPU parity of the codebook encode / decode kernels (north_star check #1: indices bit-exact, ties -> lowest)."""
import zlib

import numpy as np
import pytest
import torch

from oracle import codebook_oracle as co
from tests.codebook_cases import make_codebook_params, make_features, oracle_params

pytestmark = pytest.mark.gpu

CASES = [
    # name, C, m, ks, rows
    ("c256_m1_k128", 256, 1, [128] * 3, 5000),
    ("c256_m1_k64", 256, 1, [64] * 3, 3000),
    ("c256_m1_k256", 256, 1, [256] * 3, 3000),
    ("c256_m2_k256", 256, 2, [256] * 3, 2000),
    ("c64_m2_k64", 64, 2, [64] * 3, 4001),
    ("c256_m1_k128_ragged", 256, 1, [128] * 3, 127),
]


def library_fold(eng, fe):
    """Rebuild the oracle's folded structure from the tables the LIBRARY holds (kernel parameters)."""
    L, m, C = eng.levels, eng.m, eng.channel
    digits = eng.folded(0).astype(np.int64)
    sc, g0, bt = eng.folded(1), eng.folded(2), eng.folded(3)
    fixed, G0, B = [], [], []
    row = col = boff = 0
    for l in range(L):
        nl = m * eng.k[l]
        fixed.append((digits[row * C:(row + 3 * nl) * C].reshape(3, nl, C), sc[col:col + nl]))
        G0.append(g0[col:col + nl])
        Bl = []
        for j in range(l):
            cnt = m * eng.k[j] * nl
            Bl.append(bt[boff:boff + cnt].reshape(m, eng.k[j], nl))
            boff += cnt
        B.append(Bl)
        row += 3 * nl
        col += nl
    lib_fe = dict(fe)
    lib_fe["g0"], lib_fe["B"] = G0, B
    return lib_fe, fixed


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_encode_decode(cuda_device, case):
    from quantv2x_b200.engine import CodebookEngine

    name, C, m, ks, rows = case
    cbs, heads = make_codebook_params(zlib.crc32(name.encode()) % 1000, C, m, ks)
    p = oracle_params(cbs, heads)
    q = make_features(1, rows, C)
    delta = np.float32(0.173)
    eng = CodebookEngine(cbs, heads)

    # (a) the library's fold agrees with the oracle's independent float64 fold
    fe = co.fold_encode(p)
    lib_fe, fixed = library_fold(eng, fe)
    o_fixed = co.quantize_fold(fe)
    for l in range(len(ks)):
        np.testing.assert_allclose(lib_fe["g0"][l], fe["g0"][l], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(fixed[l][1], o_fixed[l][1], rtol=1e-12)
        recon_lib = (fixed[l][0][0] * 65536 + fixed[l][0][1] * 256 + fixed[l][0][2]) * fixed[l][1][:, None]
        np.testing.assert_allclose(recon_lib, fe["G"][l], atol=float(np.abs(fe["G"][l]).max()) * 1e-7)
        for j in range(l):
            np.testing.assert_allclose(lib_fe["B"][l][j], fe["B"][l][j], rtol=1e-9, atol=1e-12)

    # (b) kernel == exact restatement of its arithmetic on the library's own tables: bit-exact codes
    codes = eng.encode(torch.from_numpy(q).to(cuda_device), float(delta))
    torch.cuda.synchronize()
    codes_np = codes.cpu().numpy()                       # [L, m, rows]
    ref_fx = co.encode_fixed_point(lib_fe, q, delta, fixed)
    for l in range(len(ks)):
        assert np.array_equal(codes_np[l].T, ref_fx[l]), f"level {l}: kernel codes differ from fixed-point oracle"

    # (c) kernel vs canonical float64 oracle: equal except provable near-ties
    x = q.astype(np.float64) * np.float64(delta)
    c64, gaps = co.encode_fp64(p, x, return_gaps=True)
    ok_rows = np.ones(rows, bool)
    for l in range(len(ks)):
        diff = (codes_np[l].T != c64[l]).any(axis=1)
        # a row may differ only where the canonical top-2 gap is below the 24-bit fixed-point resolution,
        # or where an earlier level already differed
        assert (gaps[l].min(axis=1)[diff & ok_rows] < 1e-6).all()
        ok_rows &= ~diff
    assert ok_rows.mean() > 0.999

    # (d) decode: bit-exact vs fp32 table sum in kernel order, tolerance vs sequential float64 decode
    out = eng.decode(codes).cpu().numpy()
    const, tabs = eng.folded(4), eng.folded(5)
    exp = np.broadcast_to(const, (rows, C)).astype(np.float32).copy()
    off = 0
    for l in range(len(ks)):
        for s in range(m):
            t = tabs[off:off + ks[l] * C].reshape(ks[l], C)
            exp = exp + t[codes_np[l, s]]
            off += ks[l] * C
    assert np.array_equal(out, exp)
    d64 = co.decode_fp64(p, [codes_np[l].T.astype(np.int64) for l in range(len(ks))])
    np.testing.assert_allclose(out, d64, atol=1e-5 * np.abs(d64).max(), rtol=0)


def test_argmin_tie_lowest_index(cuda_device):
    """Duplicate codewords produce exact ties: the kernel must return the lowest index."""
    from quantv2x_b200.engine import CodebookEngine

    C, m, ks = 256, 1, [128] * 3
    cbs, heads = make_codebook_params(5, C, m, ks)
    for l in range(3):
        cbs[l][0, 64:] = cbs[l][0, :64]          # codeword k+64 == codeword k
    eng = CodebookEngine(cbs, heads)
    q = make_features(2, 1024, C)
    codes = eng.encode(torch.from_numpy(q).to(cuda_device), 0.2).cpu().numpy()
    assert codes.max() < 64


def test_encode_exact_ties_lowest_index(cuda_device):
    """Duplicate codewords make exact ties in every level: the fp32 filter cannot separate them, so the kernel must
    fall back to its float64 evaluation and pick the LOWEST index (reference torch.argmin semantics, codebook.py:131);
    the result must still equal the exact restatement bit for bit, and a duplicated index must never be chosen."""
    from quantv2x_b200.engine import CodebookEngine

    C, m, ks, rows = 256, 1, [128] * 3, 3000
    cbs, heads = make_codebook_params(77, C, m, ks)
    for l in range(3):
        cbs[l][0, 64:, :] = cbs[l][0, :64, :]          # codeword 64 + i == codeword i
    p = oracle_params(cbs, heads)
    q = make_features(3, rows, C)
    delta = np.float32(0.21)
    eng = CodebookEngine(cbs, heads)
    lib_fe, fixed = library_fold(eng, co.fold_encode(p))
    codes = eng.encode(torch.from_numpy(q).to(cuda_device), float(delta)).cpu().numpy()
    ref_fx = co.encode_fixed_point(lib_fe, q, delta, fixed)
    for l in range(3):
        assert np.array_equal(codes[l].T, ref_fx[l]), f"level {l}: kernel codes differ from the fixed-point oracle"
    # the folded columns of duplicated codewords are identical bit for bit at level 0 (same digits, scale, g0)
    assert codes[0].max() < 64, "a duplicate (higher) index won an exact tie at level 0"


@pytest.mark.parametrize("case", [(64, 1, [128] * 3, 128, 5000), (64, 1, [64, 64], 64, 777), (256, 1, [128] * 3, 72 + 4, 1234)],
                         ids=lambda c: f"c{c[0]}_m{c[1]}_k{c[2][0]}x{len(c[2])}_o{c[3]}_r{c[4]}")
def test_decode_linear_fold(cuda_device, case):
    """qv2x_decode_linear (decode . 1x1 conv . ReLU . activation quantizer folded over the codeword tables, the entry
    of the pyramid model's ego stage) against the float64 evaluation on the library's fp32 decode tables: codes equal
    except where the float64 value sits within 1e-3 of a rounding boundary (then within 1 LSB), row sums consistent;
    and against the unfolded chain decode -> FP32 GEMM -> quantizing converter within 1 LSB on < 1e-3 of the codes."""
    from quantv2x_b200 import engine as E
    from tests.codebook_cases import make_codebook_params

    C, m, ks, cout, rows = case
    rng = np.random.default_rng(rows)
    cb = E.CodebookEngine(*make_codebook_params(11, C, m, ks))
    w = (rng.normal(size=(cout, C)) * 0.5).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32) * 0.1
    nt = len(ks) * m
    kk = [k for k in ks for _ in range(m)]
    codes = np.stack([rng.integers(0, kk[i], size=rows + 13, dtype=np.uint8) for i in range(nt)]).reshape(len(ks), m, -1)
    codes_d = torch.from_numpy(codes).to(cuda_device)
    feat = cb.decode(codes_d).cpu().numpy().astype(np.float64)[:rows]
    y = feat @ w.astype(np.float64).T + b
    delta = float(np.float32(np.abs(y).max() / 200.0))
    assert E.DecodeLinearEngine.supported(cb, cout)
    eng = E.DecodeLinearEngine(cb, w, b, delta)
    q, rs = eng.forward(codes_d, rows)
    q, rs = q.cpu().numpy().astype(np.int64), rs.cpu().numpy()
    t = y / delta
    ref = np.clip(np.rint(t), 0, 255).astype(np.int64)
    near = np.abs(np.abs(t - np.floor(t)) - 0.5) < 1e-3
    assert np.array_equal(q[~near], ref[~near])
    assert np.abs(q - ref).max() <= 1
    assert np.array_equal(rs, q.sum(axis=1))
    # the unfolded chain of this library
    planar = E.HeadsEngine(w, b).forward(cb.decode(codes_d)[:rows].contiguous())
    q2 = E.quantize_nchw_to_nhwc_u8(planar.view(1, cout, 1, rows), delta).view(rows, cout).cpu().numpy().astype(np.int64)
    d = np.abs(q - q2)
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
