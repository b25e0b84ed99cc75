"""ORACLE (test infrastructure): the reference's multi-level residual multi-codebook quantizer, restated.

Follows opencood/models/sub_modules/codebook.py:
  * _multiCodebookQuantization._distance / encode     :106-131  (||x||^2 + ||c||^2 - 2 x.c, argmin over k)
  * _multiCodebookDeQuantization.decode                :192-201  (gather codeword per segment, concat)
  * _quantizerEncoder.encode                           :231-239  (z = E(x); code = argmin d(Q(z)); x' = L(z) - cb[code])
  * _quantizerDecoder.decode                           :263-269  (q = D(cb[code]); xhat = q + S(former); R(xhat))
  * UMGMQuantizer.encode / decode                      :330-343

Two evaluations:
  1. `encode_fp64` / `decode_fp64`: the canonical oracle -- the reference's sequential algorithm in float64 with
     lowest-index tie-break (np.argmin returns the first minimum, as torch.argmin does on CPU).
  2. `fold_encode` + `encode_fixed_point`: the exact arithmetic of the CUDA kernel.  Every head is affine, so the
     distance scores of all levels are affine in the input row and in the already selected codewords:
         score[l,s,k](x) = G[l][(s,k),:] . x + g0[l][(s,k)] + sum_{j<l} sum_{s'} B[l][j][s'][code[j][s'], (s,k)]
     (score = ||c||^2 - 2 h.c; the row constant ||h||^2 is dropped, it does not change the argmin).
     With x = delta * q (q uint8, the shrinker's output codes), G is carried as 24-bit fixed point per column (three
     signed base-256 digits -> three exact int32 accumulators, same trick as the transposed-conv layers) and the
     rest of the score is evaluated in float64.  Codes of this evaluation must equal the kernel's bit for bit, and
     equal the canonical oracle except on rows whose best/second-best gap is below the fixed-point resolution.
"""
from __future__ import annotations

import numpy as np

from .int_oracle import fixed_point_columns

HEADS = ("latentStageEncoder", "quantizationHead", "latentHead", "dequantizationHead", "sideHead", "restoreHead")


def params_from_state_dict(sd: dict, prefix: str = "") -> dict:
    """Pull a UMGMQuantizer's tensors out of a (reference-compatible) state_dict into float64 numpy.

    Returns {"levels": L, "m": m, "k": [k_l], "codebook": [L x [m,k,d]], "<head>": [L x (W, b) or None]}."""
    def get(name):
        v = sd.get(prefix + name)
        return None if v is None else np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, np.float64)

    levels = 0
    while get(f"_encoders.{levels}._quantizer._codebook") is not None:
        levels += 1
    p = {"levels": levels, "codebook": [get(f"_encoders.{l}._quantizer._codebook") for l in range(levels)]}
    p["m"] = p["codebook"][0].shape[0]
    p["k"] = [cb.shape[1] for cb in p["codebook"]]
    enc = {"latentStageEncoder": "_latentStageEncoder", "quantizationHead": "_quantizationHead",
           "latentHead": "_latentHead"}
    dec = {"dequantizationHead": "_dequantizationHead", "sideHead": "_sideHead", "restoreHead": "_restoreHead"}
    for head, attr in enc.items():
        p[head] = []
        for l in range(levels):
            W = get(f"_encoders.{l}.{attr}.weight")
            p[head].append(None if W is None else (W, get(f"_encoders.{l}.{attr}.bias")))
    for head, attr in dec.items():
        p[head] = []
        for l in range(levels):
            W = get(f"_decoders.{l}.{attr}.weight")
            p[head].append(None if W is None else (W, get(f"_decoders.{l}.{attr}.bias")))
    return p


def _lin(wb, x):
    W, b = wb
    return x @ W.T + b


def _gather(cb, code):
    """cb [m,k,d], code [n,m] -> [n, m*d]"""
    m = cb.shape[0]
    return np.concatenate([cb[s][code[:, s]] for s in range(m)], axis=1)


def distances(cb, h):
    """[n, m*d] x [m,k,d] -> [n,m,k] squared distances, the reference's expansion."""
    m, k, d = cb.shape
    hs = h.reshape(h.shape[0], m, d)
    x2 = (hs ** 2).sum(2, keepdims=True)
    c2 = (cb ** 2).sum(-1)
    inter = np.einsum("nmd,mkd->nmk", hs, cb)
    return x2 + c2[None] - 2 * inter


def encode_fp64(p, x, return_gaps=False):
    x = np.asarray(x, np.float64)
    codes, gaps = [], []
    for l in range(p["levels"]):
        z = _lin(p["latentStageEncoder"][l], x)
        d = distances(p["codebook"][l], _lin(p["quantizationHead"][l], z))
        code = d.argmin(-1)
        codes.append(code)
        if return_gaps:
            srt = np.sort(d, axis=-1)
            gaps.append((srt[..., 1] - srt[..., 0]) / np.maximum(np.abs(srt[..., 0]), 1e-30))
        if p["latentHead"][l] is not None:
            x = _lin(p["latentHead"][l], z) - _gather(p["codebook"][l], code)
    return (codes, gaps) if return_gaps else codes


def decode_fp64(p, codes):
    former = None
    for l in reversed(range(p["levels"])):
        q = _lin(p["dequantizationHead"][l], _gather(p["codebook"][l], codes[l]))
        xhat = q if (p["sideHead"][l] is None or former is None) else q + _lin(p["sideHead"][l], former)
        former = _lin(p["restoreHead"][l], xhat)
    return former


# ---------------------------------------------------------------------------------------------------
# folded (affine) forms -- what the CUDA kernels evaluate
# ---------------------------------------------------------------------------------------------------
def fold_encode(p):
    """Returns dict(G=[L x [m*k, C]], g0=[L x [m*k]], B=[L][j] -> [m(s'), k_j, m*k_l]) in float64."""
    L, m = p["levels"], p["m"]
    C = p["latentStageEncoder"][0][0].shape[1]
    d = C // m
    A = np.eye(C)                 # x_l = A x + a - sum_j P[j] cbvec_j
    a = np.zeros(C)
    P = []                        # P[j]: [C, C] matrix applied to the level-j codeword vector
    G, g0, B = [], [], []
    for l in range(L):
        We, be = p["latentStageEncoder"][l]
        Wq, bq = p["quantizationHead"][l]
        cb = p["codebook"][l]     # [m,k,d]
        k = cb.shape[1]
        # h = Wq (We x_l + be) + bq ; score[(s,k)] = ||c||^2 - 2 c . h_s
        M = Wq @ We
        hb = Wq @ be + bq
        Cmat = np.zeros((m * k, C))       # row (s,k) has c_{s,k} in segment s
        for s in range(m):
            Cmat[s * k:(s + 1) * k, s * d:(s + 1) * d] = cb[s]
        c2 = (cb ** 2).sum(-1).reshape(-1)
        T = -2.0 * Cmat @ M               # d score / d x_l
        G.append(T @ A)
        g0.append(c2 - 2.0 * Cmat @ hb + T @ a)
        Bl = []
        for j in range(l):
            kj = p["codebook"][j].shape[1]
            tab = np.zeros((m, kj, m * k))
            TP = -(T @ P[j])              # minus: x_l subtracts the codeword terms
            for s2 in range(m):
                tab[s2] = p["codebook"][j][s2] @ TP[:, s2 * d:(s2 + 1) * d].T
            Bl.append(tab)
        B.append(Bl)
        if p["latentHead"][l] is not None:
            Wl, bl = p["latentHead"][l]
            N = Wl @ We
            P = [N @ Pj for Pj in P] + [np.eye(C)]
            a = N @ a + Wl @ be + bl
            A = N @ A
    return {"G": G, "g0": g0, "B": B, "m": m, "k": p["k"], "levels": L, "C": C}


def fold_decode(p):
    """decode(codes) = const + sum_l sum_s T[l][s][code[l][:, s]]  (float64 tables [k, C])."""
    L, m = p["levels"], p["m"]
    C = p["restoreHead"][0][0].shape[0]
    d = p["codebook"][0].shape[2]
    # walk from the last level to the first, carrying former = F_const + sum of table terms
    tables = [None] * L
    const = None
    chain = None          # matrix applied to everything produced at deeper levels
    for l in reversed(range(L)):
        Wd, bd = p["dequantizationHead"][l]
        Wr, br = p["restoreHead"][l]
        if const is None:
            const = Wr @ bd + br
            for ll in range(l + 1, L):
                pass
            tables[l] = [p["codebook"][l][s] @ (Wr @ Wd[:, s * d:(s + 1) * d]).T for s in range(m)]
        else:
            Ws, bs = p["sideHead"][l]
            RS = Wr @ Ws
            for ll in range(l + 1, L):
                tables[ll] = [t @ RS.T for t in tables[ll]]
            const = Wr @ (bd + Ws @ const + bs) + br
            tables[l] = [p["codebook"][l][s] @ (Wr @ Wd[:, s * d:(s + 1) * d]).T for s in range(m)]
    return {"const": const, "tables": tables, "m": m, "levels": L, "C": C}


def decode_tables(fd, codes, dtype=np.float64):
    out = np.broadcast_to(fd["const"].astype(dtype), (codes[0].shape[0], fd["C"])).copy()
    for l in range(fd["levels"]):
        for s in range(fd["m"]):
            out += fd["tables"][l][s].astype(dtype)[codes[l][:, s]]
    return out


def quantize_fold(fe):
    """24-bit fixed point of every G row: digits int64 [3, N, C] and scales float64 [N], per level."""
    return [fixed_point_columns(Gl) for Gl in fe["G"]]


def encode_fixed_point(fe, q_u8, delta, fixed=None):
    """Exact restatement of the CUDA encode kernel.  q_u8 uint8 [n, C]; delta float32 activation scale."""
    if fixed is None:
        fixed = quantize_fold(fe)
    m, L = fe["m"], fe["levels"]
    q = q_u8.astype(np.float64)
    n = q.shape[0]
    codes = []
    for l in range(L):
        k = fe["k"][l]
        digits, sc = fixed[l]
        acc = [np.rint(q @ digits[g].T.astype(np.float64)).astype(np.int64) for g in range(3)]
        V = (acc[0] * 65536 + acc[1] * 256 + acc[2]).astype(np.float64)     # exact in the kernel (int64)
        score = V * (np.float64(np.float32(delta)) * sc)[None, :] + fe["g0"][l][None, :]
        for j in range(l):
            for s2 in range(m):
                score = score + fe["B"][l][j][s2][codes[j][:, s2]]
        codes.append(score.reshape(n, m, k).argmin(-1))
    return codes
