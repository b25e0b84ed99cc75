"""ORACLE (test infrastructure): the attention-fusion ego stage in its FOLDED form -- what csrc/egostage.cu
(qv2x_ego_att) evaluates -- restated in float64 numpy, so that the algebra can be checked on the CPU against the
operator-by-operator restatement (codebook_oracle.decode_tables -> fusion_oracle.att_fusion -> fusion_oracle.heads); at the end of
the file the same for the folded entry of the pyramid model (decode -> conv1 -> quantizer, csrc/decode_linear.cu).

Operators folded (reference files):
  * UMGMQuantizer.decode        opencood/models/sub_modules/codebook.py:192-201, 263-269   f = const + sum_i T_i[code_i]
  * warp_affine_simple          opencood/models/sub_modules/torch_transformation_utils.py:323-332
  * AttFusion.forward           opencood/models/fuse_modules/fusion_in_one.py:126-151 (ego query row)
  * cls / reg / dir heads       opencood/models/heter_model_baseline_mc.py:137-142
With T' = the tables with the constant added to table 0 (every pixel carries exactly one code of table 0):
  Gram[r][r'] = <T'[r], T'[r']> / sqrt(C)          HT[r] = Wh T'[r]
  score_a(p)  = sum_t w_t(a,p) sum_t' w_t'(0,p) sum_{i,j} Gram[row_i(0, q_t')][row_j(a, q_t)]
  y(p)        = bias + sum_a softmax_a(score)(p) sum_t w_t(a,p) sum_j HT[row_j(a, q_t)]
Parity of this module is pinned by tests/test_oracle_cpu.py::test_ego_fold_equals_chain (it must equal the chain of
restatements that ARE pinned to the reference's golden vectors).
"""
from __future__ import annotations

import numpy as np


def fold_tables(tables, const, w_heads):
    """tables: list of [k_i, C] decode tables, const [C], w_heads [Cout, C] -> (Gram [R, R], HT [R, Cout], rowbase)."""
    C = const.shape[0]
    t = [np.asarray(x, np.float64).copy() for x in tables]
    t[0] = t[0] + np.asarray(const, np.float64)[None, :]
    rowbase = np.cumsum([0] + [x.shape[0] for x in t[:-1]])
    T = np.concatenate(t, axis=0)
    gram = (T @ T.T) * np.float64(np.float32(1.0) / np.sqrt(np.float32(C)))
    ht = T @ np.asarray(w_heads, np.float64).T
    return gram, ht, rowbase


def taps(aff_a, H, W):
    """Bilinear taps of warp_affine_simple for one agent: (weights [H, W, 4], source pixel index [H, W, 4]); taps
    outside the map have weight 0 and a clamped index."""
    M = np.asarray(aff_a, np.float64)
    xn = (2.0 * np.arange(W) + 1.0) / W - 1.0
    yn = (2.0 * np.arange(H) + 1.0) / H - 1.0
    xs = M[0, 0] * xn[None, :] + M[0, 1] * yn[:, None] + M[0, 2]
    ys = M[1, 0] * xn[None, :] + M[1, 1] * yn[:, None] + M[1, 2]
    ix = ((xs + 1.0) * W - 1.0) / 2.0
    iy = ((ys + 1.0) * H - 1.0) / 2.0
    x0, y0 = np.floor(ix).astype(np.int64), np.floor(iy).astype(np.int64)
    tx, ty = ix - x0, iy - y0
    wts, idx = [], []
    for dy in (0, 1):
        for dx in (0, 1):
            xx, yy = x0 + dx, y0 + dy
            w = (tx if dx else 1.0 - tx) * (ty if dy else 1.0 - ty)
            valid = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
            wts.append(np.where(valid, w, 0.0))
            idx.append(np.clip(yy, 0, H - 1) * W + np.clip(xx, 0, W - 1))
    return np.stack(wts, -1), np.stack(idx, -1)


def ego_att_folded(codes, tables, const, aff, w_heads, bias, H, W):
    """codes: int [NT, n*H*W] (agent-major rows), aff [n, 2, 3] -> head maps [Cout, H*W] (float64)."""
    nt = len(tables)
    hw = H * W
    n = codes.shape[1] // hw
    gram, ht, rowbase = fold_tables(tables, const, w_heads)
    rows = np.stack([rowbase[i] + np.asarray(codes[i], np.int64) for i in range(nt)])      # [NT, n*hw]
    wt, src = zip(*[taps(aff[a], H, W) for a in range(n)])
    wt = [w.reshape(hw, 4) for w in wt]
    src = [s.reshape(hw, 4) for s in src]
    # P[p][:] = sum over the ego's taps and tables of w0 * Gram[row][:]
    P = np.zeros((hw, gram.shape[0]))
    for t in range(4):
        for i in range(nt):
            P += wt[0][:, t, None] * gram[rows[i, src[0][:, t]]]
    score = np.zeros((n, hw))
    Y = np.zeros((n, hw, ht.shape[1]))
    pix = np.arange(hw)
    for a in range(n):
        for t in range(4):
            for j in range(nt):
                r = rows[j, a * hw + src[a][:, t]]
                score[a] += wt[a][:, t] * P[pix, r]
                Y[a] += wt[a][:, t, None] * ht[r]
    score -= score.max(axis=0, keepdims=True)
    p = np.exp(score)
    p /= p.sum(axis=0, keepdims=True)
    y = (p[..., None] * Y).sum(axis=0) + np.asarray(bias, np.float64)[None, :]
    return y.T


def decode_linear_folded(codes, tables, const, w, bias, delta):
    """The folded entry of the pyramid model's ego stage (csrc/decode_linear.cu, qv2x_decode_linear): decode
    (codebook.py:192-201) -> conv1 of the first QuantBottleneck (quant_block.py:100-134; 1x1 conv, ReLU, activation
    quantizer of scale delta, zero-point 0) as a sum of folded table rows.  codes int [NT, rows]; tables list of
    [k_i, C]; w [Cout, C].  Returns (uint8 codes [rows, Cout], their row sums, the pre-quantizer values)."""
    w64 = np.asarray(w, np.float64)
    b = (np.zeros(w64.shape[0]) if bias is None else np.asarray(bias, np.float64)) + w64 @ np.asarray(const, np.float64)
    y = np.tile(b, (codes.shape[1], 1))
    for i, t in enumerate(tables):
        y += (np.asarray(t, np.float64) @ w64.T)[np.asarray(codes[i], np.int64)]
    q = np.clip(np.rint(y / np.float64(delta)), 0, 255).astype(np.uint8)
    return q, q.astype(np.int64).sum(axis=1), y
