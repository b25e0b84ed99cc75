"""Seeded random UMGMQuantizer parameters (constructor-style init of the reference, codebook.py:290-324)."""
import numpy as np

HEADS = ("latentStageEncoder", "quantizationHead", "latentHead", "dequantizationHead", "sideHead", "restoreHead")


def make_codebook_params(seed, C, m, ks):
    rng = np.random.default_rng(seed)
    levels = len(ks)
    d = C // m
    bound = 1.0 / np.sqrt(C)
    codebooks, heads = [], []
    for l, k in enumerate(ks):
        codebooks.append(rng.normal(0, np.sqrt(2.0 / (5.0 * d)), size=(m, k, d)).astype(np.float32))
        h = {}
        for name in HEADS:
            if l == levels - 1 and name in ("latentHead", "sideHead"):
                h[name] = None
                continue
            h[name] = (rng.uniform(-bound, bound, size=(C, C)).astype(np.float32),
                       rng.uniform(-bound, bound, size=C).astype(np.float32))
        heads.append(h)
    return codebooks, heads


def oracle_params(codebooks, heads):
    """The dict layout oracle.codebook_oracle functions take."""
    p = {"levels": len(codebooks), "m": codebooks[0].shape[0], "k": [c.shape[1] for c in codebooks],
         "codebook": [c.astype(np.float64) for c in codebooks]}
    for name in HEADS:
        p[name] = [None if h[name] is None else (h[name][0].astype(np.float64), h[name][1].astype(np.float64))
                   for h in heads]
    return p


def make_features(seed, rows, C, density=0.5):
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 256, size=(rows, C), dtype=np.uint8)
    q[rng.random((rows, C)) > density] = 0
    return q
