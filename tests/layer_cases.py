"""Seeded random quantized layers shared by the GPU parity tests and smoke()."""
import numpy as np


def make_input(rng, n, h, w, c, density=0.5):
    x = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    x[rng.random((n, h, w, c)) > density] = 0
    return x


def make_conv(rng, cin, cout, k, w_bits=8, groups=1):
    qmax = 2 ** w_bits - 1
    w_zp = rng.integers(qmax // 2 - qmax // 6, qmax // 2 + qmax // 6 + 1, size=cout).astype(np.float32)
    w_int = np.clip(np.rint(w_zp.reshape(-1, 1, 1, 1) + rng.normal(0, qmax / 3.5, size=(cout, cin, k, k))),
                    0, qmax).astype(np.uint8)
    w_delta = (rng.uniform(0.5, 1.5, size=cout) * 2e-3 * 255 / qmax).astype(np.float32)
    bias = rng.uniform(-0.1, 0.1, size=cout).astype(np.float32)
    in_delta = rng.uniform(0.05, 0.3, size=groups).astype(np.float32)
    # keep a healthy spread of output codes: std of acc ~ sqrt(K)*q_rms*w_rms
    k_total = cin * k * k
    est = np.sqrt(k_total) * 100.0 * (qmax / 3.5) * float(w_delta.mean()) * float(in_delta.mean())
    out_delta = np.float32(est * 2.5 / 255.0)
    return dict(w_int=w_int, w_delta=w_delta, w_zp=w_zp, bias=bias, in_delta=in_delta, out_delta=out_delta)


def make_deconv(rng, cin, cout, s, w_bits=8):
    qmax = 2 ** w_bits - 1
    w_zp = rng.integers(qmax // 2 - qmax // 6, qmax // 2 + qmax // 6 + 1, size=cin).astype(np.float32)
    w_int = np.clip(np.rint(w_zp.reshape(-1, 1, 1, 1) + rng.normal(0, qmax / 3.5, size=(cin, cout, s, s))),
                    0, qmax).astype(np.uint8)
    w_delta = (rng.uniform(0.5, 1.5, size=cin) * 2e-3 * 255 / qmax).astype(np.float32)
    bias = rng.uniform(-0.1, 0.1, size=cout).astype(np.float32)
    in_delta = rng.uniform(0.05, 0.3, size=1).astype(np.float32)
    est = np.sqrt(cin) * 100.0 * (qmax / 3.5) * float(w_delta.mean()) * float(in_delta.mean())
    out_delta = np.float32(est * 2.5 / 255.0)
    return dict(w_int=w_int, w_delta=w_delta, w_zp=w_zp, bias=bias, in_delta=in_delta, out_delta=out_delta)
