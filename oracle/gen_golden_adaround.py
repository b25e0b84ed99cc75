"""Generates tests/golden/adaround_layer.npz by running the UNMODIFIED reference (/root/reference): one QuantModule
whose weight quantizer has been replaced by the reference's AdaRoundQuantizer in hard-rounding mode
(opencood/quant/adaptive_rounding.py:46-51: x_int = floor(w / delta) + (alpha >= 0)), as block reconstruction leaves
it (block_recon.py:123-131), and whose activation scale is an nn.Parameter as LSQ fine-tuning leaves it
(block_recon.py:157-159).  Pins the export path (QuantModule.integer_weight -> libqv2x) for reconstructed models.

Run in the build container only:  python oracle/gen_golden_adaround.py"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
from opencood.quant.adaptive_rounding import AdaRoundQuantizer  # noqa: E402
from opencood.quant.quant_layer import QuantModule  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
WQ = dict(n_bits=8, channel_wise=True, scale_method="minmax")
AQ = dict(n_bits=8, channel_wise=False, scale_method="minmax", leaf_param=True, prob=1.0)


def main():
    rng = np.random.default_rng(2024)
    cin, cout, H, W = 64, 64, 12, 20
    conv = nn.Conv2d(cin, cout, 3, padding=1)
    w = (rng.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / (cin * 9))).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, cout).astype(np.float32)
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(w))
        conv.bias.copy_(torch.from_numpy(b))
    qm = QuantModule(conv, WQ, AQ).eval()
    qm.activation_function = nn.ReLU()
    in_delta = np.float32(0.031)
    xq = rng.integers(0, 256, size=(2, cin, H, W)).astype(np.uint8)
    xq[rng.random(xq.shape) < 0.5] = 0
    x = torch.from_numpy(xq.astype(np.float32) * in_delta)
    with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
        qm.weight_quantizer.set_inited(False)
        qm.weight_quantizer(qm.weight)                      # init delta / zero_point (minmax)
        qm.weight_quantizer.set_inited(True)
        ada = AdaRoundQuantizer(uaq=qm.weight_quantizer, round_mode="learned_hard_sigmoid", weight_tensor=qm.org_weight.data)
        # a "learned" rounding: push a seeded 30 % of the alphas across zero
        flip = torch.from_numpy(rng.random(w.shape) < 0.3)
        ada.alpha.data = torch.where(flip, -ada.alpha.data, ada.alpha.data)
        ada.soft_targets = False
        qm.weight_quantizer = ada
        qm.set_quant_state(True, True)
        qm.act_quantizer.set_inited(False)
        qm(x)                                               # calibrate the activation scale
        qm.act_quantizer.set_inited(True)
        qm.act_quantizer.delta = nn.Parameter(torch.tensor(float(qm.act_quantizer.delta) * 1.0371))   # LSQ'd scale
        y = qm(x)
        w_hat = ada(qm.weight)
    od = float(qm.act_quantizer.delta)
    assert float(qm.act_quantizer.zero_point) == 0.0
    codes = torch.round(y / od)
    assert float((codes * od - y).abs().max()) < 1e-4 * od
    nearest = torch.round(qm.weight / ada.delta)
    hard = torch.floor(qm.weight / ada.delta) + (ada.alpha >= 0).float()
    print("weights whose AdaRound grid point differs from round-to-nearest:", float((nearest != hard).float().mean()))
    out = dict(w=w, bias=b, alpha_nonneg=(ada.alpha.detach() >= 0).numpy(), w_delta=ada.delta.detach().reshape(-1).numpy(),
               w_zp=ada.zero_point.detach().reshape(-1).numpy(), w_hat=w_hat.numpy(), xq=xq, in_delta=in_delta,
               out_delta=np.float32(od), out_codes=codes.numpy().astype(np.uint8))
    np.savez_compressed(os.path.join(OUT, "adaround_layer.npz"), **out)
    print("adaround_layer.npz", os.path.getsize(os.path.join(OUT, "adaround_layer.npz")))


if __name__ == "__main__":
    torch.manual_seed(0)
    main()
