"""A QuantModule of the mirror package whose weight quantizer is an AdaRound-style replacement (hard rounding
floor(w / delta) + (alpha >= 0), reference opencood/quant/adaptive_rounding.py:46-51) and whose activation scale is an
nn.Parameter (LSQ, block_recon.py:157-159), rebuilt from tests/golden/adaround_layer.npz."""
import os

import numpy as np
import torch
import torch.nn as nn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class HardAdaRound(nn.Module):
    """What block reconstruction leaves in place of UniformAffineQuantizer: same delta / zero_point / n_levels
    attributes, learned up/down decision per weight."""

    def __init__(self, uaq, alpha_nonneg: torch.Tensor):
        super().__init__()
        self.n_bits, self.n_levels, self.sym = uaq.n_bits, uaq.n_levels, uaq.sym
        self.delta, self.zero_point = uaq.delta, uaq.zero_point
        self.alpha = nn.Parameter(torch.where(alpha_nonneg, torch.ones(()), -torch.ones(())))
        self.inited = True
        self.soft_targets = False

    def set_inited(self, inited=True):
        self.inited = inited

    def forward(self, x):
        x_int = torch.floor(x / self.delta) + (self.alpha >= 0).float()
        x_quant = torch.clamp(x_int + self.zero_point, 0, self.n_levels - 1)
        return (x_quant - self.zero_point) * self.delta


def build():
    from quantv2x_b200.quant.quant_layer import QuantModule

    g = np.load(os.path.join(GOLD, "adaround_layer.npz"))
    cout, cin = g["w"].shape[:2]
    conv = nn.Conv2d(cin, cout, 3, padding=1)
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(g["w"]))
        conv.bias.copy_(torch.from_numpy(g["bias"]))
    qm = QuantModule(conv, dict(n_bits=8, channel_wise=True, scale_method="minmax"),
                     dict(n_bits=8, channel_wise=False, scale_method="minmax", leaf_param=True, prob=1.0)).eval()
    qm.activation_function = nn.ReLU()
    with torch.no_grad():
        qm.weight_quantizer.set_inited(False)
        qm.weight_quantizer(qm.weight)
        qm.weight_quantizer.set_inited(True)
    qm.weight_quantizer = HardAdaRound(qm.weight_quantizer, torch.from_numpy(g["alpha_nonneg"]))
    qm.act_quantizer.delta = nn.Parameter(torch.tensor(float(g["out_delta"])))
    qm.act_quantizer.zero_point = torch.tensor(0.0)
    qm.act_quantizer.set_inited(True)
    qm.set_quant_state(True, True)
    return qm, g
