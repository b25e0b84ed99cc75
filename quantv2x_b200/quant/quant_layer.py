"""Host-side mirror of the reference's fake-quant primitives (opencood/quant/quant_layer.py).

Same class names, constructor arguments, attributes and numerical behaviour as the reference so that
calibration scripts written against ``opencood.quant`` keep working:

* ``UniformAffineQuantizer``  -- reference quant_layer.py:53-346 (asymmetric unsigned affine grid)
* ``QuantModule``             -- reference quant_layer.py:349-420

The torch code in this file is the CALIBRATION path (it is how delta / zero_point are produced,
SURVEY row a6).  Quantized INFERENCE does not run here: once every quantizer is initialised the
enclosing model exports its integer parameters to libqv2x.so (``quantv2x_b200.export``) and the
forward runs as hand-written sm_100a kernels; there is no CPU fallback for that path.
"""
from __future__ import annotations

from typing import Union

import torch
import torch.nn as nn
import torch.nn.functional as F


class StraightThrough(nn.Module):
    def forward(self, input):  # noqa: A002 - keeps the reference's argument name
        return input


def round_ste(x: torch.Tensor) -> torch.Tensor:
    """Round with a straight-through gradient."""
    return (x.round() - x).detach() + x


def lp_loss(pred, tgt, p=2.0, reduction="none"):
    err = (pred - tgt).abs().pow(p)
    return err.sum(1).mean() if reduction == "none" else err.mean()


class UniformAffineQuantizer(nn.Module):
    """q = clamp(round(x / delta) + zero_point, 0, 2**n_bits - 1);  x_hat = (q - zero_point) * delta."""

    def __init__(self, n_bits: int = 8, symmetric: bool = False, channel_wise: bool = False,
                 scale_method: str = "mse", leaf_param: bool = False, prob: float = 1.0):
        super().__init__()
        if symmetric:
            raise NotImplementedError("symmetric quantization is not supported (as in the reference)")
        assert 2 <= n_bits <= 8, "bitwidth not supported"
        self.sym = symmetric
        self.n_bits = n_bits
        self.n_levels = 2 ** n_bits
        self.delta = 1.0
        self.zero_point = 0.0
        self.inited = True
        self.leaf_param = leaf_param          # activation quantizer: EMA over calibration batches
        self.channel_wise = channel_wise
        self.eps = torch.tensor(1e-8, dtype=torch.float32)
        self.scale_method = scale_method
        self.one_side_dist = None
        self.num = 100
        self.running_min = None
        self.running_max = None
        self.prob = prob
        self.is_training = False

    # ------------------------------------------------------------------ state
    def set_inited(self, inited: bool = True):
        self.inited = inited

    def bitwidth_refactor(self, refactored_bit: int):
        assert 2 <= refactored_bit <= 8, "bitwidth not supported"
        self.n_bits = refactored_bit
        self.n_levels = 2 ** refactored_bit

    def update_quantize_range(self, x_min, x_max):
        if self.running_min is None:
            self.running_min, self.running_max = x_min, x_max
        self.running_min = 0.1 * x_min + 0.9 * self.running_min
        self.running_max = 0.1 * x_max + 0.9 * self.running_max
        return self.running_min, self.running_max

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor):
        if self.inited is False:
            self.delta, self.zero_point = self.init_quantization_scale(x.clone().detach(), self.channel_wise)
        x_int = round_ste(x / self.delta) + self.zero_point
        x_quant = torch.clamp(x_int, 0, self.n_levels - 1)
        x_dequant = (x_quant - self.zero_point) * self.delta
        if self.is_training and self.prob < 1.0:
            return torch.where(torch.rand_like(x) < self.prob, x_dequant, x)
        return x_dequant

    # ------------------------------------------------------------------ scale search
    def calculate_qparams(self, min_val, max_val):
        quant_min, quant_max = 0, self.n_levels - 1
        lo = torch.min(min_val, torch.zeros_like(min_val))
        hi = torch.max(max_val, torch.zeros_like(max_val))
        scale = torch.max((hi - lo) / float(quant_max - quant_min), self.eps.to(hi.device))
        zero_point = torch.clamp(quant_min - torch.round(lo / scale), quant_min, quant_max)
        return scale, zero_point

    def quantize(self, x, x_max, x_min):
        delta, zero_point = self.calculate_qparams(x_min, x_max)
        if self.channel_wise:
            shape = [x.shape[0]] + [1] * (x.dim() - 1)
            delta, zero_point = delta.reshape(shape), zero_point.reshape(shape)
        x_quant = torch.clamp(torch.round(x / delta) + zero_point, 0, self.n_levels - 1)
        return (x_quant - zero_point) * delta

    def _lp(self, pred, tgt, p=2.4):
        err = (pred - tgt).abs().pow(p)
        return err.flatten(1).mean(1) if self.channel_wise else err.mean()

    def _minmax(self, x, clamp_zero):
        if self.channel_wise:
            x_min, x_max = torch.aminmax(x.flatten(1), dim=1)
            if clamp_zero:
                x_max = torch.max(x_max, torch.zeros_like(x_max))
                x_min = torch.min(x_min, torch.zeros_like(x_min))
        else:
            x_min, x_max = torch.aminmax(x)
        return x_min, x_max

    def perform_2D_search(self, x):
        x_min, x_max = self._minmax(x, clamp_zero=True)
        if self.scale_method == "minmax":
            return x_min, x_max
        xrange = x_max - x_min
        best_score = torch.full_like(x_min, 1e10)
        best_min, best_max = x_min.clone(), x_max.clone()
        for i in range(1, self.num + 1):
            tmp_max = xrange / self.num * i
            tmp_delta = tmp_max / (2 ** self.n_bits - 1)
            for zp in range(0, self.n_levels):
                new_min, new_max = -zp * tmp_delta, tmp_max - zp * tmp_delta
                score = self._lp(x, self.quantize(x, new_max, new_min))
                better = score < best_score
                best_min = torch.where(better, new_min, best_min)
                best_max = torch.where(better, new_max, best_max)
                best_score = torch.min(best_score, score)
        return best_min, best_max

    def perform_1D_search(self, x):
        x_min, x_max = self._minmax(x, clamp_zero=False)
        if self.scale_method == "minmax":
            return x_min, x_max
        xrange = torch.max(x_min.abs(), x_max)
        best_score = torch.full_like(x_min, 1e10)
        best_min, best_max = x_min.clone(), x_max.clone()
        for i in range(1, self.num + 1):
            thres = xrange / self.num * i
            new_min = torch.zeros_like(x_min) if self.one_side_dist == "pos" else -thres
            new_max = torch.zeros_like(x_max) if self.one_side_dist == "neg" else thres
            score = self._lp(x, self.quantize(x, new_max, new_min))
            better = score < best_score
            best_min = torch.where(better, new_min, best_min)
            best_max = torch.where(better, new_max, best_max)
            best_score = torch.min(score, best_score)
        return best_min, best_max

    def get_x_min_x_max(self, x):
        if self.scale_method not in ("mse", "minmax"):
            raise NotImplementedError(f"scale_method {self.scale_method!r}")
        if self.one_side_dist is None:
            self.one_side_dist = "pos" if x.min() >= 0.0 else "neg" if x.max() <= 0.0 else "no"
            if self.one_side_dist != "no":
                best_min, best_max = self.perform_1D_search(x)
            else:
                best_min, best_max = self.perform_2D_search(x)
        else:
            best_min, best_max = self.perform_2D_search(x)
        if self.leaf_param:
            return self.update_quantize_range(best_min, best_max)
        return best_min, best_max

    def init_quantization_scale_channel(self, x):
        return self.calculate_qparams(*self.get_x_min_x_max(x))

    def init_quantization_scale(self, x_clone, channel_wise: bool = False):
        delta, zero_point = self.init_quantization_scale_channel(x_clone)
        if channel_wise:
            shape = [x_clone.shape[0]] + [1] * (x_clone.dim() - 1)
            delta, zero_point = delta.reshape(shape), zero_point.reshape(shape)
        return delta, zero_point

    def extra_repr(self):
        return f"bit={self.n_bits}, is_training={self.is_training}, inited={self.inited}"


class QuantModule(nn.Module):
    """Wraps nn.Conv2d / nn.ConvTranspose2d / nn.Linear with a weight and an (output) activation quantizer."""

    def __init__(self, org_module: Union[nn.Conv2d, nn.ConvTranspose2d, nn.Linear], weight_quant_params: dict = {},
                 act_quant_params: dict = {}, disable_act_quant=False):
        super().__init__()
        if isinstance(org_module, nn.Conv2d):
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                                   dilation=org_module.dilation, groups=org_module.groups)
            self.fwd_func = F.conv2d
        elif isinstance(org_module, nn.ConvTranspose2d):
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                                   output_padding=org_module.output_padding, groups=org_module.groups,
                                   dilation=org_module.dilation)
            self.fwd_func = F.conv_transpose2d
        else:
            self.fwd_kwargs = dict()
            self.fwd_func = F.linear
        self.weight = org_module.weight
        self.org_weight = org_module.weight.data.clone()
        if org_module.bias is not None:
            self.bias = org_module.bias
            self.org_bias = org_module.bias.data.clone()
        else:
            self.bias = None
            self.org_bias = None
        self.use_weight_quant = False
        self.use_act_quant = False
        self.weight_quantizer = UniformAffineQuantizer(**weight_quant_params)
        self.act_quantizer = UniformAffineQuantizer(**act_quant_params)
        self.norm_function = StraightThrough()
        self.activation_function = StraightThrough()
        self.ignore_reconstruction = False
        self.disable_act_quant = disable_act_quant
        self.trained = False

    def forward(self, input: torch.Tensor):  # noqa: A002
        if self.use_weight_quant:
            weight = self.weight_quantizer(self.weight).to(input.device)
            bias = self.bias.to(input.device) if self.bias is not None else None
        else:
            weight = self.org_weight.to(input.device)
            bias = self.org_bias.to(input.device) if self.org_bias is not None else None
        out = self.fwd_func(input, weight, bias, **self.fwd_kwargs)
        if type(self.norm_function) is nn.BatchNorm1d:
            out = self.norm_function(out.permute(0, 2, 1)).permute(0, 2, 1)
        else:
            out = self.norm_function(out)
        out = self.activation_function(out)
        if self.disable_act_quant:
            return out
        if self.use_act_quant:
            out = self.act_quantizer(out)
        return out

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant

    def extra_repr(self):
        return (f"wbit={self.weight_quantizer.n_bits}, abit={self.act_quantizer.n_bits}, "
                f"disable_act_quant={self.disable_act_quant}")

    # ------------------------------------------------------------------ export for libqv2x
    @torch.no_grad()
    def integer_weight(self):
        """(w_int uint8, delta[dim0], zero_point[dim0]) of the calibrated weight quantizer.

        Works for UniformAffineQuantizer and for AdaRound-style replacements alike: the integer grid is
        recovered from the quantizer's own output, w_int = round(w_hat / delta + zp)."""
        wq = self.weight_quantizer
        w_hat = wq(self.weight)
        delta = torch.as_tensor(wq.delta, dtype=torch.float32, device=w_hat.device)
        zp = torch.as_tensor(wq.zero_point, dtype=torch.float32, device=w_hat.device)
        shape = [w_hat.shape[0]] + [1] * (w_hat.dim() - 1)
        if delta.numel() == 1:
            delta = delta.reshape(1).expand(w_hat.shape[0])
            zp = zp.reshape(1).expand(w_hat.shape[0])
        delta, zp = delta.reshape(shape), zp.reshape(shape)
        w_int = torch.round(w_hat / delta + zp).clamp(0, wq.n_levels - 1)
        return (w_int.to(torch.uint8).cpu().numpy(), delta.reshape(-1).cpu().numpy().copy(),
                zp.reshape(-1).cpu().numpy().copy())
