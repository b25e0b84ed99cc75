"""BatchNorm folding before quantization -- mirror of opencood/quant/fold_bn.py:40-81,161-175.
W' = W * gamma / sqrt(var + eps) along the OUTPUT channel axis (dim 0 for Conv2d/Linear, dim 1 for
ConvTranspose2d);  b' = beta - gamma * mean / std (+ gamma * b / std)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .quant_layer import StraightThrough

SKIP_NAMES = ("aligner_m1", "aligner_m2", "codebook")   # reference quant_block.py:1599-1615


def _is_bn(m):
    return isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d))


def _is_absorbing(m):
    return isinstance(m, (nn.Conv2d, nn.Linear, nn.ConvTranspose2d))


@torch.no_grad()
def fold_bn_into_conv(conv: nn.Module, bn: nn.Module):
    std = torch.sqrt(bn.running_var + bn.eps)
    gamma = bn.weight if bn.affine else torch.ones_like(std)
    beta = bn.bias if bn.affine else torch.zeros_like(std)
    ratio = gamma / std
    if isinstance(conv, nn.ConvTranspose2d):
        view = (1, -1, 1, 1)
    elif isinstance(conv, nn.Conv2d):
        view = (-1, 1, 1, 1)
    elif isinstance(conv, nn.Linear):
        view = (-1, 1)
    else:
        raise TypeError(f"cannot fold BN into {type(conv)}")
    w = conv.weight.data * ratio.view(view)
    b = beta - gamma * bn.running_mean / std
    if conv.bias is not None:
        b = gamma * conv.bias / std + b
        conv.bias.data = b
    else:
        conv.bias = nn.Parameter(b)
    conv.weight.data = w
    # the reference leaves the (now unused) BN holding these stats
    bn.running_mean = bn.bias.data if bn.affine else bn.running_mean
    bn.running_var = bn.weight.data ** 2 if bn.affine else bn.running_var


def search_fold_and_remove_bn(model: nn.Module):
    """Depth-first: whenever a BN directly follows a conv/linear inside the same container, fold and
    replace the BN by StraightThrough.  Returns the last absorbing module seen (reference behaviour)."""
    model.eval()
    prev = None
    for name, m in model.named_children():
        if name in SKIP_NAMES:
            continue
        if _is_bn(m) and _is_absorbing(prev):
            fold_bn_into_conv(prev, m)
            setattr(model, name, StraightThrough())
        elif _is_absorbing(m):
            prev = m
        else:
            prev = search_fold_and_remove_bn(m)
    return prev
