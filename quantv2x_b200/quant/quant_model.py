"""``QuantModel`` -- mirror of opencood/quant/quant_model.py:7-147: folds BN, swaps registered modules for
their quantized wrappers, wraps stray Conv2d/Linear in ``QuantModule`` and toggles the quantization state."""
from __future__ import annotations

import torch.nn as nn

from .fold_bn import search_fold_and_remove_bn
from .quant_block import BaseQuantBlock, opencood_specials, specials_unquantized_names
from .quant_layer import QuantModule, StraightThrough, UniformAffineQuantizer


class QuantModel(nn.Module):
    def __init__(self, model: nn.Module, weight_quant_params: dict = {}, act_quant_params: dict = {},
                 is_fusing=True, skip_quant_module_names=None):
        super().__init__()
        self.skip_quant_module_names = tuple(skip_quant_module_names or [])
        self.weight_quant_params = dict(weight_quant_params)
        self.act_quant_params = dict(act_quant_params)
        if is_fusing:
            search_fold_and_remove_bn(model)
        self.model = model
        self._refactor(self.model, weight_quant_params, act_quant_params, fold=is_fusing)

    # the reference exposes both spellings
    def quant_module_refactor(self, module, weight_quant_params={}, act_quant_params={}, parent_name=""):
        self._refactor(module, weight_quant_params, act_quant_params, True, parent_name)

    def quant_module_refactor_wo_fuse(self, module, weight_quant_params={}, act_quant_params={}, parent_name=""):
        self._refactor(module, weight_quant_params, act_quant_params, False, parent_name)

    def _should_skip_quantization(self, full_name: str, local_name: str) -> bool:
        return any(local_name == s or full_name == s or full_name.startswith(f"{s}.")
                   for s in self.skip_quant_module_names)

    def _refactor(self, module, wq, aq, fold, parent_name=""):
        prev = None
        for name, child in module.named_children():
            full = f"{parent_name}.{name}" if parent_name else name
            if name in specials_unquantized_names or self._should_skip_quantization(full, name):
                continue
            if type(child) in opencood_specials:
                setattr(module, name, opencood_specials[type(child)](child, wq, aq))
            elif isinstance(child, (nn.Conv2d, nn.Linear)):
                prev = QuantModule(child, wq, aq)
                setattr(module, name, prev)
            elif isinstance(child, nn.BatchNorm2d) and not fold:
                if prev is not None:
                    prev.norm_function = child
                    setattr(module, name, StraightThrough())
            elif isinstance(child, (nn.ReLU, nn.ReLU6)):
                if prev is not None:
                    prev.activation_function = child
                    setattr(module, name, StraightThrough())
            elif isinstance(child, StraightThrough):
                continue
            else:
                self._refactor(child, wq, aq, fold, full)

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        for m in self.model.modules():
            if isinstance(m, (QuantModule, BaseQuantBlock)):
                m.set_quant_state(weight_quant, act_quant)

    def forward(self, input):  # noqa: A002
        return self.model(input)

    def set_first_last_layer_to_8bit(self):
        w_list, a_list = [], []
        for module in self.model.modules():
            if isinstance(module, UniformAffineQuantizer):
                (a_list if module.leaf_param else w_list).append(module)
        w_list[0].bitwidth_refactor(8)
        w_list[-1].bitwidth_refactor(8)
        a_list[-2].bitwidth_refactor(8)

    def disable_network_output_quantization(self):
        for name, module in self.model.named_modules():
            if isinstance(module, QuantModule) and name.rsplit(".", 1)[-1].startswith(("cls_head", "reg_head", "dir_head")):
                module.disable_act_quant = True

    def get_memory_footprint(self):
        total = sum(p.nelement() * p.element_size() for p in self.parameters())
        total += sum(b.nelement() * b.element_size() for b in self.buffers())
        return f"Model Memory Footprint: {total / 1024 ** 2:.2f} MB"
