"""Mirror of opencood/quant/set_act_quantize_params.py:7-30: run calibration batches through the float
(fake-quant) path with un-initialised activation quantizers so they pick up delta / zero_point (EMA over
batches, quant_layer.py:102-108), then freeze them."""
from typing import Union

import torch

from .quant_block import BaseQuantBlock
from .quant_layer import QuantModule
from .quant_model import QuantModel


def set_act_quantize_params(module: Union[QuantModel, QuantModule, BaseQuantBlock], cached_inps, channel_sizes=None,
                            batch_size: int = 8):
    module.set_quant_state(True, True)
    quantizers = [t.act_quantizer for t in module.modules()
                  if isinstance(t, (QuantModule, BaseQuantBlock)) and hasattr(t, "act_quantizer")]
    for q in quantizers:
        q.set_inited(False)
    device = next(module.parameters()).device
    with torch.no_grad():
        if isinstance(cached_inps, torch.Tensor):
            for i in range(cached_inps.size(0)):
                module(cached_inps[i].to(device))
        else:
            for i in range(min(len(cached_inps), batch_size)):
                inp = cached_inps[i]
                module(inp.to(device) if isinstance(inp, torch.Tensor) else inp)
    for q in quantizers:
        q.set_inited(True)
