"""Mirror of opencood/quant/set_weight_quantize_params.py:13-24."""
from .quant_layer import QuantModule


def set_weight_quantize_params(model):
    """Initialise delta / zero_point of every weight quantizer from the (BN-folded) weights."""
    for module in model.modules():
        if isinstance(module, QuantModule):
            module.weight_quantizer.set_inited(False)
            module.weight_quantizer(module.weight)
            module.weight_quantizer.set_inited(True)


def save_quantized_weight(model):
    for module in model.modules():
        if isinstance(module, QuantModule):
            module.weight.data = module.weight_quantizer(module.weight)
