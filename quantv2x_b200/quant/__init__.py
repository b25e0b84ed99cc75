"""Mirror of the reference's ``opencood.quant`` package surface for the quantized BEV path."""
from .fold_bn import search_fold_and_remove_bn  # noqa: F401
from .quant_block import (BaseQuantBlock, QuantBaseBEVBackbone, QuantDoubleConv, QuantDownsampleConv,  # noqa: F401
                          QuantPFNLayer, QuantPillarVFE, QuantPointPillar, opencood_specials,
                          specials_unquantized_names)
from .quant_layer import QuantModule, StraightThrough, UniformAffineQuantizer  # noqa: F401
from .quant_model import QuantModel  # noqa: F401
from .set_act_quantize_params import set_act_quantize_params  # noqa: F401
from .set_weight_quantize_params import save_quantized_weight, set_weight_quantize_params  # noqa: F401
