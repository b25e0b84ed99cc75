"""``UMGMQuantizer`` -- mirror of the reference's multi-level residual multi-codebook compressor
(opencood/models/sub_modules/codebook.py:280-343) with identical parameter names / shapes, so reference
checkpoints load with ``load_state_dict``.

Only the deterministic inference interface is provided: ``encode`` (distance argmin) and ``decode``
(gather + heads), both executed by libqv2x (quantv2x_b200.engine.CodebookEngine).  The stochastic training
``forward`` (Gumbel-softmax sampling, codebook.py:147-182, 375-408) is out of scope and raises.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Union

import torch
from torch import nn

from . import engine as E

_COMPONENTS = ["latentStageEncoder", "quantizationHead", "latentHead", "dequantizationHead", "sideHead",
               "restoreHead"]


class _LowerBound(nn.Module):              # reference codebook_utils.LowerBound (buffer only; training op)
    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))


class _Quantization(nn.Module):            # reference _multiCodebookQuantization (parameters only)
    def __init__(self, codebook: nn.Parameter):
        super().__init__()
        self._m, self._k, self._d = codebook.shape
        self._codebook = codebook
        self._temperature = nn.Parameter(torch.ones((self._m, 1)))
        self._bound = _LowerBound(1e-6)


class _DeQuantization(nn.Module):          # reference _multiCodebookDeQuantization
    def __init__(self, codebook: nn.Parameter):
        super().__init__()
        self._m, self._k, self._d = codebook.shape
        self._codebook = codebook
        self.register_buffer("_ix", torch.arange(self._m), persistent=False)


class _QuantizerEncoder(nn.Module):
    def __init__(self, quantizer, dequantizer, latentStageEncoder, quantizationHead, latentHead):
        super().__init__()
        self._quantizer = quantizer
        self._dequantizer = dequantizer
        self._latentStageEncoder = latentStageEncoder
        self._quantizationHead = quantizationHead
        self._latentHead = latentHead

    @property
    def Codebook(self):
        return self._quantizer._codebook


class _QuantizerDecoder(nn.Module):
    def __init__(self, dequantizer, dequantizationHead, sideHead, restoreHead):
        super().__init__()
        self._dequantizer = dequantizer
        self._dequantizationHead = dequantizationHead
        self._sideHead = sideHead
        self._restoreHead = restoreHead


class UMGMQuantizer(nn.Module):
    def __init__(self, channel: int, m: int, k: Union[int, List[int]], permutationRate: float,
                 components: Dict[str, Callable[[], nn.Module]]):
        super().__init__()
        if isinstance(k, int):
            k = [k]
        self._m, self._k, self._channel = m, list(k), channel
        self.ema = 0.9
        self._freqEMA = nn.ParameterList(nn.Parameter(torch.ones(m, ki) / ki, requires_grad=False) for ki in k)
        fns = [components[key] for key in _COMPONENTS]
        encoders, decoders = [], []
        for i, ki in enumerate(k):
            last = i == len(k) - 1
            latentStageEncoder, quantizationHead = fns[0](), fns[1]()
            latentHead = None if last else fns[2]()
            dequantizationHead = fns[3]()
            sideHead = None if last else fns[4]()
            restoreHead = fns[5]()
            codebook = nn.Parameter(nn.init.normal_(torch.empty(m, ki, channel // m),
                                                    std=math.sqrt(2 / (5 * channel / m))))
            quantizer, dequantizer = _Quantization(codebook), _DeQuantization(codebook)
            encoders.append(_QuantizerEncoder(quantizer, dequantizer, latentStageEncoder, quantizationHead, latentHead))
            decoders.append(_QuantizerDecoder(dequantizer, dequantizationHead, sideHead, restoreHead))
        self._encoders = nn.ModuleList(encoders)
        self._decoders = nn.ModuleList(decoders)
        self._engine = None
        self._input_scale = None      # scale of the uint8 grid the input features lie on (set_input_scale)

    @property
    def Codebooks(self):
        return [enc.Codebook for enc in self._encoders]

    # ------------------------------------------------------------------ engine
    def head_params(self):
        """(codebooks, heads) in the layout CodebookEngine / the oracle take."""
        def wb(mod):
            if mod is None:
                return None
            if not isinstance(mod, nn.Linear):
                raise ValueError("codebook heads must be nn.Linear for the folded GPU path")
            return mod.weight.detach().cpu().numpy(), mod.bias.detach().cpu().numpy()

        cbs = [enc.Codebook.detach().cpu().numpy() for enc in self._encoders]
        heads = []
        for enc, dec in zip(self._encoders, self._decoders):
            heads.append({"latentStageEncoder": wb(enc._latentStageEncoder), "quantizationHead": wb(enc._quantizationHead),
                          "latentHead": wb(enc._latentHead), "dequantizationHead": wb(dec._dequantizationHead),
                          "sideHead": wb(dec._sideHead), "restoreHead": wb(dec._restoreHead)})
        return cbs, heads

    def engine(self) -> E.CodebookEngine:
        if self._engine is None:
            self._engine = E.CodebookEngine(*self.head_params())
        return self._engine

    def reset_engine(self):
        """Call after changing parameters (e.g. load_state_dict)."""
        self._engine = None

    def set_input_scale(self, delta: float | None):
        """Scale of the uint8 grid the encoder's input lies on (the shrinker's act_quantizer.delta).  Set by
        attach_engines, so that the reference's one-argument call ``codebook.encode(flattened)`` works unchanged."""
        self._input_scale = None if delta is None else float(delta)

    @staticmethod
    def _recover_grid_scale(x: torch.Tensor) -> float:
        """delta such that x = delta * q with q integer in [0, 255], recovered from on-grid float features: the
        smallest positive value is q_min * delta for some q_min in 1..255; the largest candidate that puts every
        element on an integer wins."""
        xf = x.detach().float()
        pos = xf[xf > 0]
        if pos.numel() == 0:
            return 1.0
        if bool((xf < 0).any()):
            raise ValueError("features are negative: not a post-ReLU uint8 grid; pass `delta` explicitly")
        v_min, v_max = float(pos.min()), float(pos.max())
        sample = xf.flatten()[:: max(1, xf.numel() // 65536)]
        for k in range(1, 256):
            d = v_min / k
            if v_max / d > 255.5:
                continue
            q = sample / d
            if float((q - torch.round(q)).abs().max()) <= 2e-3:
                q_all = xf / d
                if float((q_all - torch.round(q_all)).abs().max()) <= 2e-3:
                    return d
        raise ValueError("features do not lie on a uint8 grid: quantize them first or pass `delta`")

    # ------------------------------------------------------------------ float path (offline calibration only)
    @torch.no_grad()
    def encode_float(self, x: torch.Tensor) -> List[torch.Tensor]:
        """The reference's encode in PyTorch FP32 (codebook.py:106-131, 231-239, 330-337): for OFFLINE calibration
        forwards of models whose quantizers sit behind the codebook (the pyramid model).  Inference uses encode()."""
        codes = []
        for enc in self._encoders:
            z = enc._latentStageEncoder(x)
            h = enc._quantizationHead(z)
            cb = enc.Codebook                                           # [m, k, d]
            n = h.shape[0]
            hs = h.reshape(n, self._m, -1)
            d = (hs ** 2).sum(2, keepdim=True) + (cb ** 2).sum(-1)[None] - 2 * torch.einsum("nmd,mkd->nmk", hs, cb)
            code = d.argmin(-1)                                         # [n, m]
            codes.append(code)
            if enc._latentHead is not None:
                q = cb[torch.arange(self._m)[None], code].reshape(n, -1)
                x = enc._latentHead(z) - q
        return codes

    @torch.no_grad()
    def decode_float(self, codes) -> torch.Tensor:
        """The reference's decode in PyTorch FP32 (codebook.py:192-201, 263-269, 339-343); calibration only."""
        former = None
        for dec, code in zip(reversed(list(self._decoders)), reversed(list(codes))):
            cb = dec._dequantizer._codebook
            n = code.shape[0]
            q = dec._dequantizationHead(cb[torch.arange(self._m)[None], code.long()].reshape(n, -1))
            xhat = q if former is None else q + dec._sideHead(former)
            former = dec._restoreHead(xhat)
        return former

    # ------------------------------------------------------------------ reference interface
    def encode(self, x: torch.Tensor, delta: float | None = None) -> List[torch.Tensor]:
        """x: [n, C] -> list (levels) of LongTensor [n, m], the reference's signature and structure
        (opencood/models/sub_modules/codebook.py:330-337).

        The GPU encoder consumes uint8 activation codes and their scale.  ``x`` may be those codes (then ``delta``,
        or the scale registered with ``set_input_scale``, is required), or float32 features lying on a uint8 grid --
        what the quantized shrinker emits -- in which case the scale is taken from ``delta``, else from
        ``set_input_scale`` (attach_engines registers it), else recovered exactly from the values."""
        if delta is None:
            delta = self._input_scale
        if delta is None:
            if x.dtype == torch.uint8:
                raise ValueError("encode() of uint8 codes needs their scale: pass `delta` or call set_input_scale()")
            delta = self._recover_grid_scale(x)
        if x.dtype != torch.uint8:
            q = torch.round(x / delta)
            if not torch.equal(q * delta, x.to(q.dtype)) and (q * delta - x).abs().max() > 1e-4 * delta:
                raise ValueError("features are not on the uint8 grid given by `delta`")
            x = q.clamp(0, 255).to(torch.uint8)
        codes = self.engine().encode(x.contiguous(), delta)            # [levels, m, n]
        return [codes[l].t().long() for l in range(codes.shape[0])]

    def decode(self, codes) -> torch.Tensor:
        """codes: list (levels) of [n, m] integer tensors, or the packed uint8 [levels, m, n].  -> float32 [n, C]."""
        if isinstance(codes, (list, tuple)):
            codes = torch.stack([c.t() for c in codes]).to(torch.uint8)
        return self.engine().decode(codes.contiguous())

    def forward(self, x):
        raise NotImplementedError("the stochastic Gumbel-softmax training forward is out of scope; "
                                  "use encode()/decode() (reference codebook.py:330-343)")
