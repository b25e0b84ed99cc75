"""Python handles over the C-ABI objects of libqv2x.so.  PyTorch is only the allocator / stream
provider here: tensors are passed down as raw device pointers."""
from __future__ import annotations

import ctypes
from ctypes import byref, c_int, c_void_p

import numpy as np
import torch

from . import _lib
from ._lib import LayerDesc, check


def _stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(c_void_p)


class QLayer:
    """One quantized conv / transposed-conv layer living on the GPU (qv2x_layer).

    Parameters mirror what a calibrated reference ``QuantModule`` holds
    (opencood/quant/quant_layer.py:349-410): integer weight grid, per-dim-0 weight delta / zero-point,
    bias, the activation delta(s) of the tensor(s) feeding the layer and the layer's own act quantizer.
    """

    def __init__(self, *, kind, w_int, w_delta, w_zp, bias, ksize, stride, pad, w_bits, relu, in_delta,
                 out_delta, out_zp=0.0, out_bits=8, groups=1):
        w_int = np.ascontiguousarray(w_int, dtype=np.uint8)
        w_delta = np.ascontiguousarray(w_delta, dtype=np.float32).reshape(-1)
        w_zp = np.ascontiguousarray(w_zp, dtype=np.float32).reshape(-1)
        in_delta = [float(v) for v in np.atleast_1d(np.asarray(in_delta, dtype=np.float32))]
        d = LayerDesc()
        d.kind = int(kind)
        d.groups = int(groups)
        if kind == 0:
            d.cout, d.cin = int(w_int.shape[0]), int(w_int.shape[1]) * int(groups)   # [cout][cin/groups][k][k]
        else:
            d.cin, d.cout = int(w_int.shape[0]), int(w_int.shape[1])
        d.ksize, d.stride, d.pad = int(ksize), int(stride), int(pad)
        d.w_bits, d.relu = int(w_bits), int(bool(relu))
        d.n_in_groups = len(in_delta)
        for i in range(3):
            d.in_delta[i] = in_delta[i] if i < len(in_delta) else 0.0
        d.out_delta, d.out_zero_point, d.out_bits = float(out_delta), float(out_zp), int(out_bits)
        assert w_delta.size == w_int.shape[0] and w_zp.size == w_int.shape[0]
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
        self.desc = d
        # everything needed to rebuild the layer (quantv2x_b200.serialize)
        self.spec = dict(kind=int(kind), w_int=w_int, w_delta=w_delta, w_zp=w_zp, bias=b, ksize=int(ksize),
                         stride=int(stride), pad=int(pad), w_bits=int(w_bits), relu=bool(relu),
                         in_delta=np.asarray(in_delta, np.float32), out_delta=float(out_delta), out_zp=float(out_zp),
                         out_bits=int(out_bits), groups=int(groups))
        self._h = c_void_p()
        check(_lib.lib().qv2x_layer_create(byref(d), _np_ptr(w_int), _np_ptr(w_delta), _np_ptr(w_zp),
                                            None if b is None else _np_ptr(b), byref(self._h)))
        self.needs_rowsum = bool(_lib.lib().qv2x_layer_needs_rowsum(self._h))
        self.kind, self.cin, self.cout = d.kind, d.cin, d.cout
        self.n_groups = 3 if (kind == 1 or d.n_in_groups == 3) else 1
        self.n_cols = d.cout * (d.stride * d.stride if kind == 1 else 1)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().qv2x_layer_destroy(h)
            except Exception:
                pass
            self._h = None

    def out_shape(self, hi, wi):
        ho, wo = c_int(), c_int()
        check(_lib.lib().qv2x_layer_out_shape(self._h, hi, wi, byref(ho), byref(wo)))
        return ho.value, wo.value

    def forward(self, x: torch.Tensor, *, in_cbase=0, rowsum_in=None, out=None, out_cbase=0, rowsum_out=None,
                acc_dump=None, residual=None, res_delta=None, res_cbase=0, out_f32=None):
        """x: uint8 NHWC [n, H, W, Cstride] on the GPU.  Returns the uint8 NHWC output tensor.

        Residual-block forms (reference QuantBottleneck.forward, quant_block.py:124-134): ``residual`` is added before
        the ReLU and the output quantizer -- uint8 NHWC codes with scale ``res_delta`` (identity shortcut) or a float32
        NHWC tensor (the downsample conv's output); ``out_f32`` (float32 NHWC [n, Ho, Wo, >= cout]) makes this a
        layer without an output quantizer and is returned instead of codes."""
        assert x.is_cuda and x.dtype == torch.uint8 and x.is_contiguous() and x.dim() == 4
        n, hi, wi, cs = x.shape
        ho, wo = self.out_shape(hi, wi)
        extra = None
        if residual is not None or out_f32 is not None:
            extra = _lib.LayerExtra()
            if residual is not None:
                assert residual.is_cuda and residual.is_contiguous() and tuple(residual.shape[:3]) == (n, ho, wo)
                if residual.dtype == torch.uint8:
                    extra.d_res_u8, extra.res_delta = residual.data_ptr(), float(res_delta)
                else:
                    assert residual.dtype == torch.float32
                    extra.d_res_f32 = residual.data_ptr()
                extra.res_cstride, extra.res_cbase = residual.shape[3], int(res_cbase)
            if out_f32 is not None:
                assert out_f32.is_cuda and out_f32.dtype == torch.float32 and out_f32.is_contiguous()
                assert tuple(out_f32.shape[:3]) == (n, ho, wo) and rowsum_out is None
                extra.d_out_f32, extra.out_f32_cstride = out_f32.data_ptr(), out_f32.shape[3]
        if out is None and out_f32 is None:
            out = torch.empty((n, ho, wo, self.cout), dtype=torch.uint8, device=x.device)
        if out is not None:
            assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous()
            assert tuple(out.shape[:3]) == (n, ho, wo)
        rs_arr = None
        if self.needs_rowsum:
            if rowsum_in is None:
                g = self.desc.n_in_groups
                cg = self.cin // g
                rowsum_in = [rowsum_u8(x, in_cbase + i * cg, cg) for i in range(g)]
            rs_arr = (c_void_p * len(rowsum_in))(*[c_void_p(t.data_ptr()) for t in rowsum_in])
        check(_lib.lib().qv2x_layer_forward_ex(
            self._h, n, hi, wi, c_void_p(x.data_ptr()), cs, in_cbase, rs_arr,
            None if out is None else c_void_p(out.data_ptr()), 16 if out is None else out.shape[3], out_cbase,
            None if rowsum_out is None else c_void_p(rowsum_out.data_ptr()),
            None if acc_dump is None else c_void_p(acc_dump.data_ptr()),
            None if extra is None else byref(extra), _stream_ptr()))
        return out if out_f32 is None else out_f32


def rowsum_u8(x: torch.Tensor, cbase: int, c: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """Per-pixel channel sums (int32 [n, H, W]) of channels [cbase, cbase+c) of a uint8 NHWC tensor."""
    assert x.is_cuda and x.dtype == torch.uint8 and x.is_contiguous() and x.dim() == 4
    n, h, w, cs = x.shape
    if out is None:
        out = torch.empty((n, h, w), dtype=torch.int32, device=x.device)
    check(_lib.lib().qv2x_rowsum_u8(c_void_p(x.data_ptr()), n * h * w, cs, cbase, c, c_void_p(out.data_ptr()),
                                    _stream_ptr()))
    return out


HEAD_ORDER = ("latentStageEncoder", "quantizationHead", "latentHead", "dequantizationHead", "sideHead",
              "restoreHead")


class CodebookEngine:
    """GPU codebook compressor (qv2x_codebook): UMGMQuantizer.encode / .decode of the reference
    (opencood/models/sub_modules/codebook.py:330-343).

    codebooks: list (levels) of float32 [m, k, C/m]; heads: list (levels) of dicts name -> (W [C,C], b [C]) or None.
    """

    def __init__(self, codebooks, heads):
        from ._lib import CodebookDesc

        levels = len(codebooks)
        cbs = [np.ascontiguousarray(c, dtype=np.float32) for c in codebooks]
        self.spec = dict(codebooks=cbs, heads=heads)          # quantv2x_b200.serialize
        m, _, dseg = cbs[0].shape
        d = CodebookDesc()
        d.channel, d.m, d.levels = int(m * dseg), int(m), int(levels)
        for l in range(levels):
            d.k[l] = int(cbs[l].shape[1])
        keep = list(cbs)
        cb_ptrs = (c_void_p * levels)(*[_np_ptr(c) for c in cbs])
        w_ptrs = (c_void_p * (levels * 6))()
        b_ptrs = (c_void_p * (levels * 6))()
        for l in range(levels):
            for h, name in enumerate(HEAD_ORDER):
                wb = heads[l].get(name)
                if wb is None:
                    continue
                w = np.ascontiguousarray(wb[0], dtype=np.float32)
                b = np.ascontiguousarray(wb[1], dtype=np.float32)
                keep += [w, b]
                w_ptrs[l * 6 + h] = w.ctypes.data
                b_ptrs[l * 6 + h] = b.ctypes.data
        self.desc = d
        self.channel, self.m, self.levels = d.channel, d.m, d.levels
        self.k = [d.k[l] for l in range(levels)]
        self._h = c_void_p()
        check(_lib.lib().qv2x_codebook_create(byref(d), cb_ptrs, w_ptrs, b_ptrs, byref(self._h)))
        del keep

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().qv2x_codebook_destroy(h)
            except Exception:
                pass
            self._h = None

    def encode(self, feat_u8: torch.Tensor, delta: float, out: torch.Tensor | None = None) -> torch.Tensor:
        """feat_u8: uint8 [rows, Cstride] (or any [..., Cstride] contiguous).  Returns uint8 [levels, m, rows]."""
        assert feat_u8.is_cuda and feat_u8.dtype == torch.uint8 and feat_u8.is_contiguous()
        cs = feat_u8.shape[-1]
        rows = feat_u8.numel() // cs
        if out is None:
            out = torch.empty((self.levels, self.m, rows), dtype=torch.uint8, device=feat_u8.device)
        check(_lib.lib().qv2x_codebook_encode(self._h, rows, c_void_p(feat_u8.data_ptr()), cs, float(delta),
                                              c_void_p(out.data_ptr()), _stream_ptr()))
        return out

    def decode(self, codes: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """codes: uint8 [levels, m, rows] -> float32 [rows, C]."""
        assert codes.is_cuda and codes.dtype == torch.uint8 and codes.is_contiguous()
        rows = codes.shape[-1]
        if out is None:
            out = torch.empty((rows, self.channel), dtype=torch.float32, device=codes.device)
        check(_lib.lib().qv2x_codebook_decode(self._h, rows, c_void_p(codes.data_ptr()), c_void_p(out.data_ptr()),
                                              _stream_ptr()))
        return out

    def decode_regions(self, codes: torch.Tensor, pitch: int, base_rows, rects, out: torch.Tensor) -> torch.Tensor:
        """Decode only rectangles of the row grid (see qv2x_codebook_decode_regions).  codes uint8 [levels, m, rows];
        base_rows: first row of each region's image; rects: (y0, y1, x0, x1) per region; out float32 [rows, C]."""
        br = np.ascontiguousarray(base_rows, dtype=np.int64)
        rc = np.ascontiguousarray(rects, dtype=np.int32).reshape(-1)
        check(_lib.lib().qv2x_codebook_decode_regions(self._h, codes.shape[-1], pitch, len(br), _np_ptr(br),
                                                      _np_ptr(rc), c_void_p(codes.data_ptr()),
                                                      c_void_p(out.data_ptr()), _stream_ptr()))
        return out

    def folded(self, which: int) -> np.ndarray:
        """Test hook: the folded tables held by the library (see qv2x_codebook_folded_copy)."""
        n = _lib.lib().qv2x_codebook_folded_size(self._h, which)
        dt = {0: np.int8, 1: np.float64, 2: np.float64, 3: np.float64, 4: np.float32, 5: np.float32}[which]
        buf = np.empty(n, dtype=dt)
        check(_lib.lib().qv2x_codebook_folded_copy(self._h, which, _np_ptr(buf)))
        return buf


def fuse_tile(feat: torch.Tensor, affine: torch.Tensor, mode: str, tile, out: torch.Tensor) -> torch.Tensor:
    """Warp + fuse the output tile (y0, y1, x0, x1) only; out is compact float32 [(y1-y0)*(x1-x0), C]."""
    n, h, w, c = feat.shape
    y0, y1, x0, x1 = tile
    aff = affine.to(torch.float32).reshape(n, 6).contiguous()
    check(_lib.lib().qv2x_fuse_tile({"max": 0, "att": 1}[mode], n, h, w, c, c_void_p(feat.data_ptr()),
                                    c_void_p(aff.data_ptr()), c_void_p(out.data_ptr()), y0, y1, x0, x1, _stream_ptr()))
    return out


def fuse(feat: torch.Tensor, affine, mode: str, out: torch.Tensor | None = None) -> torch.Tensor:
    """Warp + fuse one frame.  feat float32 [N, H, W, C] pixel-major (agent 0 = ego); affine [N, 2, 3]
    normalized (CUDA tensor, or a host array that is copied over); mode 'max' | 'att'.  Returns float32 [H, W, C]."""
    assert feat.is_cuda and feat.dtype == torch.float32 and feat.is_contiguous() and feat.dim() == 4
    n, h, w, c = feat.shape
    if not (isinstance(affine, torch.Tensor) and affine.is_cuda):
        affine = torch.as_tensor(np.asarray(affine, dtype=np.float32)).to(feat.device)
    aff = affine.to(torch.float32).reshape(n, 6).contiguous()
    if out is None:
        out = torch.empty((h, w, c), dtype=torch.float32, device=feat.device)
    check(_lib.lib().qv2x_fuse({"max": 0, "att": 1}[mode], n, h, w, c, c_void_p(feat.data_ptr()),
                               c_void_p(aff.data_ptr()), c_void_p(out.data_ptr()), _stream_ptr()))
    return out


def fuse_weighted(feat: torch.Tensor, score: torch.Tensor, affine, score_is_logit: bool = True,
                  out: torch.Tensor | None = None) -> torch.Tensor:
    """Score-weighted fusion of one pyramid level (reference weighted_fuse, pyramid_fuse.py:17-62).
    feat float32 [N, H, W, C] pixel-major (agent 0 = ego); score float32 [N, H, W]: the occupancy logits
    (score_is_logit, the kernel applies sigmoid + 1e-4 as quant_block.py:520 does) or ready-made scores;
    affine [N, 2, 3] normalized.  Returns float32 [H, W, C]."""
    assert feat.is_cuda and feat.dtype == torch.float32 and feat.is_contiguous() and feat.dim() == 4
    n, h, w, c = feat.shape
    assert score.is_cuda and score.dtype == torch.float32 and score.is_contiguous() and score.numel() == n * h * w
    if not (isinstance(affine, torch.Tensor) and affine.is_cuda):
        affine = torch.as_tensor(np.asarray(affine, dtype=np.float32)).to(feat.device)
    aff = affine.to(torch.float32).reshape(n, 6).contiguous()
    if out is None:
        out = torch.empty((h, w, c), dtype=torch.float32, device=feat.device)
    check(_lib.lib().qv2x_fuse_weighted(n, h, w, c, c_void_p(feat.data_ptr()), c_void_p(score.data_ptr()),
                                        1 if score_is_logit else 0, c_void_p(aff.data_ptr()),
                                        c_void_p(out.data_ptr()), _stream_ptr()))
    return out


def fuse_weighted_u8(codes: torch.Tensor, delta: float, score: torch.Tensor, affine, score_is_logit: bool = True,
                     out: torch.Tensor | None = None) -> torch.Tensor:
    """fuse_weighted on the level's uint8 codes [N, H, W, C] of scale delta: the same result as
    fuse_weighted(dequantize_u8(codes, delta), ...) bit for bit, without the FP32 copy (qv2x_fuse_weighted_u8)."""
    assert codes.is_cuda and codes.dtype == torch.uint8 and codes.is_contiguous() and codes.dim() == 4
    n, h, w, c = codes.shape
    assert c % 4 == 0
    assert score.is_cuda and score.dtype == torch.float32 and score.is_contiguous() and score.numel() == n * h * w
    if not (isinstance(affine, torch.Tensor) and affine.is_cuda):
        affine = torch.as_tensor(np.asarray(affine, dtype=np.float32)).to(codes.device)
    aff = affine.to(torch.float32).reshape(n, 6).contiguous()
    if out is None:
        out = torch.empty((h, w, c), dtype=torch.float32, device=codes.device)
    check(_lib.lib().qv2x_fuse_weighted_u8(n, h, w, c, c_void_p(codes.data_ptr()), float(delta),
                                           c_void_p(score.data_ptr()), 1 if score_is_logit else 0,
                                           c_void_p(aff.data_ptr()), c_void_p(out.data_ptr()), _stream_ptr()))
    return out


class HeadsEngine:
    """cls/reg/dir 1x1 heads as one [Cout, Cin] FP32 GEMM (qv2x_heads)."""

    def __init__(self, w: np.ndarray, bias: np.ndarray | None):
        w = np.ascontiguousarray(w, dtype=np.float32)
        self.cout, self.cin = w.shape
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
        self.spec = dict(w=w, bias=b)                         # quantv2x_b200.serialize
        self._h = c_void_p()
        check(_lib.lib().qv2x_heads_create(self.cin, self.cout, _np_ptr(w), None if b is None else _np_ptr(b),
                                            byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().qv2x_heads_destroy(h)
            except Exception:
                pass
            self._h = None

    def forward(self, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """x float32 [..., Cin] pixel-major -> float32 [Cout, pixels]."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[-1] == self.cin
        pixels = x.numel() // self.cin
        if out is None:
            out = torch.empty((self.cout, pixels), dtype=torch.float32, device=x.device)
        check(_lib.lib().qv2x_heads_forward(self._h, pixels, c_void_p(x.data_ptr()), c_void_p(out.data_ptr()),
                                            _stream_ptr()))
        return out


class EgoAttEngine:
    """The ego stage of an attention-fusion frame as one kernel (qv2x_ego_att): code planes -> head maps, i.e.
    UMGMQuantizer.decode (codebook.py:192-201) + warp_affine_simple + AttFusion.forward (fusion_in_one.py:126-151) +
    the cls/reg/dir heads, folded over the codeword tables (no feature map is formed)."""

    @staticmethod
    def supported(codebook: "CodebookEngine", heads: "HeadsEngine") -> bool:
        return bool(_lib.lib().qv2x_ego_att_supported(codebook._h, heads.cout)) and heads.cin == codebook.channel

    def __init__(self, codebook: "CodebookEngine", heads: "HeadsEngine"):
        self.cout = heads.cout
        self.nt = codebook.levels * codebook.m
        self._h = c_void_p()
        w, b = heads.spec["w"], heads.spec["bias"]
        check(_lib.lib().qv2x_ego_att_create(codebook._h, heads.cout, _np_ptr(w), None if b is None else _np_ptr(b),
                                              byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().qv2x_ego_att_destroy(h)
            except Exception:
                pass
            self._h = None

    def forward(self, codes: torch.Tensor, affine: torch.Tensor, n: int, h: int, w: int,
                out: torch.Tensor | None = None) -> torch.Tensor:
        """codes uint8 [levels, m, rows >= n*h*w] (agent-major rows, agent 0 = ego); affine CUDA float32 [n, 2, 3]
        -> float32 [Cout, h*w]."""
        assert codes.is_cuda and codes.dtype == torch.uint8 and codes.is_contiguous()
        assert codes.numel() // codes.shape[-1] == self.nt
        assert affine.is_cuda and affine.dtype == torch.float32 and affine.is_contiguous() and affine.numel() >= 6 * n
        if out is None:
            out = torch.empty((self.cout, h * w), dtype=torch.float32, device=codes.device)
        check(_lib.lib().qv2x_ego_att_forward(self._h, n, h, w, c_void_p(codes.data_ptr()), codes.shape[-1],
                                              c_void_p(affine.data_ptr()), c_void_p(out.data_ptr()), _stream_ptr()))
        return out


class DecodeLinearEngine:
    """Codebook decode + 1x1 conv + activation quantizer folded over the codeword tables (qv2x_decode_linear): code
    planes -> uint8 NHWC codes of the conv's output and their per-pixel sums.  Entry of the pyramid model's ego stage
    (UMGMQuantizer.decode -> conv1 of the first QuantBottleneck, quant_block.py:100-134)."""

    @staticmethod
    def supported(codebook: "CodebookEngine", cout: int) -> bool:
        return bool(_lib.lib().qv2x_decode_linear_supported(codebook._h, int(cout)))

    def __init__(self, codebook: "CodebookEngine", w_hat: np.ndarray, bias, out_delta: float):
        w = np.ascontiguousarray(w_hat, dtype=np.float32)
        assert w.shape[1] == codebook.channel
        self.cout = int(w.shape[0])
        self.nt = codebook.levels * codebook.m
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
        self._h = c_void_p()
        check(_lib.lib().qv2x_decode_linear_create(codebook._h, self.cout, _np_ptr(w), None if b is None else _np_ptr(b),
                                                    float(out_delta), byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().qv2x_decode_linear_destroy(h)
            except Exception:
                pass
            self._h = None

    def forward(self, codes: torch.Tensor, rows: int, want_rowsum: bool = True):
        """codes uint8 [levels, m, >= rows] -> (uint8 [rows, cout], int32 [rows] or None)."""
        assert codes.is_cuda and codes.dtype == torch.uint8 and codes.is_contiguous()
        assert codes.numel() // codes.shape[-1] == self.nt and codes.shape[-1] >= rows
        out = torch.empty((rows, self.cout), dtype=torch.uint8, device=codes.device)
        rs = torch.empty((rows,), dtype=torch.int32, device=codes.device) if want_rowsum else None
        check(_lib.lib().qv2x_decode_linear_forward(self._h, rows, c_void_p(codes.data_ptr()), codes.shape[-1],
                                                    c_void_p(out.data_ptr()),
                                                    None if rs is None else c_void_p(rs.data_ptr()), _stream_ptr()))
        return out, rs


def heads_forward_deconv_u8(heads: "HeadsEngine", x: torch.Tensor, stride: int, out_delta: float, out: torch.Tensor,
                            out_cbase: int = 0) -> torch.Tensor:
    """Transposed conv (kernel = stride) on an FP32 map [h, w, cin] as one GEMM whose epilogue applies ReLU + the
    activation quantizer and the pixel shuffle: codes go to out[:, :, out_cbase : out_cbase + c] of the uint8 NHWC
    buffer [h*stride, w*stride, C] (qv2x_heads_forward_deconv_u8)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 3 and x.shape[-1] == heads.cin
    assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and out.dim() == 3
    h, w, _ = x.shape
    assert out.shape[0] == h * stride and out.shape[1] == w * stride
    check(_lib.lib().qv2x_heads_forward_deconv_u8(heads._h, h, w, c_void_p(x.data_ptr()), int(stride), float(out_delta),
                                                  c_void_p(out.data_ptr()), out.shape[2], int(out_cbase),
                                                  _stream_ptr()))
    return out


class PillarEngine:
    """libqv2x handle of the quantized PointPillars front end (qv2x_pillar_*): pillars -> uint8 NHWC BEV map."""

    def __init__(self, w_hat: np.ndarray, bias, *, nx: int, ny: int, voxel_size, offset, pre_quant, out_quant):
        """w_hat float32 [64, 10] fake-quantized weights (BN folded); bias [64] or None;
        pre_quant: None or (delta, zero_point, bits) of the Linear's own (pre-ReLU) activation quantizer;
        out_quant: (delta, zero_point, bits) of the PFN block's post-ReLU quantizer."""
        from ._lib import PillarDesc

        w_hat = np.ascontiguousarray(w_hat, dtype=np.float32)
        self.spec = dict(w_hat=w_hat, bias=None if bias is None else np.ascontiguousarray(bias, dtype=np.float32),
                         nx=int(nx), ny=int(ny), voxel_size=tuple(float(v) for v in voxel_size),
                         offset=tuple(float(v) for v in offset), pre_quant=pre_quant, out_quant=out_quant)
        d = PillarDesc()
        d.cout, d.n_feat, d.max_points = int(w_hat.shape[0]), int(w_hat.shape[1]), 32
        d.nx, d.ny = int(nx), int(ny)
        for i in range(3):
            d.voxel_size[i], d.offset[i] = float(voxel_size[i]), float(offset[i])
        d.has_pre_quant = 0 if pre_quant is None else 1
        if pre_quant is not None:
            d.pre_delta, d.pre_zero_point, d.pre_bits = float(pre_quant[0]), float(pre_quant[1]), int(pre_quant[2])
        else:
            d.pre_delta, d.pre_zero_point, d.pre_bits = 1.0, 0.0, 8
        d.out_delta, d.out_zero_point, d.out_bits = float(out_quant[0]), float(out_quant[1]), int(out_quant[2])
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
        self.nx, self.ny, self.cout = d.nx, d.ny, d.cout
        self.out_delta = float(out_quant[0])
        self._h = c_void_p()
        check(_lib.lib().qv2x_pillar_create(byref(d), _np_ptr(w_hat), None if b is None else _np_ptr(b),
                                            byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().qv2x_pillar_destroy(h)
            except Exception:
                pass
            self._h = None

    def forward(self, voxel_features: torch.Tensor, voxel_coords: torch.Tensor, voxel_num_points: torch.Tensor,
                batch: int, out: torch.Tensor | None = None, rowsum_out: torch.Tensor | None = None,
                assume_zero: bool = False) -> torch.Tensor:
        """voxel_features float32 [M, 32, 4], voxel_coords int [M, 4] (batch, z, y, x), voxel_num_points int [M]
        (CUDA tensors) -> uint8 BEV codes [batch, ny, nx, 64].  rowsum_out int32 [batch, ny, nx]: also receives the
        per-cell sum of the 64 codes (what the backbone plan's first conv needs, Plan.forward(..., rowsum_in=)).
        assume_zero: `out` (and `rowsum_out`) are all-zero already -- skip the clear (see clear())."""
        assert voxel_features.is_cuda and voxel_features.dim() == 3 and voxel_features.shape[1:] == (32, 4)
        f = voxel_features.to(torch.float32).contiguous()
        c = voxel_coords.to(torch.int32).contiguous()
        n = voxel_num_points.to(torch.int32).contiguous()
        if out is None:
            out = torch.empty((batch, self.ny, self.nx, self.cout), dtype=torch.uint8, device=f.device)
        assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous()
        assert tuple(out.shape) == (batch, self.ny, self.nx, self.cout)
        if rowsum_out is not None:
            assert rowsum_out.is_cuda and rowsum_out.dtype == torch.int32 and rowsum_out.is_contiguous()
            assert tuple(rowsum_out.shape) == (batch, self.ny, self.nx)
        fn = _lib.lib().qv2x_pillar_scatter if assume_zero else _lib.lib().qv2x_pillar_forward_rs
        check(fn(self._h, int(f.shape[0]), c_void_p(f.data_ptr()), c_void_p(c.data_ptr()), c_void_p(n.data_ptr()),
                 int(batch), c_void_p(out.data_ptr()),
                 None if rowsum_out is None else c_void_p(rowsum_out.data_ptr()), _stream_ptr()))
        return out

    def clear(self, voxel_coords: torch.Tensor, batch: int, bev: torch.Tensor,
              rowsum: torch.Tensor | None = None) -> None:
        """Zero the cells (and row sums) of these pillars in a map forward() filled: after its consumer has run, the
        map is all-zero again and the next forward(..., assume_zero=True) needs no 9 MB-per-agent clear."""
        assert voxel_coords.is_cuda and voxel_coords.dtype == torch.int32 and voxel_coords.is_contiguous()
        assert bev.is_cuda and bev.dtype == torch.uint8 and tuple(bev.shape) == (batch, self.ny, self.nx, self.cout)
        check(_lib.lib().qv2x_pillar_clear(self._h, int(voxel_coords.shape[0]), c_void_p(voxel_coords.data_ptr()),
                                           int(batch), c_void_p(bev.data_ptr()),
                                           None if rowsum is None else c_void_p(rowsum.data_ptr()), _stream_ptr()))


class PostProcessEngine:
    """libqv2x handle of the GPU detection post-processing (qv2x_postprocess_*).

    anchor_cfg: the yaml's anchor generator config list (one dict per class: anchor_sizes [[l, w, h]],
    anchor_rotations, anchor_bottom_heights, align_center, feature_map_stride), as
    VoxelPostprocessor3Heads.generate_anchor_box reads it; lidar_range / grid_wh as in the yaml."""

    def __init__(self, anchor_cfg, lidar_range, grid_wh, *, score_threshold: float, nms_threshold: float,
                 box_range, max_candidates: int = 8192, top: int = 1000):
        from ._lib import PostprocessDesc

        d = PostprocessDesc()
        strides = {c["feature_map_stride"] for c in anchor_cfg}
        rots = [tuple(c["anchor_rotations"]) for c in anchor_cfg]
        if len(strides) != 1 or len(set(rots)) != 1:
            raise NotImplementedError("all classes must share the feature-map stride and the rotation list")
        st = strides.pop()
        d.W, d.H = int(grid_wh[0] // st), int(grid_wh[1] // st)
        d.n_classes, d.n_rotations = len(anchor_cfg), len(rots[0])
        for c, cfg in enumerate(anchor_cfg):
            if len(cfg["anchor_sizes"]) != 1 or len(cfg["anchor_bottom_heights"]) != 1:
                raise NotImplementedError("one anchor size and one bottom height per class")
            if cfg["align_center"]:
                xs, ys = (lidar_range[3] - lidar_range[0]) / d.W, (lidar_range[4] - lidar_range[1]) / d.H
                xo, yo = xs / 2, ys / 2
            else:
                xs, ys = (lidar_range[3] - lidar_range[0]) / (d.W - 1), (lidar_range[4] - lidar_range[1]) / (d.H - 1)
                xo, yo = 0.0, 0.0
            d.anchor_x0[c], d.anchor_y0[c] = lidar_range[0] + xo, lidar_range[1] + yo
            d.anchor_dx[c], d.anchor_dy[c] = xs, ys
            d.anchor_z[c] = cfg["anchor_bottom_heights"][0]
            l, w, h = cfg["anchor_sizes"][0]
            d.anchor_hwl[c][0], d.anchor_hwl[c][1], d.anchor_hwl[c][2] = h, w, l
        for r, v in enumerate(rots[0]):
            d.anchor_rot[r] = v
        d.score_threshold, d.nms_threshold = float(score_threshold), float(nms_threshold)
        d.range_lo[0], d.range_lo[1] = box_range[0], box_range[1]
        d.range_hi[0], d.range_hi[1] = box_range[3], box_range[4]
        d.max_candidates, d.top = int(max_candidates), int(top)
        self.top, self.hw = int(top), d.H * d.W
        self.channels = d.n_classes * d.n_classes * d.n_rotations + 7 * d.n_classes * d.n_rotations
        self._h = c_void_p()
        check(_lib.lib().qv2x_postprocess_create(byref(d), byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().qv2x_postprocess_destroy(h)
            except Exception:
                pass
            self._h = None

    def alloc_outputs(self, device):
        """Static output buffers for forward_into: (corners [top, 4, 2] f64, scores [top] f64, labels [top] i32,
        boxes [top, 7] f64, counts [2] i32 = boxes kept / candidates above the threshold)."""
        return (torch.empty((self.top, 4, 2), dtype=torch.float64, device=device),
                torch.empty((self.top,), dtype=torch.float64, device=device),
                torch.empty((self.top,), dtype=torch.int32, device=device),
                torch.empty((self.top, 7), dtype=torch.float64, device=device),
                torch.zeros((2,), dtype=torch.int32, device=device))

    def forward_into(self, preds: torch.Tensor, outs) -> None:
        """Asynchronous form (no host synchronisation, capturable in a CUDA graph): results go to `outs` from
        alloc_outputs(); outs[4][0] holds the number of boxes."""
        assert preds.is_cuda and preds.dtype == torch.float32 and preds.is_contiguous()
        p = preds if preds.dim() == 2 else preds.reshape(-1, self.hw)
        assert p.shape[1] == self.hw and p.shape[0] >= self.channels
        corners, scores, labels, boxes, n = outs
        check(_lib.lib().qv2x_postprocess_forward(self._h, c_void_p(p.data_ptr()), c_void_p(corners.data_ptr()),
                                                  c_void_p(scores.data_ptr()), c_void_p(labels.data_ptr()),
                                                  c_void_p(boxes.data_ptr()), c_void_p(n.data_ptr()),
                                                  c_void_p(n.data_ptr() + 4), _stream_ptr()))

    def forward(self, preds: torch.Tensor):
        """preds float32 [>= cls + reg channels, H*W] (channel-major head maps, cls first then reg) ->
        (corners [K, 4, 2], scores [K], labels [K], boxes [K, 7]) float64 / int32 CUDA tensors in NMS pick order."""
        assert preds.is_cuda and preds.dtype == torch.float32 and preds.is_contiguous()
        p = preds.reshape(preds.shape[0] if preds.dim() == 2 else -1, self.hw) if preds.dim() != 2 else preds
        assert p.shape[1] == self.hw and p.shape[0] >= self.channels
        dev = preds.device
        corners = torch.empty((self.top, 4, 2), dtype=torch.float64, device=dev)
        scores = torch.empty((self.top,), dtype=torch.float64, device=dev)
        labels = torch.empty((self.top,), dtype=torch.int32, device=dev)
        boxes = torch.empty((self.top, 7), dtype=torch.float64, device=dev)
        n = torch.zeros((2,), dtype=torch.int32, device=dev)
        check(_lib.lib().qv2x_postprocess_forward(self._h, c_void_p(p.data_ptr()), c_void_p(corners.data_ptr()),
                                                  c_void_p(scores.data_ptr()), c_void_p(labels.data_ptr()),
                                                  c_void_p(boxes.data_ptr()), c_void_p(n.data_ptr()),
                                                  c_void_p(n.data_ptr() + 4), _stream_ptr()))
        k = int(n[0].item())
        self.last_candidates = int(n[1].item())
        return corners[:k], scores[:k], labels[:k], boxes[:k]


def heads_forward_tile(heads: "HeadsEngine", x: torch.Tensor, out_ptr: int, tile_w: int, out_w: int, out_pixels: int):
    """Heads on a compact tile x float32 [tile_pixels, Cin]; output o of tile pixel (ty, tx) goes to
    out_ptr[o * out_pixels + ty * out_w + tx] (out_ptr: raw device address, possibly in a peer GPU's memory)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[-1] == heads.cin
    pixels = x.numel() // heads.cin
    check(_lib.lib().qv2x_heads_forward_tile(heads._h, pixels, c_void_p(x.data_ptr()), c_void_p(out_ptr), tile_w, out_w,
                                             out_pixels, _stream_ptr()))


def push_planes(codes: torch.Tensor, dst_plane_stride: int, dst_row0: int, peer_ptrs):
    """Store codes uint8 [levels, m, rows_local] into every peer's code buffer (see qv2x_push_planes)."""
    assert codes.is_cuda and codes.dtype == torch.uint8 and codes.is_contiguous()
    planes, rows_local = codes.shape[0] * codes.shape[1], codes.shape[2]
    arr = (c_void_p * len(peer_ptrs))(*[c_void_p(int(p)) for p in peer_ptrs])
    check(_lib.lib().qv2x_push_planes(c_void_p(codes.data_ptr()), planes, rows_local, dst_plane_stride, dst_row0, arr,
                                      len(peer_ptrs), _stream_ptr()))


def scatter_planes(codes: torch.Tensor, dst_plane_stride: int, dst_row0: int, peer_ptrs):
    """All-to-all of code planes: codes uint8 [levels, m, n_peers * rows] (frame-major rows); peer p receives the rows
    of frame p (see qv2x_scatter_planes)."""
    assert codes.is_cuda and codes.dtype == torch.uint8 and codes.is_contiguous()
    planes, rows_all = codes.shape[0] * codes.shape[1], codes.shape[2]
    assert rows_all % len(peer_ptrs) == 0
    arr = (c_void_p * len(peer_ptrs))(*[c_void_p(int(p)) for p in peer_ptrs])
    check(_lib.lib().qv2x_scatter_planes(c_void_p(codes.data_ptr()), planes, rows_all // len(peer_ptrs),
                                         dst_plane_stride, dst_row0, arr, len(peer_ptrs), _stream_ptr()))


def quantize_nchw_to_nhwc_u8(x: torch.Tensor, delta: float, zero_point: float = 0.0, bits: int = 8,
                             out: torch.Tensor | None = None, out_cbase: int = 0) -> torch.Tensor:
    """float32 NCHW -> uint8 NHWC activation codes (module-boundary converter)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    n, c, h, w = x.shape
    if out is None:
        out = torch.empty((n, h, w, c), dtype=torch.uint8, device=x.device)
    check(_lib.lib().qv2x_quantize_nchw_to_nhwc_u8(c_void_p(x.data_ptr()), n, c, h * w, float(delta),
                                                   float(zero_point), bits, c_void_p(out.data_ptr()), out.shape[3],
                                                   out_cbase, _stream_ptr()))
    return out


def dequant_nhwc_u8_to_nchw_f32(x: torch.Tensor, delta: float, zero_point: float = 0.0) -> torch.Tensor:
    assert x.is_cuda and x.dtype == torch.uint8 and x.is_contiguous() and x.dim() == 4
    n, h, w, c = x.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    check(_lib.lib().qv2x_dequant_nhwc_u8_to_nchw_f32(c_void_p(x.data_ptr()), n, c, h * w, float(delta),
                                                      float(zero_point), c_void_p(out.data_ptr()), _stream_ptr()))
    return out


def dequantize_u8(x: torch.Tensor, delta: float, out: torch.Tensor | None = None) -> torch.Tensor:
    """uint8 codes -> float32 fl(delta * code), same shape / layout."""
    assert x.is_cuda and x.dtype == torch.uint8 and x.is_contiguous() and x.numel() % 16 == 0
    if out is None:
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    check(_lib.lib().qv2x_dequant_u8(c_void_p(x.data_ptr()), x.numel(), float(delta), c_void_p(out.data_ptr()),
                                     _stream_ptr()))
    return out


def nchw_to_nhwc_f32(x: torch.Tensor) -> torch.Tensor:
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    n, c, h, w = x.shape
    out = torch.empty((n, h, w, c), dtype=torch.float32, device=x.device)
    check(_lib.lib().qv2x_nchw_to_nhwc_f32(c_void_p(x.data_ptr()), n, c, h * w, c_void_p(out.data_ptr()),
                                           _stream_ptr()))
    return out


def nhwc_to_nchw_f32(x: torch.Tensor) -> torch.Tensor:
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    n, h, w, c = x.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    check(_lib.lib().qv2x_nhwc_to_nchw_f32(c_void_p(x.data_ptr()), n, c, h * w, c_void_p(out.data_ptr()),
                                           _stream_ptr()))
    return out


class Plan:
    """A fixed launch sequence of QLayers over numbered buffers (qv2x_plan): one modality's quantized
    BaseBEVBackbone + DownsampleConv.  steps: list of (QLayer, in_buf, in_cbase, out_buf, out_cbase)."""

    def __init__(self, steps, buf_channels):
        from ._lib import PlanStep

        self.layers = [s[0] for s in steps]          # keep the layers alive
        self.wiring = [(int(ib), int(ic), int(ob), int(oc)) for (_, ib, ic, ob, oc) in steps]
        arr = (PlanStep * len(steps))()
        for i, (layer, ib, ic, ob, oc) in enumerate(steps):
            arr[i].layer, arr[i].in_buf, arr[i].in_cbase, arr[i].out_buf, arr[i].out_cbase = layer._h, ib, ic, ob, oc
        bc = (c_int * len(buf_channels))(*buf_channels)
        self.buf_channels = list(buf_channels)
        self._h = c_void_p()
        check(_lib.lib().qv2x_plan_create(arr, len(steps), bc, len(buf_channels), byref(self._h)))
        self._ws = {}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().qv2x_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    def out_shape(self, h, w):
        ho, wo, c = c_int(), c_int(), c_int()
        check(_lib.lib().qv2x_plan_out_shape(self._h, h, w, byref(ho), byref(wo), byref(c)))
        return ho.value, wo.value, c.value

    def workspace(self, n, h, w, device, slot=0):
        """Scratch activations of one forward; `slot` separates frames that are in flight concurrently."""
        key = (n, h, w, str(device), slot)
        if key not in self._ws:
            nbytes = ctypes.c_size_t()
            check(_lib.lib().qv2x_plan_workspace_bytes(self._h, n, h, w, byref(nbytes)))
            self._ws[key] = torch.empty(nbytes.value, dtype=torch.uint8, device=device)
        return self._ws[key]

    def forward(self, x: torch.Tensor, out: torch.Tensor | None = None, dump_step: int = -1, acc_dump=None, slot=0,
                rowsum_in: torch.Tensor | None = None):
        """x uint8 NHWC [n, H, W, C0] -> uint8 NHWC [n, ho, wo, C_last].  rowsum_in: int32 [n, H, W] per-pixel channel
        sums of x when its producer already has them (PillarEngine.forward(..., rowsum_out=))."""
        assert x.is_cuda and x.dtype == torch.uint8 and x.is_contiguous() and x.dim() == 4
        n, h, w, c = x.shape
        assert c == self.buf_channels[0], f"input has {c} channels, plan expects {self.buf_channels[0]}"
        ho, wo, co = self.out_shape(h, w)
        if out is None:
            out = torch.empty((n, ho, wo, co), dtype=torch.uint8, device=x.device)
        ws = self.workspace(n, h, w, x.device, slot)
        if rowsum_in is not None:
            assert rowsum_in.is_cuda and rowsum_in.dtype == torch.int32 and rowsum_in.is_contiguous()
            assert tuple(rowsum_in.shape) == (n, h, w)
        check(_lib.lib().qv2x_plan_forward_rs(self._h, n, h, w, c_void_p(x.data_ptr()),
                                              None if rowsum_in is None else c_void_p(rowsum_in.data_ptr()),
                                              c_void_p(out.data_ptr()), c_void_p(ws.data_ptr()), ws.numel(), dump_step,
                                              None if acc_dump is None else c_void_p(acc_dump.data_ptr()),
                                              _stream_ptr()))
        return out
