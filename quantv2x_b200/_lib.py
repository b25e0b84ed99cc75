"""ctypes binding of libqv2x.so (C ABI declared in include/qv2x.h).

The library is the product: there is no Python/CPU fallback.  Importing this module never needs a GPU
(the shared object links the CUDA runtime statically and resolves driver symbols lazily), but every
compute entry point fails loudly if the library or an sm_100 device is missing.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_longlong, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# QV2X_LIB selects another build of the same library (the bring-up build with role traces, `make debug`)
LIB_PATH = os.environ.get("QV2X_LIB") or os.path.join(_HERE, "libqv2x.so")


class Qv2xError(RuntimeError):
    pass


class _SizedStructure(ctypes.Structure):
    """Descriptors start with `struct_size` = sizeof(struct): the library rejects a mirror whose layout differs."""

    def __init__(self, *args, **kw):
        super().__init__(*args, **kw)
        self.struct_size = ctypes.sizeof(type(self))


class LayerDesc(_SizedStructure):
    """Mirror of qv2x_layer_desc (include/qv2x.h)."""

    _fields_ = [
        ("struct_size", c_uint32),
        ("kind", c_int),
        ("cin", c_int),
        ("cout", c_int),
        ("ksize", c_int),
        ("stride", c_int),
        ("pad", c_int),
        ("w_bits", c_int),
        ("relu", c_int),
        ("n_in_groups", c_int),
        ("in_delta", c_float * 3),
        ("out_delta", c_float),
        ("out_zero_point", c_float),
        ("out_bits", c_int),
        ("groups", c_int),
    ]


class LayerExtra(_SizedStructure):
    """Mirror of qv2x_layer_extra (include/qv2x.h): shortcut input and FP32 output of residual-block convs."""

    _fields_ = [
        ("struct_size", c_uint32),
        ("d_res_u8", c_void_p),
        ("d_res_f32", c_void_p),
        ("res_delta", c_float),
        ("res_cstride", c_int),
        ("res_cbase", c_int),
        ("d_out_f32", c_void_p),
        ("out_f32_cstride", c_int),
    ]


_lib = None


def _declare(lib):
    lib.qv2x_last_error.restype = c_char_p
    lib.qv2x_last_error.argtypes = []
    lib.qv2x_version.restype = c_int
    lib.qv2x_device_check.argtypes = [c_int]
    lib.qv2x_launch_count.restype = c_longlong
    lib.qv2x_layer_create.argtypes = [POINTER(LayerDesc), c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_void_p)]
    lib.qv2x_layer_destroy.argtypes = [c_void_p]
    lib.qv2x_layer_destroy.restype = None
    lib.qv2x_layer_needs_rowsum.argtypes = [c_void_p]
    lib.qv2x_layer_out_shape.argtypes = [c_void_p, c_int, c_int, POINTER(c_int), POINTER(c_int)]
    lib.qv2x_layer_forward.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, POINTER(c_void_p),
                                       c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.qv2x_rowsum_u8.argtypes = [c_void_p, c_longlong, c_int, c_int, c_int, c_void_p, c_void_p]


def lib():
    """Load libqv2x.so once; raise if it has not been built (python __graft_entry__.py / make -C csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Qv2xError(f"{LIB_PATH} is missing: build it with `make -C quantv2x_b200/csrc` "
                            "(there is no CPU fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for fn in _DECLARERS:
            fn(handle)
        _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        raise Qv2xError(f"libqv2x error {rc}: {lib().qv2x_last_error().decode()}")


def exported_symbols():
    """Names declared in include/qv2x.h (used by the CPU test that the library exports all of them)."""
    import re

    hdr = os.path.join(os.path.dirname(_HERE), "include", "qv2x.h")
    text = open(hdr).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qv2x_[a-z0-9_]+)\s*\(", text)))


class PillarDesc(_SizedStructure):
    """Mirror of qv2x_pillar_desc (include/qv2x.h)."""

    _fields_ = [("struct_size", c_uint32), ("n_feat", c_int), ("cout", c_int), ("max_points", c_int), ("nx", c_int), ("ny", c_int),
                ("voxel_size", c_float * 3), ("offset", c_float * 3), ("has_pre_quant", c_int),
                ("pre_delta", c_float), ("pre_zero_point", c_float), ("pre_bits", c_int),
                ("out_delta", c_float), ("out_zero_point", c_float), ("out_bits", c_int)]


def _declare_pillar(lib):
    lib.qv2x_pillar_create.argtypes = [POINTER(PillarDesc), c_void_p, c_void_p, POINTER(c_void_p)]
    lib.qv2x_pillar_destroy.argtypes = [c_void_p]
    lib.qv2x_pillar_destroy.restype = None
    lib.qv2x_pillar_forward.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    lib.qv2x_pillar_forward_rs.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                           c_void_p]
    lib.qv2x_pillar_scatter.argtypes = lib.qv2x_pillar_forward_rs.argtypes
    lib.qv2x_pillar_clear.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]


class PostprocessDesc(_SizedStructure):
    """Mirror of qv2x_postprocess_desc (include/qv2x.h)."""

    _fields_ = [("struct_size", c_uint32), ("H", c_int), ("W", c_int), ("n_classes", c_int), ("n_rotations", c_int),
                ("anchor_x0", ctypes.c_double * 4), ("anchor_y0", ctypes.c_double * 4),
                ("anchor_dx", ctypes.c_double * 4), ("anchor_dy", ctypes.c_double * 4),
                ("anchor_z", ctypes.c_double * 4), ("anchor_hwl", (ctypes.c_double * 3) * 4),
                ("anchor_rot", ctypes.c_double * 4), ("score_threshold", ctypes.c_double),
                ("nms_threshold", c_float), ("range_lo", ctypes.c_double * 2), ("range_hi", ctypes.c_double * 2),
                ("max_candidates", c_int), ("top", c_int)]


def _declare_postprocess(lib):
    lib.qv2x_postprocess_create.argtypes = [POINTER(PostprocessDesc), POINTER(c_void_p)]
    lib.qv2x_postprocess_destroy.argtypes = [c_void_p]
    lib.qv2x_postprocess_destroy.restype = None
    lib.qv2x_postprocess_forward.argtypes = [c_void_p] + [c_void_p] * 8


class CodebookDesc(_SizedStructure):
    """Mirror of qv2x_codebook_desc (include/qv2x.h)."""

    _fields_ = [("struct_size", c_uint32), ("channel", c_int), ("m", c_int), ("levels", c_int), ("k", c_int * 4)]


def _declare_codebook(lib):
    lib.qv2x_codebook_create.argtypes = [POINTER(CodebookDesc), POINTER(c_void_p), POINTER(c_void_p),
                                         POINTER(c_void_p), POINTER(c_void_p)]
    lib.qv2x_codebook_destroy.argtypes = [c_void_p]
    lib.qv2x_codebook_destroy.restype = None
    lib.qv2x_codebook_encode.argtypes = [c_void_p, c_longlong, c_void_p, c_int, c_float, c_void_p, c_void_p]
    lib.qv2x_codebook_decode.argtypes = [c_void_p, c_longlong, c_void_p, c_void_p, c_void_p]
    lib.qv2x_codebook_folded_size.argtypes = [c_void_p, c_int]
    lib.qv2x_codebook_folded_size.restype = c_longlong
    lib.qv2x_codebook_folded_copy.argtypes = [c_void_p, c_int, c_void_p]


_DECLARERS = [_declare, _declare_codebook]


def _declare_fusion(lib):
    lib.qv2x_fuse.argtypes = [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.qv2x_heads_create.argtypes = [c_int, c_int, c_void_p, c_void_p, POINTER(c_void_p)]
    lib.qv2x_heads_destroy.argtypes = [c_void_p]
    lib.qv2x_heads_destroy.restype = None
    lib.qv2x_heads_forward.argtypes = [c_void_p, c_longlong, c_void_p, c_void_p, c_void_p]
    lib.qv2x_quantize_nchw_to_nhwc_u8.argtypes = [c_void_p, c_int, c_int, c_longlong, c_float, c_float, c_int,
                                                  c_void_p, c_int, c_int, c_void_p]
    lib.qv2x_dequant_nhwc_u8_to_nchw_f32.argtypes = [c_void_p, c_int, c_int, c_longlong, c_float, c_float, c_void_p,
                                                     c_void_p]
    lib.qv2x_nchw_to_nhwc_f32.argtypes = [c_void_p, c_int, c_int, c_longlong, c_void_p, c_void_p]
    lib.qv2x_nhwc_to_nchw_f32.argtypes = [c_void_p, c_int, c_int, c_longlong, c_void_p, c_void_p]


_DECLARERS.append(_declare_fusion)


def _declare_ego_att(lib):
    lib.qv2x_codebook_desc_get.argtypes = [c_void_p, POINTER(CodebookDesc)]
    lib.qv2x_ego_att_supported.argtypes = [c_void_p, c_int]
    lib.qv2x_ego_att_create.argtypes = [c_void_p, c_int, c_void_p, c_void_p, POINTER(c_void_p)]
    lib.qv2x_ego_att_destroy.argtypes = [c_void_p]
    lib.qv2x_ego_att_destroy.restype = None
    lib.qv2x_ego_att_forward.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_longlong, c_void_p, c_void_p,
                                         c_void_p]


_DECLARERS.append(_declare_ego_att)


def _declare_decode_linear(lib):
    lib.qv2x_decode_linear_supported.argtypes = [c_void_p, c_int]
    lib.qv2x_decode_linear_create.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_float, POINTER(c_void_p)]
    lib.qv2x_decode_linear_destroy.argtypes = [c_void_p]
    lib.qv2x_decode_linear_destroy.restype = None
    lib.qv2x_decode_linear_forward.argtypes = [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p]


_DECLARERS.append(_declare_decode_linear)


def _declare_fuse_u8(lib):
    lib.qv2x_fuse_weighted_u8.argtypes = [c_int, c_int, c_int, c_int, c_void_p, c_float, c_void_p, c_int, c_void_p,
                                          c_void_p, c_void_p]


    lib.qv2x_heads_forward_deconv_u8.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int, c_float, c_void_p, c_int,
                                                 c_int, c_void_p]


_DECLARERS.append(_declare_fuse_u8)


class PlanStep(ctypes.Structure):
    """Mirror of qv2x_plan_step (include/qv2x.h)."""

    _fields_ = [("layer", c_void_p), ("in_buf", c_int), ("in_cbase", c_int), ("out_buf", c_int), ("out_cbase", c_int)]


def _declare_peak(lib):
    lib.qv2x_int8_mma_peak.argtypes = [POINTER(ctypes.c_double), POINTER(ctypes.c_double), c_void_p]


_DECLARERS.append(_declare_peak)


def _declare_plan(lib):
    lib.qv2x_plan_create.argtypes = [POINTER(PlanStep), c_int, POINTER(c_int), c_int, POINTER(c_void_p)]
    lib.qv2x_plan_destroy.argtypes = [c_void_p]
    lib.qv2x_plan_destroy.restype = None
    lib.qv2x_plan_out_shape.argtypes = [c_void_p, c_int, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int)]
    lib.qv2x_plan_workspace_bytes.argtypes = [c_void_p, c_int, c_int, c_int, POINTER(ctypes.c_size_t)]
    lib.qv2x_plan_forward.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, ctypes.c_size_t,
                                      c_int, c_void_p, c_void_p]
    lib.qv2x_layer_desc_get.argtypes = [c_void_p, POINTER(LayerDesc)]
    lib.qv2x_plan_forward_rs.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                         ctypes.c_size_t, c_int, c_void_p, c_void_p]


_DECLARERS.append(_declare_plan)


def _declare_tiles(lib):
    lib.qv2x_fuse_tile.argtypes = [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                   c_int, c_int, c_void_p]
    lib.qv2x_codebook_decode_regions.argtypes = [c_void_p, c_longlong, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                                 c_void_p, c_void_p]
    lib.qv2x_heads_forward_tile.argtypes = [c_void_p, c_longlong, c_void_p, c_void_p, c_int, c_longlong, c_longlong,
                                            c_void_p]
    lib.qv2x_push_planes.argtypes = [c_void_p, c_int, c_longlong, c_longlong, c_longlong, c_void_p, c_int, c_void_p]
    lib.qv2x_scatter_planes.argtypes = [c_void_p, c_int, c_longlong, c_longlong, c_longlong, c_void_p, c_int, c_void_p]
    lib.qv2x_layer_forward_ex.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, POINTER(c_void_p),
                                          c_void_p, c_int, c_int, c_void_p, c_void_p, POINTER(LayerExtra), c_void_p]
    lib.qv2x_dequant_u8.argtypes = [c_void_p, c_longlong, c_float, c_void_p, c_void_p]
    lib.qv2x_fuse_weighted.argtypes = [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                       c_void_p]
    lib.qv2x_set_debug_flags.argtypes = [c_int]
    lib.qv2x_set_debug_flags.restype = None
    lib.qv2x_debug_trace.argtypes = [c_void_p]
    lib.qv2x_debug_trace.restype = None


_DECLARERS.append(_declare_tiles)
_DECLARERS.append(_declare_pillar)
_DECLARERS.append(_declare_postprocess)
