"""Serialized PTQ engine and wire format.

* ``save_pipeline / load_pipeline``: everything a calibrated ``CollabPipeline`` needs -- integer weight grids,
  quantizer parameters, plan wiring, codebook + heads parameters, PointPillars front end -- in ONE ``.npz`` file, so a
  deployment loads the engines without the float model, the yaml or a calibration pass.  (The reference cannot
  persist PTQ results at all: tools/inference_mc_quant.py recalibrates on every run.)
* ``pack_codes / unpack_codes``: the message an agent sends -- a 24-byte header and the code planes packed to
  ceil(log2 k) bits per code (7 bits at k = 128: 92.4 KB per agent at 100 x 352, instead of the reference's pickled
  int64 lists, ~845 KB, tools/inference_mc_codebook_encdec_cached.py:117-134).
Host-side code: numpy only (the engines themselves are built by quantv2x_b200.engine on load).
"""
from __future__ import annotations

import json
import struct

import numpy as np

FORMAT_VERSION = 1
_MAGIC = b"QV2X"
# the six nn.Linear heads of one codebook level (reference codebook.py _components order)
HEAD_NAMES = ("latentStageEncoder", "quantizationHead", "latentHead", "dequantizationHead", "sideHead", "restoreHead")


# ------------------------------------------------------------------------------------------- engine file
def _put(store: dict, prefix: str, spec: dict, meta: dict):
    for k, v in spec.items():
        if isinstance(v, np.ndarray):
            store[f"{prefix}.{k}"] = v
        else:
            meta[f"{prefix}.{k}"] = v


def pipeline_state(pipe) -> tuple[dict, dict]:
    """(arrays, meta) of a CollabPipeline: numpy arrays keyed by name + a JSON-able dict."""
    arrays, meta = {}, {"format": FORMAT_VERSION}
    plan = pipe.fused.plan
    meta["plan.buf_channels"] = list(plan.buf_channels)
    meta["plan.wiring"] = [list(w) for w in plan.wiring]
    meta["plan.n_layers"] = len(plan.layers)
    for i, layer in enumerate(plan.layers):
        _put(arrays, f"layer{i}", {k: v for k, v in layer.spec.items() if v is not None}, meta)
        meta[f"layer{i}.has_bias"] = layer.spec["bias"] is not None
    cb = pipe.codebook.spec
    meta["codebook.levels"] = len(cb["codebooks"])
    for l, c in enumerate(cb["codebooks"]):
        arrays[f"codebook.{l}"] = c
        for name, wb in cb["heads"][l].items():
            if wb is not None:
                arrays[f"codebook.{l}.{name}.w"] = np.asarray(wb[0], np.float32)
                arrays[f"codebook.{l}.{name}.b"] = np.asarray(wb[1], np.float32)
    arrays["heads.w"] = pipe.heads.spec["w"]
    if pipe.heads.spec["bias"] is not None:
        arrays["heads.bias"] = pipe.heads.spec["bias"]
    meta.update({"feat_delta": pipe.feat_delta, "fusion_mode": pipe.fusion_mode, "bev_hw": [pipe.H, pipe.W],
                 "bev_delta": getattr(pipe, "bev_delta", None),
                 "in_deltas": list(pipe.fused.in_deltas), "in_group_channels": list(pipe.fused.in_group_channels),
                 "out_deltas": list(pipe.fused.out_deltas), "out_group_channels": list(pipe.fused.out_group_channels)})
    pil = getattr(pipe, "pillar_engine", None)
    meta["pillar"] = pil is not None
    if pil is not None:
        sp = pil.spec
        arrays["pillar.w_hat"] = sp["w_hat"]
        if sp["bias"] is not None:
            arrays["pillar.bias"] = sp["bias"]
        meta["pillar.spec"] = {k: (list(v) if isinstance(v, tuple) else v) for k, v in sp.items()
                               if k not in ("w_hat", "bias")}
    return arrays, meta


def _npz_path(path: str) -> str:
    """np.savez appends '.npz' to other names: save and load agree on the same normalised path."""
    return path if str(path).endswith(".npz") else str(path) + ".npz"


def save_pipeline(pipe, path: str) -> None:
    arrays, meta = pipeline_state(pipe)
    np.savez(_npz_path(path), __meta__=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)


def load_pipeline(path: str, device):
    """Rebuild the libqv2x engines of a saved pipeline on `device` (a CUDA device; there is no CPU path)."""
    from . import engine as E
    from .export import BlockEngine
    from .pipeline import CollabPipeline

    z = np.load(_npz_path(path))
    meta = json.loads(bytes(z["__meta__"]).decode())
    if meta.get("format") != FORMAT_VERSION:
        raise ValueError(f"unsupported engine file format {meta.get('format')}")
    layers = []
    for i in range(meta["plan.n_layers"]):
        g = lambda k: z[f"layer{i}.{k}"]
        m = lambda k: meta[f"layer{i}.{k}"]
        layers.append(E.QLayer(kind=m("kind"), w_int=g("w_int"), w_delta=g("w_delta"), w_zp=g("w_zp"),
                               bias=g("bias") if m("has_bias") else None, ksize=m("ksize"), stride=m("stride"),
                               pad=m("pad"), w_bits=m("w_bits"), relu=m("relu"), in_delta=g("in_delta"),
                               out_delta=m("out_delta"), out_zp=m("out_zp"), out_bits=m("out_bits"),
                               groups=meta.get(f"layer{i}.groups", 1)))
    steps = [(layers[i],) + tuple(w) for i, w in enumerate(meta["plan.wiring"])]
    plan = E.Plan(steps, meta["plan.buf_channels"])
    fused = BlockEngine(plan, meta["in_deltas"], meta["in_group_channels"], meta["out_deltas"],
                        meta["out_group_channels"])
    cbs, heads = [], []
    for l in range(meta["codebook.levels"]):
        cbs.append(z[f"codebook.{l}"])
        heads.append({name: ((z[f"codebook.{l}.{name}.w"], z[f"codebook.{l}.{name}.b"])
                             if f"codebook.{l}.{name}.w" in z.files else None) for name in HEAD_NAMES})
    pipe = CollabPipeline(fused, meta["feat_delta"], E.CodebookEngine(cbs, heads),
                          E.HeadsEngine(z["heads.w"], z["heads.bias"] if "heads.bias" in z.files else None),
                          meta["fusion_mode"], tuple(meta["bev_hw"]), device)
    pipe.bev_delta = meta["bev_delta"]
    if meta["pillar"]:
        sp = meta["pillar.spec"]
        pq = None if sp["pre_quant"] is None else tuple(sp["pre_quant"])
        pipe.pillar_engine = E.PillarEngine(z["pillar.w_hat"], z["pillar.bias"] if "pillar.bias" in z.files else None,
                                            nx=sp["nx"], ny=sp["ny"], voxel_size=sp["voxel_size"], offset=sp["offset"],
                                            pre_quant=pq, out_quant=tuple(sp["out_quant"]))
    return pipe


# ------------------------------------------------------------------------------------------- wire format
def _bits(k: int) -> int:
    return max(1, int(np.ceil(np.log2(k))))


def pack_codes(codes: np.ndarray, k: int) -> bytes:
    """codes uint8 [levels, m, rows] with values < k  ->  header + planes packed to ceil(log2 k) bits per code.

    header (24 bytes, little endian): magic 'QV2X', version u16, bits u16, levels u16, m u16, k u32, rows u64."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    levels, m, rows = codes.shape
    if not 2 <= k <= 256:
        raise ValueError(f"dictionary size {k} outside 2..256 (codes travel as bytes)")
    if codes.size and int(codes.max()) >= k:
        raise ValueError("code out of range")
    b = _bits(k)
    head = _MAGIC + struct.pack("<HHHHIQ", FORMAT_VERSION, b, levels, m, k, rows)
    if b == 8:
        return head + codes.tobytes()
    bits = np.unpackbits(codes.reshape(-1, 1), axis=1)[:, 8 - b:]          # [n, b], most significant bit first
    return head + np.packbits(bits.reshape(-1)).tobytes()


def unpack_codes(msg: bytes) -> tuple[np.ndarray, int]:
    """Inverse of pack_codes: returns (codes uint8 [levels, m, rows], k)."""
    if msg[:4] != _MAGIC:
        raise ValueError("not a qv2x code message")
    ver, b, levels, m, k, rows = struct.unpack("<HHHHIQ", msg[4:24])
    if ver != FORMAT_VERSION:
        raise ValueError(f"unsupported wire format {ver}")
    # the message comes from another agent: never trust its header
    if not 2 <= k <= 256 or not 1 <= b <= 8 or b != _bits(k):
        raise ValueError(f"inconsistent header: k={k}, {b} bits per code")
    if levels < 1 or m < 1:
        raise ValueError("inconsistent header: no planes")
    n = levels * m * rows
    body = np.frombuffer(msg, dtype=np.uint8, offset=24)
    if b == 8:
        if body.size != n:
            raise ValueError("truncated message")
        out = body.reshape(levels, m, rows).copy()
        if out.size and int(out.max()) >= k:
            raise ValueError("code out of range in message")
        return out, k
    if body.size != (n * b + 7) // 8:
        raise ValueError("truncated message")
    bits = np.unpackbits(body)[:n * b].reshape(n, b)
    full = np.zeros((n, 8), np.uint8)
    full[:, 8 - b:] = bits
    out = np.packbits(full, axis=1).reshape(levels, m, rows)
    if out.size and int(out.max()) >= k:
        raise ValueError("code out of range in message")
    return out, k
