"""Frame-level pipeline of the quantized cooperative forward (BEV level).

    per agent GPU : BEV uint8 -> quantized backbone + shrinker (qv2x_plan) -> codebook encode -> byte planes
    exchange      : levels*m byte planes per agent (105.6 KB at m=1, H*W=35200) -> ego GPU
    ego GPU       : decode all agents -> warp + max/att fusion -> cls/reg/dir heads

Mirrors the data flow of the reference model forward (heter_baseline_collab_codebook_mc.py:71-169) with the
deterministic encode/decode split of heter_pyramid_collab_codebook_mc_encdec.py:33-208; every stage is a
libqv2x kernel sequence (no torch compute on the path).  Buffers are allocated once per agent count so the
whole frame can be captured in a CUDA graph.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import engine as E


class CollabPipeline:
    def __init__(self, fused_engine, feat_delta: float, codebook: E.CodebookEngine, heads: E.HeadsEngine,
                 fusion_mode: str, bev_hw, device):
        self.fused = fused_engine            # export.BlockEngine over backbone + shrinker
        self.feat_delta = float(feat_delta)  # activation scale of the shrinker output
        self.codebook = codebook
        self.heads = heads
        self.fusion_mode = fusion_mode
        self.device = device
        self.H, self.W = bev_hw
        self.ho, self.wo, self.c_feat = fused_engine.plan.out_shape(self.H, self.W)
        self.hw = self.ho * self.wo
        self._enc_buf = {}
        self._ego_buf = {}
        # attention fusion: decode, warp, fuse and heads fold into one kernel over the codeword tables (no feature
        # map in HBM); max fusion (not linear) and codebooks whose head table exceeds shared memory keep the
        # three-kernel chain.  QV2X_EGO_CHAIN=1 forces the chain (comparison runs).
        self.ego_att = None
        if (fusion_mode == "att" and os.environ.get("QV2X_EGO_CHAIN", "0") != "1"
                and E.EgoAttEngine.supported(codebook, heads)):
            self.ego_att = E.EgoAttEngine(codebook, heads)

    # ------------------------------------------------------------------ agent side
    def bev_from_inputs(self, data_dict, modality: str = "m1") -> torch.Tensor:
        """uint8 BEV codes [n, H, W, C] for the frame's agents.  BEV-level callers pass them directly as
        data_dict['inputs_<m>']['bev_u8']; pillar-level inputs go through the PFN + scatter kernel."""
        inp = data_dict[f"inputs_{modality}"]
        if "bev_u8" in inp:
            return inp["bev_u8"]
        pillar = getattr(self, "pillar_engine", None)
        if pillar is None:
            raise RuntimeError("pillar-level input needs the PFN + scatter engine: attach_engines() builds it when "
                               "the PointPillar encoder is quantized and calibrated; otherwise pass "
                               "inputs_m1['bev_u8'] (uint8 [n, H, W, 64])")
        n = len(data_dict["agent_modality_list"]) if "agent_modality_list" in data_dict else int(
            data_dict["record_len"].sum())
        dev = self.device
        return pillar.forward(inp["voxel_features"].to(dev), inp["voxel_coords"].to(dev),
                              inp["voxel_num_points"].to(dev), n)

    def encode_buffers(self, n, slot=0):
        """Buffers of the agent-side stage; `slot` separates frames that are in flight concurrently."""
        if (n, slot) not in self._enc_buf:
            d = self.device
            self._enc_buf[(n, slot)] = dict(
                feat=torch.empty((n, self.ho, self.wo, self.c_feat), dtype=torch.uint8, device=d),
                codes=torch.empty((self.codebook.levels, self.codebook.m, n * self.hw), dtype=torch.uint8, device=d))
        return self._enc_buf[(n, slot)]

    def encode_agents(self, bev_u8: torch.Tensor, slot=0, rowsum: torch.Tensor | None = None) -> torch.Tensor:
        """bev_u8 uint8 [n, H, W, C_bev] -> codes uint8 [levels, m, n*hw] (agent-major rows).  rowsum: the map's
        per-cell channel sums when its producer emitted them (encode_pillars)."""
        n = bev_u8.shape[0]
        b = self.encode_buffers(n, slot)
        self.fused.forward_u8(bev_u8, out=b["feat"], slot=slot, rowsum_in=rowsum)
        self.codebook.encode(b["feat"], self.feat_delta, out=b["codes"])
        return b["codes"]

    def encode_pillars(self, voxel_features, voxel_coords, voxel_num_points, n: int, slot=0,
                       bev_out: torch.Tensor | None = None) -> torch.Tensor:
        """The agent stage from the model's own input: pillars -> PointPillars front end (BEV codes + their per-cell
        sums in one kernel) -> backbone + shrinker plan -> codebook encode.  Returns codes uint8 [levels, m, n*hw].

        The stage's own BEV map and row sums are kept ALL-ZERO between calls: the front end only scatters, and once
        the plan has consumed the map the cells of this frame's pillars are zeroed again (3 MB instead of a 76 MB
        clear per 8-agent frame).  A caller-provided `bev_out` keeps its contents (full clear, as before)."""
        if getattr(self, "pillar_engine", None) is None:
            raise RuntimeError("no pillar engine attached (the PointPillar encoder is not quantized / calibrated)")
        pe = self.pillar_engine
        key = ("pillar", n, slot)
        if key not in self._enc_buf:
            self._enc_buf[key] = dict(
                bev=torch.zeros((n, pe.ny, pe.nx, pe.cout), dtype=torch.uint8, device=self.device),
                rowsum=torch.zeros((n, pe.ny, pe.nx), dtype=torch.int32, device=self.device))
        pb = self._enc_buf[key]
        if bev_out is not None:
            if "rowsum_ext" not in pb:
                pb["rowsum_ext"] = torch.empty((n, pe.ny, pe.nx), dtype=torch.int32, device=self.device)
            pe.forward(voxel_features, voxel_coords, voxel_num_points, n, out=bev_out, rowsum_out=pb["rowsum_ext"])
            return self.encode_agents(bev_out, slot, rowsum=pb["rowsum_ext"])
        coords = voxel_coords.to(torch.int32).contiguous()
        pe.forward(voxel_features, coords, voxel_num_points, n, out=pb["bev"], rowsum_out=pb["rowsum"],
                   assume_zero=True)
        codes = self.encode_agents(pb["bev"], slot, rowsum=pb["rowsum"])
        pe.clear(coords, n, pb["bev"], pb["rowsum"])
        return codes

    # ------------------------------------------------------------------ ego side
    def ego_buffers(self, n, slot=0):
        if (n, slot) not in self._ego_buf:
            d = self.device
            self._ego_buf[(n, slot)] = dict(
                feat=torch.empty((n, self.ho, self.wo, self.c_feat), dtype=torch.float32, device=d),
                fused=torch.empty((self.ho, self.wo, self.c_feat), dtype=torch.float32, device=d),
                preds=torch.empty((self.heads.cout, self.hw), dtype=torch.float32, device=d))
        return self._ego_buf[(n, slot)]

    def decode_fuse_heads(self, codes: torch.Tensor, affine: torch.Tensor, slot=0) -> torch.Tensor:
        """codes uint8 [levels, m, n*hw]; affine CUDA float32 [n, 2, 3] -> preds float32 [Cout, hw]."""
        n = codes.shape[-1] // self.hw
        if self.ego_att is not None:
            key = ("att", n, slot)
            if key not in self._ego_buf:
                self._ego_buf[key] = torch.empty((self.heads.cout, self.hw), dtype=torch.float32, device=self.device)
            aff = affine if (affine.dtype == torch.float32 and affine.is_contiguous()) else \
                affine.to(torch.float32).contiguous()
            return self.ego_att.forward(codes, aff, n, self.ho, self.wo, out=self._ego_buf[key])
        return self.decode_fuse_heads_chain(codes, affine, slot)

    def decode_fuse_heads_chain(self, codes: torch.Tensor, affine: torch.Tensor, slot=0) -> torch.Tensor:
        """The same stage as three kernels through decoded FP32 feature maps (decode -> warp + fuse -> heads)."""
        n = codes.shape[-1] // self.hw
        b = self.ego_buffers(n, slot)
        self.codebook.decode(codes, out=b["feat"].view(n * self.hw, self.c_feat))
        E.fuse(b["feat"], affine, self.fusion_mode, out=b["fused"])
        self.heads.forward(b["fused"], out=b["preds"])
        return b["preds"]

    def forward(self, bev_u8: torch.Tensor, affine: torch.Tensor) -> torch.Tensor:
        return self.decode_fuse_heads(self.encode_agents(bev_u8), affine)

    # ------------------------------------------------------------------ ego side, one output tile (multi-GPU)
    def source_rects(self, affine_host: np.ndarray, tile):
        """Per agent, the (y0, y1, x0, x1) source rectangle that the bilinear warp of output tile `tile` samples
        from (bounding box of the affinely mapped tile corners, padded by 2 pixels)."""
        y0, y1, x0, x1 = tile
        H, W = self.ho, self.wo
        rects = []
        for M in np.asarray(affine_host, dtype=np.float64).reshape(-1, 2, 3):
            xs, ys = [], []
            for i in (y0, y1 - 1):
                for j in (x0, x1 - 1):
                    xn, yn = (2.0 * j + 1.0) / W - 1.0, (2.0 * i + 1.0) / H - 1.0
                    xs.append(((M[0, 0] * xn + M[0, 1] * yn + M[0, 2] + 1.0) * W - 1.0) / 2.0)
                    ys.append(((M[1, 0] * xn + M[1, 1] * yn + M[1, 2] + 1.0) * H - 1.0) / 2.0)
            lo_x, hi_x = int(np.floor(min(xs))) - 2, int(np.floor(max(xs))) + 4
            lo_y, hi_y = int(np.floor(min(ys))) - 2, int(np.floor(max(ys))) + 4
            cx0, cx1 = min(max(lo_x, 0), W), min(max(hi_x, 0), W)
            cy0, cy1 = min(max(lo_y, 0), H), min(max(hi_y, 0), H)
            rects.append((cy0, max(cy1, cy0), cx0, max(cx1, cx0)))
        return rects

    def ego_tile_buffers(self, n, tile, slot=0):
        key = (n, tuple(tile), slot)
        if key not in self._ego_buf:
            d = self.device
            tp = (tile[1] - tile[0]) * (tile[3] - tile[2])
            self._ego_buf[key] = dict(
                feat=torch.zeros((n, self.ho, self.wo, self.c_feat), dtype=torch.float32, device=d),
                fused=torch.empty((tp, self.c_feat), dtype=torch.float32, device=d),
                preds=torch.empty((self.heads.cout, tp), dtype=torch.float32, device=d))
        return self._ego_buf[key]

    def decode_fuse_heads_tile(self, codes: torch.Tensor, affine: torch.Tensor, affine_host, tile,
                               slot=0) -> torch.Tensor:
        """The ego stage for ONE output tile (y0, y1, x0, x1): decode only the source rectangles the tile samples
        from, warp + fuse the tile, run the heads on it.  Returns compact preds [Cout, tile_pixels].  Per-pixel
        arithmetic is identical to decode_fuse_heads, so tiles assembled from several GPUs equal the 1-GPU result."""
        n = codes.shape[-1] // self.hw
        b = self.ego_tile_buffers(n, tile, slot)
        rects = self.source_rects(affine_host, tile)
        self.codebook.decode_regions(codes, self.wo, [a * self.hw for a in range(n)], rects,
                                     b["feat"].view(n * self.hw, self.c_feat))
        E.fuse_tile(b["feat"], affine, self.fusion_mode, tile, b["fused"])
        self.heads.forward(b["fused"], out=b["preds"])
        return b["preds"]

    def decode_fuse_heads_tile_to(self, codes: torch.Tensor, affine: torch.Tensor, affine_host, tile, out_ptr: int,
                                  slot=0) -> None:
        """As decode_fuse_heads_tile, but the head maps of the tile are stored at their place in a FULL
        [Cout, ho*wo] float32 map that starts at the raw device address `out_ptr` -- possibly the ego GPU's
        peer-mapped result buffer (distributed.PeerExchange), which makes the gather of head tiles unnecessary."""
        n = codes.shape[-1] // self.hw
        b = self.ego_tile_buffers(n, tile, slot)
        rects = self.source_rects(affine_host, tile)
        self.codebook.decode_regions(codes, self.wo, [a * self.hw for a in range(n)], rects,
                                     b["feat"].view(n * self.hw, self.c_feat))
        E.fuse_tile(b["feat"], affine, self.fusion_mode, tile, b["fused"])
        y0, _, x0, x1 = tile
        E.heads_forward_tile(self.heads, b["fused"], out_ptr + 4 * (y0 * self.wo + x0), x1 - x0, self.wo, self.hw)

    # ------------------------------------------------------------------ CUDA graphs
    def _capture(self, fn):
        """Capture `fn` (library launches on static buffers) into a CUDA graph: one launch replays the ~25 kernels
        of a stage without host round trips."""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(2):
                fn()                       # warm-up: one-time attribute / workspace setup must not be captured
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        return g, out

    def capture_encode(self, bev_static: torch.Tensor, slot=0):
        """Graph of encode_agents over a STATIC input buffer; returns (graph, codes tensor it writes)."""
        return self._capture(lambda: self.encode_agents(bev_static, slot))

    def capture_ego(self, codes_static: torch.Tensor, affine_static: torch.Tensor, slot=0):
        """Graph of decode_fuse_heads over STATIC code / pose buffers; returns (graph, preds tensor it writes)."""
        return self._capture(lambda: self.decode_fuse_heads(codes_static, affine_static, slot))

    def split_preds(self, preds: torch.Tensor, n_cls: int, n_reg: int, n_dir: int):
        """[Cout, hw] -> dict of NCHW tensors as the reference model returns them (batch 1)."""
        p = preds.view(1, -1, self.ho, self.wo)
        return {"cls_preds": p[:, :n_cls], "reg_preds": p[:, n_cls:n_cls + n_reg],
                "dir_preds": p[:, n_cls + n_reg:n_cls + n_reg + n_dir], "preds_tensor": p}


def heads_from_quant_modules(cls_head, reg_head, dir_head) -> E.HeadsEngine:
    """Concatenate the three 1x1 head QuantModules (act-quant disabled) into one FP32 GEMM with the
    de-quantized fake-quant weights the reference would use (quant_layer.py:392-398)."""
    ws, bs = [], []
    for nm, qm in (("cls_head", cls_head), ("reg_head", reg_head), ("dir_head", dir_head)):
        if not hasattr(qm, "weight_quantizer"):
            raise NotImplementedError(f"{nm} is not a QuantModule (listed in skip_quant_module_names?): the heads "
                                      "engine takes the fake-quant weights of wrapped heads")
        if not qm.disable_act_quant and qm.use_act_quant:
            raise NotImplementedError(f"{nm} keeps an output quantizer (disable_output_head_quantization: false); "
                                      "call QuantModel.disable_network_output_quantization() -- the heads kernel "
                                      "writes FP32 predictions (reference quant_model.py:129-136)")
        if not isinstance(qm.activation_function, (torch.nn.Identity, type(None))) and \
                type(qm.activation_function).__name__ not in ("StraightThrough", "Identity"):
            raise NotImplementedError(f"{nm} has a fused activation {type(qm.activation_function).__name__}")
        with torch.no_grad():
            w = qm.weight_quantizer(qm.weight) if qm.use_weight_quant else qm.org_weight
            b = qm.bias if qm.use_weight_quant else qm.org_bias
        ws.append(w.detach().reshape(w.shape[0], -1).cpu().numpy())
        bs.append(np.zeros(w.shape[0], np.float32) if b is None else b.detach().cpu().numpy())
    return E.HeadsEngine(np.concatenate(ws, 0), np.concatenate(bs, 0))
