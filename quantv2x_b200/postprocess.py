"""Detection post-processing front end (mirror of the reference's VoxelPostprocessor3Heads.post_process call,
opencood/data_utils/post_processor/voxel_postprocessor_3heads.py:318-477; used by tools/inference_mc_quant.py:581-606
inside its timed region): head maps in, boxes after rotated NMS out, on the GPU (engine.PostProcessEngine ->
qv2x_postprocess_*).  Single-frame / ego-only (the cooperative model already fused the agents)."""
from __future__ import annotations

import torch

from .engine import PostProcessEngine


def anchor_config_from_hypes(post_cfg: dict):
    """The per-class anchor generator list the reference reads from postprocess.anchor_args
    ('anchor_generator_config', voxel_postprocessor_3heads.py:28-40); single-class yamls (l / w / h / r) are mapped
    to the same form."""
    aa = post_cfg["anchor_args"]
    if "anchor_generator_config" in aa:
        return list(aa["anchor_generator_config"])
    import math
    return [dict(anchor_sizes=[[aa["l"], aa["w"], aa["h"]]], anchor_rotations=[math.radians(r) for r in aa["r"]],
                 anchor_bottom_heights=[-1.0], align_center=True, feature_map_stride=aa["feature_stride"])]


class PostProcessor:
    def __init__(self, hypes: dict, grid_wh, anchor_cfg=None, score_threshold: float | None = None):
        post = hypes["postprocess"]
        cfg = anchor_cfg or anchor_config_from_hypes(post)
        rng = post["anchor_args"]["cav_lidar_range"]
        thr = post["target_args"]["score_threshold"] if score_threshold is None else score_threshold
        self.engine = PostProcessEngine(cfg, rng, grid_wh, score_threshold=thr, nms_threshold=post["nms_thresh"],
                                        box_range=post.get("gt_range", rng))

    def post_process(self, output_dict: dict) -> dict:
        """output_dict: what the model returns (preds_tensor [1, 72, H, W], or cls_preds + reg_preds).  Returns
        pred_corners [K, 4, 2], pred_scores [K], pred_labels [K] (1-based), pred_boxes [K, 7] (x y z h w l yaw),
        CUDA tensors in NMS pick order."""
        if "preds_tensor" in output_dict:
            p = output_dict["preds_tensor"]
        else:
            p = torch.cat([output_dict["cls_preds"], output_dict["reg_preds"]], dim=1)
        assert p.shape[0] == 1, "one frame at a time"
        p = p[0].reshape(p.shape[1], -1).contiguous().float()
        corners, scores, labels, boxes = self.engine.forward(p)
        return {"pred_corners": corners, "pred_scores": scores, "pred_labels": labels, "pred_boxes": boxes}
