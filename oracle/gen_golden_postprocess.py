"""Generates tests/golden/postprocess.npz by running the UNMODIFIED reference VoxelPostprocessor3Heads
(opencood/data_utils/post_processor/voxel_postprocessor_3heads.py: generate_anchor_box :63-127, post_process :318-477,
delta_to_boxes3d :581-635) and box_utils_mc (boxes_to_corners_3d, project_box3d, nms_rotated, range mask) on seeded
head maps.  Build container only:  python oracle/gen_golden_postprocess.py

Two third-party pieces are absent here and are stood in for, nothing else of the reference is touched:
  * opencood.utils.box_overlaps -- a Cython extension (training-time anchor matching only): an empty stub module
    (and opencood.visualization.*, plotting helpers on top of the absent matplotlib / open3d);
  * shapely.geometry.Polygon    -- used by nms_rotated through common_utils.convert_format / compute_iou for
    `a.intersection(b).area / a.union(b).area`: a stand-in class for CONVEX polygons whose areas come from the same
    Sutherland-Hodgman clip + shoelace formula as oracle/postprocess_oracle.py.  So this fixture pins everything of
    the post-processor EXCEPT the polygon area arithmetic, which tests/test_oracle_cpu.py checks against closed forms.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import postprocess_oracle as pp  # noqa: E402
from oracle import ref_shim  # noqa: E402

ref_shim.install()
sys.modules["opencood.utils.box_overlaps"] = types.SimpleNamespace(bbox_overlaps=None)
# plotting helpers (matplotlib / open3d are absent): never called by post_process
for _name in ("opencood.visualization", "opencood.visualization.vis_utils", "opencood.visualization.vis_utils_mc",
              "opencood.visualization.simple_vis", "opencood.visualization.debug_plot"):
    sys.modules[_name] = ref_shim._StubModule(_name)
    sys.modules[_name].__path__ = []


class _Area:
    def __init__(self, area):
        self.area = area


class ConvexPolygon:
    def __init__(self, pts):
        self.pts = np.asarray(pts, np.float64)
        self.area = pp._area(self.pts)

    def intersection(self, other):
        c = pp._clip(self.pts, other.pts)
        return _Area(pp._area(c) if c is not None else 0.0)

    def union(self, other):
        return _Area(self.area + other.area - self.intersection(other).area)


from opencood.utils import common_utils  # noqa: E402

common_utils.Polygon = ConvexPolygon
from opencood.data_utils.post_processor.voxel_postprocessor_3heads import VoxelPostprocessor3Heads  # noqa: E402

from tests.test_golden_cpu import POST_CFG as CFG  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    lidar_range, grid_wh = [-12.8, -6.4, -3, 12.8, 6.4, 1], (64, 32)
    cfg = [dict(c, class_name=n, matched_threshold=0.6, unmatched_threshold=0.45)
           for c, n in zip(CFG, ("vehicle", "pedestrian", "truck"))]
    params = dict(order="hwl", nms_thresh=0.15, target_args=dict(score_threshold=0.5),
                  anchor_args=dict(W=grid_wh[0], H=grid_wh[1], cav_lidar_range=lidar_range,
                                   anchor_generator_config=cfg))
    post = VoxelPostprocessor3Heads(params, train=False)
    anchors, per_loc = post.generate_anchor_box()
    all_anchors = np.stack(anchors)                                   # [C, H, W, A, 7]
    out = {"all_anchors": all_anchors.astype(np.float64), "lidar_range": np.array(lidar_range), "grid_wh": np.array(grid_wh)}
    rng = np.random.default_rng(3)
    for trial in range(2):
        preds = rng.normal(size=(72, 16, 32)).astype(np.float32)
        preds[:18] = preds[:18] * 2.0 - 3.0
        preds[18:60] *= 0.3
        data = {"ego": {"transformation_matrix": torch.eye(4), "all_anchors": torch.from_numpy(all_anchors).float(),
                        "num_anchors_per_location": per_loc}}
        outd = {"ego": {"cls_preds": torch.from_numpy(preds[None, :18]), "reg_preds": torch.from_numpy(preds[None, 18:60]),
                        "dir_preds": torch.from_numpy(preds[None, 60:72])}}
        with torch.no_grad():
            box3d, score_labels = post.post_process(data, outd)
        out[f"t{trial}.preds"] = preds
        out[f"t{trial}.box3d"] = box3d.numpy().astype(np.float32)              # [K, 8, 3] corners
        out[f"t{trial}.scores"] = score_labels[:, 0].numpy().astype(np.float32)
        out[f"t{trial}.labels"] = score_labels[:, 1].numpy().astype(np.int64)
        print(trial, "boxes", box3d.shape[0])
    np.savez_compressed(os.path.join(OUT, "postprocess.npz"), **out)
    print("postprocess.npz", os.path.getsize(os.path.join(OUT, "postprocess.npz")))


if __name__ == "__main__":
    main()
