"""ORACLE (test infrastructure): detection post-processing restated in numpy, used on BOTH sides of the
"boxes after NMS must be identical" check (north_star check #3).

Follows the reference's VoxelPostprocessor3Heads:
  generate_anchor_box   opencood/data_utils/post_processor/voxel_postprocessor_3heads.py:63-127
  post_process          :318-477  (sigmoid, max over class, score threshold, box decode, rotated NMS, range mask)
  delta_to_boxes3d      :581-635
  boxes_to_corners_3d   opencood/utils/box_utils_mc.py:200-246 (order 'hwl')
  nms_rotated           opencood/utils/box_utils_mc.py:665-710 (top-1000, greedy, IoU > thresh suppressed)
The reference computes polygon IoU with shapely==2.0.0 / GEOS (requirements.txt:13), which is absent here and
on the GPU box; the IoU of two convex quadrilaterals is restated with Sutherland-Hodgman clipping + the shoelace
formula.  Pinning: tests/golden/postprocess.npz (oracle/gen_golden_postprocess.py) holds the outputs of the
reference's own generate_anchor_box / post_process run with a stand-in for shapely's Polygon built on this same
clip -- it pins anchors, score / label selection, box decoding, corner order, NMS order and the range mask, but not
the polygon AREA arithmetic, which tests/test_golden_cpu.py::test_polygon_iou_closed_forms checks against closed
forms (parity with GEOS itself stays unpinned at that third-party boundary).
"""
from __future__ import annotations

import numpy as np

GT_RANGE = [-100, -40, -15, 100, 40, 15]          # opencood/data_utils/datasets/__init__.py:25


def generate_anchors(cfg_list, lidar_range, grid_wh, order="hwl"):
    """Returns all_anchors [num_class, H, W, n_rot, 7] (x, y, z, h, w, l, yaw) and anchors per location."""
    out, per_loc = [], []
    for cfg in cfg_list:
        gw, gh = grid_wh[0] // cfg["feature_map_stride"], grid_wh[1] // cfg["feature_map_stride"]
        sizes, rots, heights = np.array(cfg["anchor_sizes"]), np.array(cfg["anchor_rotations"]), cfg["anchor_bottom_heights"]
        per_loc.append(len(rots) * len(sizes) * len(heights))
        if cfg["align_center"]:
            xs, ys = (lidar_range[3] - lidar_range[0]) / gw, (lidar_range[4] - lidar_range[1]) / gh
            xo, yo = xs / 2, ys / 2
        else:
            xs, ys = (lidar_range[3] - lidar_range[0]) / (gw - 1), (lidar_range[4] - lidar_range[1]) / (gh - 1)
            xo, yo = 0, 0
        x = np.arange(lidar_range[0] + xo, lidar_range[3] + 1e-5, step=xs)
        y = np.arange(lidar_range[1] + yo, lidar_range[4] + 1e-5, step=ys)
        X, Y, Z = np.meshgrid(x, y, np.array(heights))
        a = np.concatenate([X, Y, Z], axis=-1)                                  # [H, W, 3]
        size = np.tile(sizes.reshape(1, -1, 3), (*a.shape[:2], 1))
        size = size[..., [2, 1, 0]] if order == "hwl" else size[..., [0, 2, 1]]
        a = np.concatenate([a, size], axis=-1)
        a = np.tile(a[:, :, None, :], (1, 1, len(rots), 1))
        r = np.tile(rots.reshape(1, 1, -1, 1), (*a.shape[:2], len(sizes), 1))
        out.append(np.concatenate([a, r], axis=-1))
    return np.stack(out), per_loc


def delta_to_boxes3d(deltas, anchors):
    """deltas [N, 7], anchors [N, 7] (xyzhwl yaw) -> boxes [N, 7]."""
    d = np.sqrt(anchors[:, 4] ** 2 + anchors[:, 5] ** 2)
    b = np.zeros_like(deltas)
    b[:, 0] = deltas[:, 0] * d + anchors[:, 0]
    b[:, 1] = deltas[:, 1] * d + anchors[:, 1]
    b[:, 2] = deltas[:, 2] * anchors[:, 3] + anchors[:, 2]
    b[:, 3:6] = np.exp(deltas[:, 3:6]) * anchors[:, 3:6]
    b[:, 6] = deltas[:, 6] + anchors[:, 6]
    return b


def bev_corners(boxes):
    """[N, 7] hwl boxes -> [N, 4, 2] BEV corners in the reference's corner order (first four of the 8)."""
    l, w, yaw = boxes[:, 5], boxes[:, 4], boxes[:, 6]
    tx = np.array([1, 1, -1, -1]) / 2.0
    ty = np.array([-1, 1, 1, -1]) / 2.0
    cx, cy = l[:, None] * tx, w[:, None] * ty
    c, s = np.cos(yaw)[:, None], np.sin(yaw)[:, None]
    # points @ [[c, s], [-s, c]]
    x = cx * c - cy * s + boxes[:, 0:1]
    y = cx * s + cy * c + boxes[:, 1:2]
    return np.stack([x, y], axis=-1)


def _area(poly):
    x, y = poly[:, 0], poly[:, 1]
    return 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))


def _clip(subject, clip):
    """Sutherland-Hodgman: clip convex polygon `subject` by convex polygon `clip` (both [k, 2])."""
    def signed(p):
        x, y = p[:, 0], p[:, 1]
        return np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))
    if signed(clip) < 0:
        clip = clip[::-1]
    out = [tuple(p) for p in subject]
    for i in range(len(clip)):
        a, b = clip[i], clip[(i + 1) % len(clip)]
        inp, out = out, []
        if not inp:
            break

        def inside(p):
            return (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0]) >= 0

        def inter(p, q):
            d1 = (b[0] - a[0], b[1] - a[1])
            d2 = (q[0] - p[0], q[1] - p[1])
            den = d1[0] * d2[1] - d1[1] * d2[0]
            t = ((p[0] - a[0]) * d2[1] - (p[1] - a[1]) * d2[0]) / den
            return (a[0] + t * d1[0], a[1] + t * d1[1])

        s = inp[-1]
        for e in inp:
            if inside(e):
                if not inside(s):
                    out.append(inter(s, e))
                out.append(e)
            elif inside(s):
                out.append(inter(s, e))
            s = e
    return np.array(out) if len(out) >= 3 else None


def polygon_iou(p, q):
    ap, aq = _area(p), _area(q)
    c = _clip(p, q)
    inter = _area(c) if c is not None else 0.0
    union = ap + aq - inter
    return inter / union if union > 0 else 0.0


def nms_rotated(corners, scores, threshold, top=1000):
    if corners.shape[0] == 0:
        return np.array([], dtype=np.int32)
    ixs = scores.argsort()[::-1][:top]
    pick = []
    while len(ixs) > 0:
        i = ixs[0]
        pick.append(i)
        iou = np.array([polygon_iou(corners[i], corners[j]) for j in ixs[1:]], dtype=np.float32)
        ixs = np.delete(np.delete(ixs, np.where(iou > threshold)[0] + 1), 0)
    return np.array(pick, dtype=np.int32)


def post_process(cls_preds, reg_preds, anchors, score_threshold=0.2, nms_thresh=0.15, gt_range=GT_RANGE):
    """cls_preds [1, A*C*C, H, W], reg_preds [1, 7*A*C, H, W] (float), anchors [C, H, W, A, 7].
    Returns (corners [K, 4, 2], scores [K], labels [K], boxes [K, 7]) after NMS and the range mask."""
    all_anchors = anchors.transpose(1, 2, 0, 3, 4).reshape(-1, 7)
    n = all_anchors.shape[0]
    prob = 1.0 / (1.0 + np.exp(-cls_preds.astype(np.float64).transpose(0, 2, 3, 1)))
    prob = prob.reshape(1, n, -1)
    cls_pred = prob.max(-1)[0]
    labels = prob.argmax(-1)[0] + 1
    reg = reg_preds.astype(np.float64).transpose(0, 2, 3, 1).reshape(n, 7)
    boxes = delta_to_boxes3d(reg, all_anchors)
    mask = cls_pred > score_threshold
    boxes, scores, labels = boxes[mask], cls_pred[mask], labels[mask]
    if boxes.shape[0] == 0:
        return np.zeros((0, 4, 2)), scores, labels, boxes
    corners = bev_corners(boxes)
    keep = nms_rotated(corners, scores, nms_thresh)
    corners, scores, labels, boxes = corners[keep], scores[keep], labels[keep], boxes[keep]
    lo, hi = np.array(gt_range[:2]), np.array(gt_range[3:5])
    m = np.all((corners >= lo) & (corners <= hi), axis=(1, 2))
    return corners[m], scores[m], labels[m], boxes[m]
