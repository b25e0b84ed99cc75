"""GPU detection post-processing against the numpy restatement of the reference's
VoxelPostprocessor3Heads (oracle/postprocess_oracle.py): same boxes, in the same NMS pick order."""
import os

import numpy as np
import pytest
import torch

from oracle import postprocess_oracle as pp
from tests.test_golden_cpu import GOLD
from tests.test_golden_cpu import POST_CFG as CFG

pytestmark = pytest.mark.gpu


def _compare(preds, lidar_range, grid_wh, thr, nms, box_range, cuda_device):
    from quantv2x_b200.engine import PostProcessEngine

    anchors, _ = pp.generate_anchors(CFG, lidar_range, grid_wh)
    ref = pp.post_process(preds[None, :18], preds[None, 18:60], anchors, thr, nms, box_range)
    eng = PostProcessEngine(CFG, lidar_range, grid_wh, score_threshold=thr, nms_threshold=nms, box_range=box_range)
    h, w = preds.shape[1:]
    got = eng.forward(torch.from_numpy(np.ascontiguousarray(preds.reshape(preds.shape[0], h * w))).to(cuda_device))
    corners, scores, labels, boxes = [t.cpu().numpy() for t in got]
    assert len(scores) == len(ref[1]), (len(scores), len(ref[1]))
    # same pick order unless two scores are equal to the last bit (then only the order of those two may differ)
    o_ref, o_got = np.lexsort((ref[3][:, 0], -ref[1])), np.lexsort((boxes[:, 0], -scores))
    np.testing.assert_allclose(scores[o_got], ref[1][o_ref], rtol=1e-12)
    assert np.array_equal(labels[o_got], ref[2][o_ref])
    np.testing.assert_allclose(boxes[o_got], ref[3][o_ref], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(corners[o_got], ref[0][o_ref], rtol=1e-9, atol=1e-9)
    return len(scores), eng.last_candidates


def test_random_heads(cuda_device):
    rng = np.random.default_rng(0)
    lidar_range, grid_wh = [-12.8, -6.4, -3, 12.8, 6.4, 1], (64, 32)
    box_range = [-12.8, -6.4, -15, 12.8, 6.4, 15]
    for trial in range(3):
        preds = rng.normal(size=(72, 16, 32)).astype(np.float32)
        preds[:18] = preds[:18] * 2.0 - 3.0               # a few hundred anchors above 0.5
        preds[18:60] *= 0.3
        k, cand = _compare(preds, lidar_range, grid_wh, 0.5, 0.15, box_range, cuda_device)
        assert cand > 100 and 5 < k < cand
    # nothing above the threshold
    preds = np.full((72, 16, 32), -10.0, np.float32)
    k, cand = _compare(preds, lidar_range, grid_wh, 0.5, 0.15, box_range, cuda_device)
    assert k == 0 and cand == 0


@pytest.mark.parametrize("fusion", ["att", "max"])
def test_reference_golden_heads(cuda_device, fusion):
    """The reference's own head outputs (golden): boxes from the GPU post-processing == boxes from the restatement."""
    g = np.load(os.path.join(GOLD, f"e2e_{fusion}.npz"))
    preds = g["preds"][0]
    prob = 1 / (1 + np.exp(-preds[:18].astype(np.float64)))
    thr = float(np.quantile(prob.transpose(1, 2, 0).reshape(-1, 3).max(-1), 0.97))
    k, _ = _compare(preds, [-12.8, -6.4, -3, 12.8, 6.4, 1], (64, 32), thr, 0.15, [-12.8, -6.4, -15, 12.8, 6.4, 15],
                    cuda_device)
    assert k >= 10
