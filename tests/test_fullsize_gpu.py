"""Parity at the BENCHMARKED shape: one agent at 200 x 704 through the real [3, 5, 8] backbone +
shrinker plan and the codebook encoder, as bench.py builds and calibrates it.
  * uint8 features (100 x 352 x 256) == the chained integer oracle, all 24 layers, bit for bit;
  * code indices == the exact restatement of the encoder's arithmetic (encode_fixed_point), every row."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_one_agent_at_200x704_against_the_integer_oracle(cuda_device):
    import bench
    from oracle import codebook_oracle as co
    from oracle import int_oracle
    from quantv2x_b200.export import attach_engines, export_spec
    from quantv2x_b200.synthetic import synthetic_pillars

    torch.set_num_threads(os.cpu_count() or 8)
    q, bev_delta = bench.build_calibrated_model(cuda_device, "att", 8)
    attach_engines(q, bev_delta=bev_delta, device=cuda_device)
    pipe = q.model._pipelines["m1"]
    enc = q.hypes["model"]["args"]["m1"]["encoder_args"]
    pil = [torch.from_numpy(t).to(cuda_device) for t in synthetic_pillars(7, 1, enc["lidar_range"], enc["voxel_size"], 6000)]
    bev = pipe.pillar_engine.forward(*pil, 1)
    assert tuple(bev.shape) == (1, 200, 704, 64)
    codes = pipe.encode_agents(bev)                                    # uint8 [3, 1, 35200]
    torch.cuda.synchronize()
    feat = pipe.encode_buffers(1)["feat"].cpu().numpy()
    assert feat.shape == (1, 100, 352, 256)

    spec = export_spec(q, bev_delta)
    bev_np = bev.cpu().numpy()
    assert 0.03 < (bev_np.max(-1) > 0).mean() < 0.06, "BEV occupancy is not the 6000-pillar frame"
    _, feat_ref = int_oracle.backbone_chain(spec, bev_np)
    assert feat_ref.std() > 5, "degenerate activations"
    bad = np.argwhere(feat != feat_ref)
    assert bad.shape[0] == 0, f"{bad.shape[0]} feature bytes differ from the integer oracle, first at {bad[:3]}"

    from tests.test_codebook_gpu import library_fold

    p = co.params_from_state_dict(q.model.codebook.state_dict())
    lib_fe, fixed = library_fold(q.model.codebook.engine(), co.fold_encode(p))   # the library's own folded tables
    rows = feat.reshape(-1, 256)
    ref_codes = co.encode_fixed_point(lib_fe, rows, np.float32(pipe.feat_delta), fixed)
    got = codes.cpu().numpy()
    for l in range(3):
        assert np.array_equal(got[l].T.astype(np.int64), np.asarray(ref_codes[l]).astype(np.int64)), f"level {l}"
