"""Generates tests/golden/fusion_weighted.npz and tests/golden/pyramid_blocks.npz by running the UNMODIFIED reference `weighted_fuse`
(opencood/models/fuse_modules/pyramid_fuse.py:17-62) with the score preparation of
QuantPyramidFusion.forward_collab (opencood/quant/quant_block.py:516-520) on seeded inputs, at the three pyramid
level shapes (C = 64 / 128 / 256 on H, H/2, H/4), and the reference QuantBottleneck (opencood/quant/quant_block.py:
100-134) over seeded ResNeXt blocks (tests/pyramid_cases.py).  Build container only:  python oracle/gen_golden_pyramid.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
from opencood.models.fuse_modules.pyramid_fuse import weighted_fuse  # noqa: E402
from opencood.utils.transformation_utils import normalize_pairwise_tfm  # noqa: E402

from opencood.models.sub_modules.resblock import BasicBlock, Bottleneck  # noqa: E402
from opencood.quant.quant_block import QuantBasicBlock, QuantBottleneck  # noqa: E402
from opencood.quant.quant_layer import UniformAffineQuantizer  # noqa: E402

from quantv2x_b200.synthetic import synthetic_poses  # noqa: E402
from tests.pyramid_cases import (BASIC_CASES, BLOCK_CASES, GROUPS, IN_DELTA, PYRAMID_AGENTS, PYRAMID_CFG,  # noqa: E402
                                 basic_tensors, block_tensors, pyramid_tensors)

WQ = dict(n_bits=8, channel_wise=True, scale_method="minmax")
AQ = dict(n_bits=8, channel_wise=False, scale_method="minmax", leaf_param=True, prob=1.0)

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    rng = np.random.default_rng(21)
    out = {}
    N = 4
    poses = synthetic_poses(N)
    # one far-away agent: every tap falls outside its map, so its warped score is exactly 0 everywhere
    poses[0, 0, 3, 0, 3] = 500.0
    aff = normalize_pairwise_tfm(torch.from_numpy(poses).float(), 80.0, 281.6, 1)
    out["poses"] = poses
    out["affine"] = aff.numpy().astype(np.float32)
    for lvl, (C, H, W) in enumerate([(64, 16, 24), (128, 8, 12), (256, 4, 6)]):
        feat = (rng.standard_normal((N, C, H, W)) * (rng.random((N, C, H, W)) > 0.3)).astype(np.float32)
        occ = (rng.standard_normal((N, 1, H, W)) * 3).astype(np.float32)
        with torch.no_grad():
            score = torch.sigmoid(torch.from_numpy(occ)) + 1e-4
            fused = weighted_fuse(torch.from_numpy(feat), score, torch.tensor([N]), aff, False)[0]
        out[f"l{lvl}.feat"] = feat
        out[f"l{lvl}.occ"] = occ
        out[f"l{lvl}.fused"] = fused.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "fusion_weighted.npz"), **out)
    print("fusion_weighted.npz", os.path.getsize(os.path.join(OUT, "fusion_weighted.npz")))


def with_bias(conv, wb):
    m = torch.nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding,
                        groups=conv.groups, bias=True)
    with torch.no_grad():
        m.weight.copy_(torch.from_numpy(wb[0]))
        m.bias.copy_(torch.from_numpy(wb[1]))
    return m


def gen_blocks():
    out = {}
    Bottleneck.expansion = 1                                   # as PyramidFusion.__init__ sets it (pyramid_fuse.py:71)
    for idx, (name, inplanes, planes, stride, H, W) in enumerate(BLOCK_CASES):
        t, q_in = block_tensors(idx)
        down = None
        if "down" in t:
            down = torch.nn.Sequential(torch.nn.Conv2d(inplanes, planes, 1, stride=stride, bias=False),
                                       torch.nn.Identity())
        # BN is folded into the convs before quantization (quant_model.py:14), i.e. the norm layers are identities
        b = Bottleneck(inplanes, planes, stride, down, groups=GROUPS, base_width=4, norm_layer=torch.nn.Identity)
        b.conv1, b.conv2, b.conv3 = with_bias(b.conv1, t["conv1"]), with_bias(b.conv2, t["conv2"]), with_bias(b.conv3, t["conv3"])
        if down is not None:
            b.downsample[0] = with_bias(b.downsample[0], t["down"])
        qb = QuantBottleneck(b, WQ, AQ).eval()
        qb.set_quant_state(True, True)
        quantizers = [m for m in qb.modules() if isinstance(m, UniformAffineQuantizer)]
        x = torch.from_numpy(q_in.astype(np.float32) * IN_DELTA)
        taps = {}
        hooks = [getattr(qb, n).register_forward_hook(lambda m, i, o, n=n: taps.__setitem__(n, o.detach().clone()))
                 for n in ("conv1", "conv2", "conv3")]
        if down is not None:
            hooks.append(qb.downsample.register_forward_hook(lambda m, i, o: taps.__setitem__("down", o.detach().clone())))
        with torch.no_grad():
            for q in quantizers:
                q.set_inited(False)
            qb(x)
            for q in quantizers:
                q.set_inited(True)
            y = qb(x)
        for h in hooks:
            h.remove()
        out[f"{name}.q_in"] = q_in
        names = ["conv1", "conv2", "conv3"] + (["down"] if down is not None else [])
        for n in names:
            m = qb.downsample if n == "down" else getattr(qb, n)
            out[f"{name}.{n}.w_delta"] = m.weight_quantizer.delta.detach().numpy().astype(np.float32).reshape(-1)
            out[f"{name}.{n}.w_zp"] = m.weight_quantizer.zero_point.detach().numpy().astype(np.float32).reshape(-1)
        for n, q, val in (("conv1", qb.conv1.act_quantizer, taps["conv1"]), ("conv2", qb.conv2.act_quantizer, taps["conv2"]),
                          ("out", qb.act_quantizer, y)):
            d, z = float(q.delta), float(q.zero_point)
            assert z == 0.0
            codes = torch.round(val / d)
            assert float((codes * d - val).abs().max()) < 1e-4 * d
            out[f"{name}.{n}.act_delta"] = np.float32(d)
            out[f"{name}.{n}.codes"] = codes.numpy().astype(np.uint8)
        out[f"{name}.conv3.out"] = taps["conv3"].numpy().astype(np.float32)       # FP32, no quantizer
        if down is not None:
            out[f"{name}.down.out"] = taps["down"].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "pyramid_blocks.npz"), **out)
    print("pyramid_blocks.npz", os.path.getsize(os.path.join(OUT, "pyramid_blocks.npz")))


def gen_basic_blocks():
    """QuantBasicBlock (quant_block.py:68-97) over the seeded two-conv blocks -> tests/golden/basic_blocks.npz."""
    out = {}
    for idx, (name, inplanes, planes, stride, H, W) in enumerate(BASIC_CASES):
        t, q_in = basic_tensors(idx)
        down = None
        if "down" in t:
            down = torch.nn.Sequential(torch.nn.Conv2d(inplanes, planes, 1, stride=stride, bias=False),
                                       torch.nn.Identity())
        b = BasicBlock(inplanes, planes, stride, down, norm_layer=torch.nn.Identity)
        b.conv1, b.conv2 = with_bias(b.conv1, t["conv1"]), with_bias(b.conv2, t["conv2"])
        if down is not None:
            b.downsample[0] = with_bias(b.downsample[0], t["down"])
        qb = QuantBasicBlock(b, WQ, AQ).eval()
        qb.set_quant_state(True, True)
        quantizers = [m for m in qb.modules() if isinstance(m, UniformAffineQuantizer)]
        x = torch.from_numpy(q_in.astype(np.float32) * IN_DELTA)
        taps = {}
        hooks = [qb.conv1.register_forward_hook(lambda m, i, o: taps.__setitem__("conv1", o.detach().clone()))]
        if down is not None:
            hooks.append(qb.downsample.register_forward_hook(lambda m, i, o: taps.__setitem__("down", o.detach().clone())))
        with torch.no_grad():
            for q in quantizers:
                q.set_inited(False)
            qb(x)
            for q in quantizers:
                q.set_inited(True)
            y = qb(x)
        for h in hooks:
            h.remove()
        out[f"{name}.q_in"] = q_in
        for n, q, val in (("conv1", qb.conv1.act_quantizer, taps["conv1"]), ("out", qb.act_quantizer, y)):
            d = float(q.delta)
            assert float(q.zero_point) == 0.0
            codes = torch.round(val / d)
            assert float((codes * d - val).abs().max()) < 1e-4 * d
            out[f"{name}.{n}.act_delta"] = np.float32(d)
            out[f"{name}.{n}.codes"] = codes.numpy().astype(np.uint8)
        if down is not None:
            out[f"{name}.down.out"] = taps["down"].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "basic_blocks.npz"), **out)
    print("basic_blocks.npz", os.path.getsize(os.path.join(OUT, "basic_blocks.npz")))


def gen_backbone():
    """QuantPyramidFusion.forward_collab (quant_block.py:504-541) of a small ResNeXt pyramid backbone: per-level
    codes of every agent, occupancy logits and fused features."""
    import opencood.quant.quant_block as qblock
    from opencood.models.fuse_modules.pyramid_fuse import PyramidFusion

    t, x = pyramid_tensors()
    pf = PyramidFusion(dict(PYRAMID_CFG), 64).eval()
    ident = torch.nn.Identity()
    for li, nb in enumerate(PYRAMID_CFG["layer_nums"]):
        for bi in range(nb):
            b, tb = getattr(pf.resnet, f"layer{li}")[bi], t[f"l{li}.b{bi}"]
            b.conv1, b.conv2, b.conv3 = with_bias(b.conv1, tb["conv1"]), with_bias(b.conv2, tb["conv2"]), with_bias(b.conv3, tb["conv3"])
            b.bn1 = b.bn2 = b.bn3 = ident                      # BN folded before quantization (quant_model.py:14)
            assert (b.downsample is not None) == ("down" in tb) and b.conv2.groups == GROUPS
            if b.downsample is not None:
                b.downsample[0], b.downsample[1] = with_bias(b.downsample[0], tb["down"]), ident
        setattr(pf, f"single_head_{li}", with_bias(getattr(pf, f"single_head_{li}"), t[f"head{li}"]))
        up = pf.deblocks[li][0]
        nu = torch.nn.ConvTranspose2d(up.in_channels, up.out_channels, up.kernel_size, stride=up.stride, bias=True)
        with torch.no_grad():
            nu.weight.copy_(torch.from_numpy(t[f"up{li}"][0]))
            nu.bias.copy_(torch.from_numpy(t[f"up{li}"][1]))
        pf.deblocks[li][0], pf.deblocks[li][1] = nu, ident
    q = qblock.QuantPyramidFusion(pf, WQ, AQ).eval()
    for m in q.modules():                                      # as QuantModel.set_quant_state (quant_model.py:107-110)
        if isinstance(m, (qblock.QuantModule, qblock.BaseQuantBlock)):
            m.set_quant_state(True, True)
    quantizers = [m for m in q.modules() if isinstance(m, UniformAffineQuantizer)]
    poses = synthetic_poses(PYRAMID_AGENTS)
    aff = normalize_pairwise_tfm(torch.from_numpy(poses).float(), 80.0, 281.6, 1)
    rl = torch.tensor([PYRAMID_AGENTS])
    fused_rec, feat_rec = [], {}
    orig_fuse = qblock.weighted_fuse

    def rec_fuse(xx, score, record_len, affine, align):
        o = orig_fuse(xx, score, record_len, affine, align)
        fused_rec.append(o.detach().clone())
        return o

    qblock.weighted_fuse = rec_fuse
    hooks = [getattr(q.resnet, f"layer{li}").register_forward_hook(lambda m, i, o, li=li: feat_rec.__setitem__(li, o.detach().clone()))
             for li in range(3)]
    with torch.no_grad():
        xt = torch.from_numpy(x)
        for m in quantizers:
            m.set_inited(False)
        q.forward_collab(xt, rl, aff)
        for m in quantizers:
            m.set_inited(True)
        fused_rec.clear()
        final, occ_list = q.forward_collab(xt, rl, aff)
    qblock.weighted_fuse = orig_fuse
    for h in hooks:
        h.remove()
    out = {"poses": poses, "affine": aff.numpy().astype(np.float32)}
    for li, nb in enumerate(PYRAMID_CFG["layer_nums"]):
        for bi in range(nb):
            b = getattr(q.resnet, f"layer{li}")[bi]
            for n, aq in (("conv1", b.conv1.act_quantizer), ("conv2", b.conv2.act_quantizer), ("out", b.act_quantizer)):
                assert float(aq.zero_point) == 0.0
                out[f"l{li}.b{bi}.{n}.act_delta"] = np.float32(float(aq.delta))
        d = float(getattr(q.resnet, f"layer{li}")[-1].act_quantizer.delta)
        codes = torch.round(feat_rec[li] / d)
        assert float((codes * d - feat_rec[li]).abs().max()) < 1e-4 * d
        out[f"l{li}.codes"] = codes.numpy().astype(np.uint8)
        out[f"l{li}.occ"] = occ_list[li].numpy().astype(np.float32)
        # single_head_i is wrapped WITH an active output quantizer (quant_block.py:474-478): its logits are on a grid
        hq = getattr(q, f"single_head_{li}").act_quantizer
        out[f"head{li}.act_delta"] = np.float32(float(hq.delta))
        out[f"head{li}.act_zp"] = np.float32(float(hq.zero_point))
        out[f"head{li}.act_bits"] = np.int32(int(hq.n_bits))
        out[f"l{li}.fused"] = fused_rec[li][0].numpy().astype(np.float32)
        # the deblock of this level: its 128 channels of the final [1, 384, H, W] feature, on its own grid
        aq = q.deblocks[li][0].act_quantizer
        du = float(aq.delta)
        assert float(aq.zero_point) == 0.0
        part = final[0, 128 * li:128 * (li + 1)]
        cu = torch.round(part / du)
        assert float((cu * du - part).abs().max()) < 1e-4 * du
        out[f"up{li}.act_delta"] = np.float32(du)
        out[f"up{li}.codes"] = cu.numpy().astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "pyramid_backbone.npz"), **out)
    print("pyramid_backbone.npz", os.path.getsize(os.path.join(OUT, "pyramid_backbone.npz")))


def gen_state_dict_layout():
    """Parameter / buffer names and shapes of the reference PyramidFusion and BasicBlock, so that the test can check
    that reference checkpoints load into quantv2x_b200.pyramid_modules unchanged."""
    import json

    from opencood.models.fuse_modules.pyramid_fuse import PyramidFusion

    out = {}
    for tag, nums in (("small", PYRAMID_CFG["layer_nums"]), ("heal", [3, 5, 8])):
        cfg = dict(PYRAMID_CFG, layer_nums=nums)
        out[f"pyramid_fusion.{tag}"] = {k: list(v.shape) for k, v in PyramidFusion(cfg, 64).state_dict().items()}
    down = torch.nn.Sequential(torch.nn.Conv2d(64, 128, 1, stride=2, bias=False), torch.nn.BatchNorm2d(128))
    out["basic_block"] = {k: list(v.shape) for k, v in BasicBlock(64, 128, 2, down).state_dict().items()}
    with open(os.path.join(OUT, "pyramid_state_dict_layout.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("pyramid_state_dict_layout.json", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    torch.manual_seed(0)
    gen_state_dict_layout()
    main()
    gen_blocks()
    gen_basic_blocks()
    gen_backbone()
