#!/usr/bin/env python
"""bench.py -- W8A8 fused BEV frames/s of the quantized cooperative-perception forward on B200.

A cooperative frame: N_AGENTS (ego + 7) agents each run the PointPillars front end, the W8A8 BEV backbone +
shrinker and the codebook encoder; the code planes travel to the ego, which decodes, warps, fuses (attention) and runs
the detection heads.  On 1 GPU a "step" is one frame.  On G GPUs the agents are sharded over the ranks (8/G per GPU)
and a step is a batch of G consecutive frames: rank r runs the agent stage of ITS agents for the G frames in one launch
sequence (so every launch has the 1-GPU shape, 8 agent maps), the uint8 code planes travel all-to-all (frame f's
planes to rank f, peer-memory stores by qv2x_scatter_planes), and rank f runs the ego stage of frame f.  Per-GPU work
per step does not depend on G (weak scaling); `value` = frames of all ranks / time.  QV2X_MGPU=tiles selects the
latency-oriented round-1 layout instead (one frame per step, ego output tiled over the ranks, strong scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Prints ONE JSON line (rank 0).  See the repository's task contract for the keys.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_AGENTS = 8
BEV_H, BEV_W, BEV_C = 200, 704, 64
PILLARS = 6000
METRIC = "W8A8 fused BEV frames/s (ego + 7 agents, PointPillars V2X-Real 704x200, codebook m=1 k=128, att fusion)"
# algorithmic work per agent, SURVEY section 8(d): backbone blocks + shrinker (int8 GEMM-able)
GMAC_INT8_PER_AGENT = 75.26
SHRINK0_GMAC_PER_AGENT = 31.144       # the dominant kernel's layer: conv3x3 384->256 on 100x352
SHRINK1_GMAC_PER_AGENT = 20.763       # conv3x3 256->256 on 100x352


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--fusion", default="att", choices=["att", "max"])
    ap.add_argument("--w-bits", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    # parity / sweep configurations of BASELINE.json (the default is the metric's configuration)
    ap.add_argument("--agents", type=int, default=8, help="agents in the frame (ego included): 2, 4 or 8")
    ap.add_argument("--dict-size", type=int, default=0, help="codebook size k (0 = the yaml's 128)")
    ap.add_argument("--model", default="att", choices=["att", "pyramid"],
                    help="att (default): the metric's model; pyramid: an auxiliary line for the pyramid-fusion model "
                         "(SURVEY 8(f)-2: HeterPyramidCollabCodebookMC, codebook C=64 m=1 k=128), 1 GPU")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- model
# dram__bytes_read.sum + dram__bytes_write.sum of the shrinker's first conv per agent (profiles/r2_ncu_shrink0.txt)
NCU_SHRINK0_DRAM_BYTES_PER_AGENT = 19.08e6


def build_calibrated_model(device, fusion, w_bits, seed=1234, dict_size=0):
    """Seeded float model from the yaml -> QuantModel -> weight qparams -> one calibration forward (float path,
    on `device`) -> frozen W8A8.  Returns (qmodel, bev_delta)."""
    import torch

    from quantv2x_b200 import yaml_utils
    from quantv2x_b200.quant import QuantModel, set_weight_quantize_params
    from quantv2x_b200.synthetic import seeded_init

    hypes = yaml_utils.load_yaml(yaml_utils.default_config(fusion))
    if dict_size:
        hypes["model"]["args"]["codebook"]["dict_size"] = int(dict_size)
    torch.manual_seed(seed)
    model = yaml_utils.create_model(hypes).eval()
    seeded_init(model, seed)
    wq = dict(n_bits=w_bits, channel_wise=True, scale_method="minmax")
    aq = dict(n_bits=8, channel_wise=False, scale_method="minmax", leaf_param=True, prob=1.0)
    q = QuantModel(model, wq, aq).eval()
    q.disable_network_output_quantization()
    q.to(device)
    set_weight_quantize_params(q)
    # Pillar-level calibration (SURVEY 8(d) recipe): one synthetic agent's pillars through the float fake-quant
    # model -- PointPillar encoder included, so the BEV grid's scale is the PFN block's own quantizer
    from quantv2x_b200.synthetic import synthetic_pillars
    enc_args = hypes["model"]["args"]["m1"]["encoder_args"]
    vf, vc, vn = synthetic_pillars(99, 1, enc_args["lidar_range"], enc_args["voxel_size"], PILLARS)
    data = {"inputs_m1": {"voxel_features": torch.from_numpy(vf).to(device),
                          "voxel_coords": torch.from_numpy(vc).to(device),
                          "voxel_num_points": torch.from_numpy(vn).to(device)},
            "agent_modality_list": ["m1"], "pairwise_t_matrix": torch.eye(4).view(1, 1, 1, 4, 4).repeat(1, 5, 5, 1, 1),
            "record_len": torch.tensor([1])}
    mods = [m for m in q.modules() if hasattr(m, "act_quantizer")]
    q.set_quant_state(True, True)
    for m in mods:
        m.act_quantizer.set_inited(False)
    with torch.no_grad():
        q.model.calibration_forward(data)
    for m in mods:
        m.act_quantizer.set_inited(True)
    q.set_quant_state(True, True)
    bev_delta = q.model.encoder_m1.bev_delta()
    q.hypes = hypes
    return q, bev_delta


def clock_sampler(stop, out, device_index=0):
    """SM clock / throttle reasons while the timed region runs (NVML every 5 ms; nvidia-smi as a fallback)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        while not stop.is_set():
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            try:
                r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            out.append((float(sm), float(mx), [k for k, v in bits.items() if r & v]))
            stop.wait(0.005)
        return
    except Exception:
        pass
    q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                str(device_index)], capture_output=True, text=True, timeout=5)
            if r.returncode == 0 and r.stdout.strip():
                c = [v.strip() for v in r.stdout.strip().split("\n")[0].split(",")]
                out.append((float(c[0]), float(c[1]), [n for n, v in zip(names, c[2:6]) if v.lower().startswith("active")]))
        except Exception:
            pass
        stop.wait(0.05)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
    sm = sorted(s[0] for s in samples)
    reasons = sorted({r for s in samples for r in s[2]})
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": samples[0][1], "reasons": reasons, "samples": len(samples)}


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_frames_per_s(spec, pspec, enc_args, n_agents, steps, warmup, budget_s=240.0):
    """The reference's CPU fake-quant path (oracle port, oracle/frame_ref.py) on WHOLE frames: every step runs all
    `n_agents` agents from their pillars (PointPillars front end, backbone, shrinker, encode) and the ego stage
    (decode, warp, fuse, heads) -- the work one step of the GPU arm does -- with all host threads.  Steps are cut
    short only if the wall budget runs out (then the line says how many were timed)."""
    import torch

    from oracle import frame_ref
    from oracle.fusion_oracle import normalize_pairwise_tfm
    from quantv2x_b200.synthetic import synthetic_pillars, synthetic_poses

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vf, vc, vn = synthetic_pillars(0, n_agents, enc_args["lidar_range"], enc_args["voxel_size"], PILLARS)
    aff = normalize_pairwise_tfm(synthetic_poses(n_agents), 80.0, 281.6, 1.0)[0, 0, :n_agents]
    times = []
    t_start = time.time()
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        frame_ref.frame_from_pillars(spec, pspec, vf, vc, vn, n_agents, aff)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if time.time() - t_start > budget_s and len(times) >= 1:
            break
    t = float(np.mean(times))
    sample = (f"{len(times)} whole frame(s) of {n_agents} agents from pillars ({PILLARS} per agent): PointPillars front "
              f"end + backbone + shrinker + codebook encode per agent, then decode + warp + {spec['fusion']} fusion + "
              f"heads; {warmup} warm-up frame(s); torch {torch.__version__} CPU fp32 fake-quant, {cores} threads")
    return 1.0 / t, cores, sample, t, len(times)


# ----------------------------------------------------------------------------------------------- main
def pyramid_main(args):
    """Auxiliary bench line of the pyramid-fusion model driver (1 GPU): one frame = pillars of N agents -> front end ->
    agent ResNet backbone -> codebook encode | decode -> ResNeXt pyramid over all agents -> deblocks -> shrink conv ->
    heads.  The agent stage is launched from Python, the ego stage replays a CUDA graph (capture_decode)."""
    import torch

    from quantv2x_b200 import _lib, yaml_utils
    from quantv2x_b200.pyramid_model import attach_pyramid_engines
    from quantv2x_b200.quant import QuantModel, set_weight_quantize_params
    from quantv2x_b200.synthetic import seeded_init, seeded_init_codebook, synthetic_pillars, synthetic_poses

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    dev = torch.device("cuda", 0)
    n = args.agents
    here = os.path.dirname(os.path.abspath(yaml_utils.__file__))
    hy = yaml_utils.load_yaml(os.path.join(here, "hypes_yaml/v2x_real/Codebook/Pyramid/lidar_pyramid_stage3.yaml"))
    model = yaml_utils.create_model(hy).eval()
    seeded_init(model, 1234)
    seeded_init_codebook(model.codebook, 4321)
    q = QuantModel(model, dict(n_bits=8, channel_wise=True, scale_method="minmax"),
                   dict(n_bits=8, channel_wise=False, scale_method="minmax", leaf_param=True, prob=1.0)).eval()
    q.disable_network_output_quantization()
    q.to(dev)
    set_weight_quantize_params(q)
    enc = hy["model"]["args"]["m1"]["encoder_args"]

    def frame(seed, agents, device):
        vf, vc, vn = synthetic_pillars(seed, agents, enc["lidar_range"], enc["voxel_size"], PILLARS)
        return {"inputs_m1": {"voxel_features": torch.from_numpy(vf).to(device),
                              "voxel_coords": torch.from_numpy(vc).to(device),
                              "voxel_num_points": torch.from_numpy(vn).to(device)},
                "agent_modality_list": ["m1"] * agents,
                "pairwise_t_matrix": torch.from_numpy(synthetic_poses(agents)).float(),
                "record_len": torch.tensor([agents])}

    mods = [m for m in q.modules() if hasattr(m, "act_quantizer")]
    q.set_quant_state(True, True)
    for m in mods:
        m.act_quantizer.set_inited(False)
    with torch.no_grad():
        q.model.calibration_forward(frame(99, 2, dev))          # offline calibration (torch body)
    for m in mods:
        m.act_quantizer.set_inited(True)
    attach_pyramid_engines(q, device=dev)
    mdl = q.model
    host = frame(0, n, "cpu")
    host_in = {k: v.pin_memory() for k, v in host["inputs_m1"].items()}
    data = dict(host, inputs_m1={k: v.to(dev) for k, v in host_in.items()})
    codes, _, info = mdl.encode_features(data)
    codes_u8 = torch.stack([c.t() for c in codes]).to(torch.uint8).contiguous()
    graph, gout = mdl.capture_decode(codes_u8, info)
    eng = mdl._engines
    l0 = _lib.lib().qv2x_launch_count()
    mdl.decode_features(codes_u8, info)
    ego_launches = _lib.lib().qv2x_launch_count() - l0

    def step(upload):
        if upload:
            for k, v in host_in.items():
                data["inputs_m1"][k].copy_(v, non_blocking=True)
        c, _, _ = mdl.encode_features(data)
        for l in range(len(c)):                    # the wire payload: byte planes into the graph's static buffer
            codes_u8[l].copy_(c[l].t())
        graph.replay()
        return gout["preds_tensor"]

    preds_host = torch.empty(tuple(gout["preds_tensor"].shape), dtype=torch.float32).pin_memory()

    def timed(k, upload):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            p = step(upload)
            if upload:
                preds_host.copy_(p, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    timed(max(args.warmup, 3), False)
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clock_sampler, args=(stop, samples, 0), daemon=True)
    th.start()
    l0 = _lib.lib().qv2x_launch_count()
    ms = timed(args.steps, False)
    enc_launches = (_lib.lib().qv2x_launch_count() - l0) // args.steps
    timed(2, True)
    e2e_ms = timed(args.steps, True)
    stop.set()
    th.join(timeout=2)
    import hashlib
    in_bytes = sum(int(v.numel()) * v.element_size() for v in host_in.values())
    line = {"metric": f"W8A8 pyramid-fusion frames/s (ego + {n - 1} agents, PointPillars V2X-Real 704x200, ResNet agent "
                      "backbone, codebook C=64 m=1 k=128, ResNeXt pyramid [3,5,8], shrink conv, heads)",
            "value": 1e3 * args.steps / ms, "unit": "frames/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 (int8 tensor cores, int32 accumulate)", "data": "synthetic",
            "preds_sha1": hashlib.sha1(gout["preds_tensor"].cpu().numpy().tobytes()).hexdigest(),
            "config": {"workload": f"pyramid model, ego+{n - 1} agents, {PILLARS} pillars per agent; SURVEY 8(f)-2 "
                                   "(auxiliary line: BASELINE.json's metric is the att-fusion model)",
                       "agents": n, "l2": "activations of the 16 ResNeXt blocks over all agents (> 126 MB per frame) "
                                          "are rewritten every step; the input pillars are the same frame"},
            "e2e": {"value": 1e3 * args.steps / e2e_ms, "unit": "frames/s", "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": int(preds_host.numel() * 4)},
            "gpu_launches": int((enc_launches + ego_launches) * args.steps), "clocks": summarize_clocks(samples),
            "roofline": None, "cpu_baseline": None,
            "note": "agent stage launched from Python, ego stage replayed from a CUDA graph; see "
                    "profiles/r2_launches_pyramid_model.txt for the per-kernel split"}
    print(json.dumps(line), flush=True)


def main():
    global N_AGENTS, METRIC
    args = parse()
    if args.model == "pyramid":
        if int(os.environ.get("RANK", "0")) == 0:
            pyramid_main(args)
        return
    if args.agents != 8 or args.dict_size or args.w_bits != 8 or args.fusion != "att":
        N_AGENTS = args.agents
        METRIC = (f"W{args.w_bits}A8 fused BEV frames/s (ego + {N_AGENTS - 1} agents, PointPillars V2X-Real 704x200, "
                  f"codebook m=1 k={args.dict_size or 128}, {args.fusion} fusion)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"ego+{N_AGENTS - 1} agents ({N_AGENTS}), {PILLARS} pillars x 32 points per agent -> BEV {BEV_W}x{BEV_H}x{BEV_C} uint8, "
                          f"W{args.w_bits}A8 backbone+shrinker, codebook m=1 k={args.dict_size or 128} x3 levels, {args.fusion} fusion, "
                          "heads 72 ch; SURVEY 8(d)",
              "agents": N_AGENTS, "agents_per_gpu": N_AGENTS // max(world, 1), "parallelism": f"agents/{world}gpu",
              "l2": "per-frame working set (72 MB inputs + ~45 MB activations per agent) exceeds the 126 MB L2; a "
                    "256 MB buffer is also rewritten between timed steps"}

    if args.impl == "reference":
        if rank != 0:
            return
        import torch

        from quantv2x_b200.export import export_spec, pillar_spec
        q, bev_delta = build_calibrated_model(torch.device("cpu"), args.fusion, args.w_bits, dict_size=args.dict_size)
        spec = export_spec(q, bev_delta)
        pspec = pillar_spec(q.model.encoder_m1)
        enc_args = q.hypes["model"]["args"]["m1"]["encoder_args"]
        fps, cores, sample, t, done = cpu_reference_frames_per_s(spec, pspec, enc_args, N_AGENTS, args.steps,
                                                                  min(args.warmup, 1), budget_s=240.0)
        config["l2"] = "n/a (host arm): every step recomputes the whole frame from its pillars"
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": done, "steps_requested": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 fake-quant (simulated u8)", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist

    from quantv2x_b200 import _lib
    from quantv2x_b200.export import attach_engines, export_spec
    from quantv2x_b200.synthetic import synthetic_poses
    from quantv2x_b200.collab_model import normalize_pairwise_tfm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the "
                         "CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    _lib.check(_lib.lib().qv2x_device_check(local_rank))
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert N_AGENTS % world == 0, "agent count must divide over the GPUs"
    per = N_AGENTS // world
    # Multi-GPU modes (both shard the AGENTS over the ranks, 8/G per rank, and exchange integer code planes):
    #   frames (default): a step is a batch of G consecutive frames.  Rank r runs the agent stage of ITS agents for all
    #       G frames in one launch sequence (8 agent maps per launch, the 1-GPU kernel shapes), the code planes travel
    #       all-to-all (frame f's planes go to rank f) and rank f runs the whole ego stage of frame f: the ego role
    #       rotates over the ranks.  Per-GPU work per step is constant in G (weak scaling).
    #   tiles: one frame per step; every rank receives all code planes and computes one TILE of the ego output
    #       (lowest latency, but 8/G-agent launches leave the SMs under-filled; round-1 design).
    MGPU = os.environ.get("QV2X_MGPU", "frames") if world > 1 else "single"
    assert MGPU in ("frames", "tiles", "single")
    F = world if MGPU == "frames" else 1          # frames per step
    n_img = per * F                               # agent maps per launch on this rank
    config["frames_per_step"] = F
    if world > 1:
        config["parallelism"] = (f"agents/{world}gpu, {F} frames per step, ego role rotates over the ranks"
                                 if MGPU == "frames" else f"agents/{world}gpu, ego output tiled over the ranks")

    q, bev_delta = build_calibrated_model(device, args.fusion, args.w_bits, dict_size=args.dict_size)
    attach_engines(q, bev_delta=bev_delta, device=device)
    pipe = q.model._pipelines["m1"]

    # this rank's agents as PILLARS in the reference's dict schema (voxel_features / voxel_coords / voxel_num_points);
    # the frame starts at the PointPillars front end (qv2x_pillar_forward), as the reference model does
    from quantv2x_b200.synthetic import synthetic_pillars
    enc_args = q.hypes["model"]["args"]["m1"]["encoder_args"]
    pillar = getattr(pipe, "pillar_engine", None)
    if pillar is None:
        raise SystemExit("the calibrated PointPillar encoder did not yield a pillar engine")

    def pillar_set(seed):
        """(features [per*P, 32, 4] f32, coords [per*P, 4] i32 with LOCAL agent index, num_points [per*P] i32), host."""
        vf, vc, vn = synthetic_pillars(seed, N_AGENTS, enc_args["lidar_range"], enc_args["voxel_size"], PILLARS)
        sel = (vc[:, 0] >= rank * per) & (vc[:, 0] < (rank + 1) * per)
        vc = vc[sel].copy()
        vc[:, 0] -= rank * per
        return (torch.from_numpy(np.ascontiguousarray(vf[sel])), torch.from_numpy(vc),
                torch.from_numpy(np.ascontiguousarray(vn[sel])))

    pil_frame0 = pillar_set(0)
    poses = torch.from_numpy(synthetic_poses(N_AGENTS)).float()
    aff = normalize_pairwise_tfm(poses, 80.0, 281.6, 1)[0, 0, :N_AGENTS].contiguous().to(device)
    aff_host = aff.cpu().numpy()
    levels, m, hw = pipe.codebook.levels, pipe.codebook.m, pipe.hw
    preds_host = torch.empty((pipe.heads.cout, hw), dtype=torch.float32).pin_memory()

    from quantv2x_b200.distributed import all_gather_code_planes, gather_pred_tiles, rank_tile
    from quantv2x_b200.engine import push_planes

    # ---- frames in flight and input rotation.
    # INFLIGHT frames are processed concurrently on their own streams and buffer sets (a serving loop pipelines
    # frames: frame i+1's backbone overlaps frame i's exchange + ego stage); every frame still runs the complete
    # path.  The input of step i is buffer i % R of a pool of R distinct frames whose total size exceeds the 126 MB
    # L2, so no step finds its input in L2 (this replaces the flush buffer, which would serialise the pipeline).
    INFLIGHT = int(os.environ.get("QV2X_INFLIGHT", "2" if world == 1 else "4"))     # short per-rank stages: deeper pipeline
    in_bytes = F * sum(int(t.numel()) * t.element_size() for t in pil_frame0)      # one step's input on this rank
    R = max(INFLIGHT, -(-(140 * 1024 * 1024) // in_bytes))
    R += R % INFLIGHT
    base = tuple(t.to(device) for t in pil_frame0)

    def frame_variant(v):
        """Frame v of this rank's agents: distinct frames of the same statistics -- the point clouds of frame 0 with
        every pillar moved to another cell (v = 0: frame 0 itself)."""
        f0, c0, n0 = base
        if v == 0:
            return f0, c0, n0
        shift_y, shift_x = 3 * v, 5 * v
        c = c0.clone()
        dy = (c[:, 2] + shift_y) % BEV_H - c[:, 2]
        dx = (c[:, 3] + shift_x) % BEV_W - c[:, 3]
        c[:, 2] += dy
        c[:, 3] += dx
        f = f0.clone()
        live = (torch.arange(32, device=device)[None, :] < n0[:, None]).to(f.dtype)
        f[:, :, 0] += dx[:, None].to(f.dtype) * enc_args["voxel_size"][0] * live
        f[:, :, 1] += dy[:, None].to(f.dtype) * enc_args["voxel_size"][1] * live
        return f, c, n0

    pil_pool = []
    for r in range(R):
        # the input of step r: frames r*F .. r*F+F-1, frame-major agent maps (map index = frame * per + local agent)
        fs, cs, ns = [], [], []
        for fi in range(F):
            f_, c_, n_ = frame_variant(r * F + fi)
            c_ = c_.clone()
            c_[:, 0] += fi * per
            fs.append(f_), cs.append(c_), ns.append(n_)
        pil_pool.append((torch.cat(fs).contiguous(), torch.cat(cs).contiguous(), torch.cat(ns).contiguous()))
    pil_host = tuple(t.cpu().pin_memory() for t in pil_pool[0])
    config["inflight"] = INFLIGHT
    config["input"] = (f"pillars [{n_img}x{PILLARS}, 32, 4] f32 + coords + point counts per rank and step "
                       f"({in_bytes / 2**20:.1f} MB)")
    config["l2"] = (f"the input of step i is pillar set i % {R} of {R} distinct frames ({R * in_bytes / 2**20:.0f} MB per "
                    "rank, more than the 126 MB L2); no flush between steps because frames are pipelined")

    # CUDA graphs over static buffers: one replay per stage instead of ~25 launches (collectives stay outside).
    #   1 GPU : encode graph (8 agents) -> ego graph (decode + warp/fuse + heads on the whole map)
    #   G GPUs: encode graph (8/G agents) -> all_gather of the uint8 code planes -> ego graph on this rank's output
    #           TILE (decode only the source rectangles the tile samples, fuse, heads) -> gather of the head tiles
    lc0 = _lib.lib().qv2x_launch_count()
    g_enc, codes_local = [None] * R, [None] * INFLIGHT
    for r in range(R):
        g_enc[r], codes_local[r % INFLIGHT] = pipe._capture(
            lambda r=r: pipe.encode_pillars(*pil_pool[r], n_img, slot=r % INFLIGHT))
    lc1 = _lib.lib().qv2x_launch_count()
    g_ego, preds_dev = [None] * INFLIGHT, [None] * INFLIGHT
    recv_codes, codes_full, recv_preds = [None] * INFLIGHT, [None] * INFLIGHT, [None] * INFLIGHT
    tile = rank_tile(rank, world, pipe.ho, pipe.wo) if MGPU == "tiles" else None
    # Multi-GPU exchange: peer-memory stores from our own kernels + two device barriers (PeerExchange); NCCL
    # all-gather / gather only if the symmetric-memory mapping cannot be set up on this box.
    px = None
    if world > 1 and not os.environ.get("QV2X_NCCL_EXCHANGE"):
        try:
            from quantv2x_b200.distributed import PeerExchange
            px = PeerExchange(device, levels, m, N_AGENTS * hw, pipe.heads.cout, hw, slots=INFLIGHT)
        except Exception as exc:          # noqa: BLE001
            if rank == 0:
                print(f"[bench] peer-memory exchange unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
            px = None
        ok = torch.tensor([1 if px is not None else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            px = None
    if world == 1:
        config["exchange"] = "none (1 GPU)"
    elif MGPU == "frames":
        config["exchange"] = ("all-to-all of code planes by peer-memory stores (qv2x_scatter_planes) + device barriers"
                              if px else "NCCL all_to_all of code planes")
    else:
        config["exchange"] = "peer-memory stores + device barriers" if px else "NCCL all_gather + gather"
    for sl in range(INFLIGHT):
        if world == 1:
            g_ego[sl], preds_dev[sl] = pipe.capture_ego(codes_local[sl], aff, slot=sl)
        elif MGPU == "frames":
            codes_full[sl] = (px.codes_full(sl) if px is not None else
                              torch.zeros((levels, m, N_AGENTS * hw), dtype=torch.uint8, device=device))
            if px is None:
                recv_codes[sl] = torch.empty((world, levels, m, per * hw), dtype=torch.uint8, device=device)
            g_ego[sl], preds_dev[sl] = pipe.capture_ego(codes_full[sl], aff, slot=sl)
        elif px is not None:
            codes_full[sl] = px.codes_full(sl)
            g_ego[sl], _ = pipe._capture(
                lambda sl=sl: pipe.decode_fuse_heads_tile_to(codes_full[sl], aff, aff_host, tile, px.preds_ptr(0, sl),
                                                             slot=sl))
            preds_dev[sl] = px.preds_full(sl)
        else:
            recv_codes[sl] = torch.empty((world, levels, m, per * hw), dtype=torch.uint8, device=device)
            codes_full[sl] = torch.zeros((levels, m, N_AGENTS * hw), dtype=torch.uint8, device=device)
            g_ego[sl], preds_dev[sl] = pipe._capture(
                lambda sl=sl: pipe.decode_fuse_heads_tile(codes_full[sl], aff, aff_host, tile, slot=sl))
            recv_preds[sl] = (torch.empty((world,) + tuple(preds_dev[sl].shape), dtype=torch.float32, device=device)
                              if rank == 0 else None)
    lc2 = _lib.lib().qv2x_launch_count()
    # kernels per replay = launches recorded while capturing (2 warm-up calls + 1 captured call per graph)
    launches_per_step = (lc1 - lc0) // (3 * R) + (lc2 - lc1) // (3 * INFLIGHT)
    streams = [torch.cuda.Stream() for _ in range(INFLIGHT)]

    from quantv2x_b200.distributed import all_to_all_code_planes
    from quantv2x_b200.engine import scatter_planes

    def step(i):
        """Step i on the CURRENT stream: input buffer i % R, buffer set i % INFLIGHT.  Returns the head maps of the
        frame this rank fused (frames mode: frame `rank` of the step; tiles mode: the frame, on rank 0 only)."""
        r, sl = i % R, i % INFLIGHT
        g_enc[r].replay()
        if world == 1:
            g_ego[sl].replay()
            return preds_dev[sl]
        exchange_codes(sl)
        g_ego[sl].replay()
        return collect_preds(sl)

    def exchange_codes(sl):
        if MGPU == "frames":
            if px is not None:
                scatter_planes(codes_local[sl], N_AGENTS * hw, rank * per * hw, px.code_ptrs(sl))
                px.barrier(sl)          # every rank's planes of MY frame have landed in my buffer
            else:
                codes_full[sl].copy_(all_to_all_code_planes(codes_local[sl], recv=recv_codes[sl]))
            return
        if px is not None:
            push_planes(codes_local[sl], N_AGENTS * hw, rank * per * hw, px.code_ptrs(sl))
            px.barrier(sl)              # every rank's planes have landed in this rank's buffer
        else:
            codes_full[sl].copy_(all_gather_code_planes(codes_local[sl], hw, recv=recv_codes[sl]))

    def collect_preds(sl):
        if MGPU == "frames":
            if px is not None:
                px.barrier(sl)          # every rank has consumed its code buffer: the slot may be refilled
            return preds_dev[sl]
        if px is not None:
            px.barrier(sl)              # every rank's head tile has landed in the ego rank's buffer
            return preds_dev[sl] if rank == 0 else None
        return gather_pred_tiles(preds_dev[sl], pipe.ho, pipe.wo, dst=0, recv=recv_preds[sl])

    def phase_times(k=10):
        """Device time of the step's phases on this rank (CUDA events between them, one frame at a time)."""
        names = (["encode_graph", "exchange_codes", "ego_graph", "gather_preds" if MGPU == "tiles" else "slot_barrier"]
                 if world > 1 else ["encode_graph", "ego_graph"])
        acc = [0.0] * len(names)
        for i in range(k):
            r, sl = i % R, i % INFLIGHT
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            evs[0].record()
            g_enc[r].replay()
            evs[1].record()
            if world == 1:
                g_ego[sl].replay()
                evs[2].record()
            else:
                exchange_codes(sl)
                evs[2].record()
                g_ego[sl].replay()
                evs[3].record()
                collect_preds(sl)
                evs[4].record()
            torch.cuda.synchronize()
            for j in range(len(names)):
                acc[j] += evs[j].elapsed_time(evs[j + 1])
        return {n: a / k for n, a in zip(names, acc)}

    def timed_loop(fn, k):
        """k calls of fn on the current stream, each bracketed by CUDA events (used for single kernels)."""
        evs = []
        for _ in range(k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)

    def run_steps(k, inflight):
        """k steps, `inflight` frames concurrently (step i on stream i % inflight); one event pair around all."""
        main = torch.cuda.current_stream()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(main)
        for st in streams[:inflight]:
            st.wait_event(start)
        for i in range(k):
            with torch.cuda.stream(streams[i % inflight]):
                step(i)
        for st in streams[:inflight]:
            main.wait_stream(st)
        end.record(main)
        torch.cuda.synchronize()
        return start.elapsed_time(end)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run_steps(max(args.warmup, 3), INFLIGHT)
    sync_all()
    stop, samples = threading.Event(), []
    th = None
    if rank == 0:
        th = threading.Thread(target=clock_sampler, args=(stop, samples, local_rank), daemon=True)
        th.start()
    sync_all()
    torch.cuda.nvtx.range_push("timed")          # ncu --nvtx --nvtx-include "timed/" lists exactly these launches
    total_ms = run_steps(args.steps, INFLIGHT)
    torch.cuda.nvtx.range_pop()
    sync_all()
    serial_ms = run_steps(args.steps, 1)          # one frame at a time: the single-frame latency
    sync_all()
    phases = phase_times()
    sync_all()
    launches = launches_per_step * args.steps      # kernels of this library replayed through the CUDA graphs
    t = torch.tensor([total_ms, serial_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, serial_ms = float(t[0].item()), float(t[1].item())
    ms_per_step = total_ms / args.steps
    fps = 1e3 * F / ms_per_step

    # ---- e2e: host buffers in, host result out, through the same public calls.  Every step copies ITS inputs from
    # pinned host memory into its input buffer and reads ITS result back; the copies run on their own streams so
    # that step i+1's upload and step i-1's readback overlap step i, as a serving loop would.  One event pair
    # brackets all K steps including the first (un-overlapped) upload and the last readback.
    s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
    preds_stage = [torch.empty((pipe.heads.cout, hw), dtype=torch.float32, device=device) for _ in range(INFLIGHT)]
    preds_hosts = [preds_host] + [torch.empty_like(preds_host).pin_memory() for _ in range(INFLIGHT - 1)]
    # detection post-processing on the GPU (qv2x_postprocess_*: score threshold, box decode, rotated NMS) -- inside the
    # reference's timed region too (inference_mc_quant.py:581-606, on the CPU there); one handle per frame in flight
    from quantv2x_b200.postprocess import PostProcessor
    owns_result = rank == 0 or MGPU == "frames"      # frames mode: every rank fuses (and post-processes) one frame
    ppe, pp_out, pp_host = [], [], []
    pp_threshold = None
    if owns_result:
        grid_wh = (BEV_W, BEV_H)
        # Random-init heads never reach the yaml's score threshold (0.2), which would leave NMS without work; the
        # bench sets the threshold to the score that ~300 anchors of frame 0 exceed (a busy real frame), so that the
        # timed post-processing includes candidate ranking, the rotated-IoU matrix and the greedy pass.
        n_cls = 2 * 3 * 3
        p0 = step(0) if world == 1 else preds_dev[0]            # (multi-GPU: the last warm-up step's result)
        torch.cuda.synchronize()
        sc = torch.sigmoid(p0[:n_cls].float().flatten())
        pp_threshold = float(torch.topk(sc, 300).values[-1].item())
        if os.environ.get("QV2X_BENCH_DEBUG"):
            e_dbg = PostProcessor(q.hypes, grid_wh, score_threshold=pp_threshold).engine
            r_dbg = e_dbg.forward(p0.contiguous())
            print(f"[bench debug] threshold {pp_threshold}, elements above {(sc > pp_threshold).sum().item()}, "
                  f"candidates {e_dbg.last_candidates}, boxes {r_dbg[0].shape[0]}, cls logits min/max "
                  f"{p0[:n_cls].min().item():.3f}/{p0[:n_cls].max().item():.3f}", file=sys.stderr)
        for _ in range(INFLIGHT):
            e = PostProcessor(q.hypes, grid_wh, score_threshold=pp_threshold).engine
            ppe.append(e)
            outs = e.alloc_outputs(device)
            pp_out.append(outs)
            pp_host.append(tuple(torch.empty_like(o, device="cpu").pin_memory() for o in outs))
    box_bytes = sum(int(o.numel()) * o.element_size() for o in pp_out[0]) if owns_result else 0

    def e2e_run(k, boxes):
        """boxes=True: the step's result is the detection list (post-processing on the GPU, D2H = boxes);
        boxes=False: the head maps are read back (10.1 MB), as round 1 measured."""
        ev = lambda: torch.cuda.Event()
        copied, enc_done = [None] * R, [None] * R
        ego_done, d2h_done = [None] * INFLIGHT, [None] * INFLIGHT
        main = torch.cuda.current_stream()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main.synchronize()
        start.record(main)
        s_h2d.wait_event(start)
        for st in streams:
            st.wait_event(start)

        def upload(i):
            r = i % R
            with torch.cuda.stream(s_h2d):
                if enc_done[r] is not None:
                    s_h2d.wait_event(enc_done[r])          # the graph that read this buffer R steps ago is done
                for dst, src in zip(pil_pool[r], pil_host):
                    dst.copy_(src, non_blocking=True)
                copied[r] = ev()
                copied[r].record(s_h2d)

        upload(0)
        for i in range(k):
            r, sl = i % R, i % INFLIGHT
            if i + 1 < k:
                upload(i + 1)
            st = streams[sl]
            with torch.cuda.stream(st):
                st.wait_event(copied[r])
                p = step(i)
                enc_done[r] = ev()
                enc_done[r].record(st)
                if owns_result:
                    if d2h_done[sl] is not None:
                        st.wait_event(d2h_done[sl])
                    if boxes:
                        ppe[sl].forward_into(p, pp_out[sl])
                    else:
                        preds_stage[sl].copy_(p, non_blocking=True)
                ego_done[sl] = ev()
                ego_done[sl].record(st)
            if owns_result:
                with torch.cuda.stream(s_d2h):
                    s_d2h.wait_event(ego_done[sl])
                    if boxes:
                        for dst, src in zip(pp_host[sl], pp_out[sl]):
                            dst.copy_(src, non_blocking=True)
                    else:
                        preds_hosts[sl].copy_(preds_stage[sl], non_blocking=True)
                    d2h_done[sl] = ev()
                    d2h_done[sl].record(s_d2h)
        for st in streams:
            main.wait_stream(st)
        main.wait_stream(s_d2h)
        end.record(main)
        torch.cuda.synchronize()
        return start.elapsed_time(end)

    e2e_fps = {}
    for boxes in (True, False):
        e2e_run(3, boxes)
        sync_all()
        e2e_ms = e2e_run(args.steps, boxes)
        sync_all()
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_fps[boxes] = 1e3 * F / (float(t.item()) / args.steps)
    n_boxes = int(pp_host[0][4][0].item()) if rank == 0 else 0
    n_cand = int(pp_host[0][4][1].item()) if rank == 0 else 0

    # the frame's result as a checksum: integers travel between the GPUs and every output pixel is computed by the
    # same arithmetic whatever the tiling, so this must be identical for every --gpus N
    p = step(0)
    preds_sha1 = None
    if rank == 0:
        import hashlib
        preds_sha1 = hashlib.sha1(p.detach().cpu().numpy().tobytes()).hexdigest()

    # ---- roofline of the dominant kernel: the shrinker's first conv (3x3, 384 -> 256 over the 3-scale concat input,
    # igemm_kernel<128,128,3,RequantEpilogue<3>>: the largest layer, 31.1 of 75.3 GMAC per agent, and the largest
    # single share of the frame), timed alone over this rank's agents; the second shrinker conv is reported beside it.
    roof = None
    if rank == 0:
        from quantv2x_b200.engine import rowsum_u8
        reps = max(args.steps, 10)

        def time_layer(layer, cin, groups):
            xin = torch.randint(0, 256, (n_img, pipe.ho, pipe.wo, cin), dtype=torch.uint8, device=device)
            cg = cin // groups
            rs = [rowsum_u8(xin, i * cg, cg) for i in range(groups)]
            yout = torch.empty((n_img, pipe.ho, pipe.wo, 256), dtype=torch.uint8, device=device)
            for _ in range(3):
                layer.forward(xin, rowsum_in=rs, out=yout)
            torch.cuda.synchronize()
            return timed_loop(lambda: layer.forward(xin, rowsum_in=rs, out=yout), reps) / reps

        k0_ms = time_layer(pipe.fused.plan.layers[-2], 384, 3)
        k1_ms = time_layer(pipe.fused.plan.layers[-1], 256, 1)
        ach0 = 2.0 * SHRINK0_GMAC_PER_AGENT * 1e9 * n_img / (k0_ms * 1e-3) / 1e12
        ach1 = 2.0 * SHRINK1_GMAC_PER_AGENT * 1e9 * n_img / (k1_ms * 1e-3) / 1e12
        # int8 tensor-pipe peak, measured live two ways (MEASURED_PEAKS.json has bf16 only):
        #  (a) the raw tcgen05.mma kind::i8 rate by the library's own issue loop (qv2x_int8_mma_peak): the pipe's ceiling;
        #  (b) a library int8 GEMM (cuBLASLt through torch._int_mm, 8192^3, best of 10): what a tuned GEMM kernel reaches.
        # The roofline fraction is against (a), the larger; (b) and 2 x bf16 are reported beside it.
        import ctypes
        tops, mhz = ctypes.c_double(0.0), ctypes.c_double(0.0)
        _lib.check(_lib.lib().qv2x_int8_mma_peak(ctypes.byref(tops), ctypes.byref(mhz),
                                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        peak_raw = float(tops.value)
        a8 = torch.randint(-100, 100, (8192, 8192), dtype=torch.int8, device=device)
        b8 = torch.randint(-100, 100, (8192, 8192), dtype=torch.int8, device=device)
        for _ in range(3):
            torch._int_mm(a8, b8)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a8, b8)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        peak_lib = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
        del a8, b8
        peak = max(peak_raw, peak_lib)
        roof = {"bound": "tensor",
                "kernel": "igemm_kernel<128,128,3,FixedEpilogue<3>,HALO> (shrinker conv3x3 384->256 over the 3-scale concat)",
                "achieved": ach0, "peak": peak, "unit": "TOP/s (int8)", "frac": ach0 / peak,
                "peak_source": "live: raw tcgen05.mma kind::i8 issue loop of libqv2x (qv2x_int8_mma_peak, M128 N256 K32, "
                               f"operands resident in smem, SM clock {mhz.value:.0f} MHz); MEASURED_PEAKS.json has no int8 entry",
                "peak_cublaslt_int8": peak_lib, "frac_vs_cublaslt": ach0 / peak_lib,
                # dram__bytes_read + dram__bytes_write of this kernel from one `ncu --set full` capture
                # (profiles/r2_ncu_shrink0.txt: 4 agents), per agent, scaled to this rank's agent count; algorithmic:
                # 13.5 MB in + 9 MB out + 0.9 MB weights per agent (the output mostly stays in L2 for the next layer)
                "traffic": NCU_SHRINK0_DRAM_BYTES_PER_AGENT * n_img, "traffic_unit": "bytes per launch (ncu, per agent x agents)",
                "us_per_launch": k0_ms * 1e3,
                "other_kernels": [{"kernel": "igemm_kernel<256,128,1,FixedEpilogueC<1>,HALO> (shrinker conv3x3 256->256)",
                                   "achieved": ach1, "frac": ach1 / peak, "frac_vs_cublaslt": ach1 / peak_lib,
                                   "us_per_launch": k1_ms * 1e3}],
                "step_tensor_frac": 2.0 * GMAC_INT8_PER_AGENT * 1e9 * n_img / (ms_per_step * 1e-3) / 1e12 / peak,
                "step_tensor_frac_vs_cublaslt": 2.0 * GMAC_INT8_PER_AGENT * 1e9 * n_img / (ms_per_step * 1e-3) / 1e12 / peak_lib}
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        if os.path.exists(peaks_file):
            try:
                pk = json.load(open(peaks_file))
                roof["bf16_tflops_measured"] = pk.get("bf16_tflops")
                if pk.get("hbm_gbs"):
                    hbm_peak, hbm_src = float(pk["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy bandwidth)"
            except Exception:
                pass
        # HBM-bound kernels of the ego stage (whole map, all agents), timed alone on cold inputs: the decode gather
        # writes N x 36 MB of features from 3 code bytes per row; the fuse kernel reads them once and writes 36 MB
        from quantv2x_b200 import engine as E
        n_all = N_AGENTS
        codes_all = torch.randint(0, pipe.codebook.k[0], (levels, m, n_all * hw), dtype=torch.uint8, device=device)
        eb = pipe.ego_buffers(n_all, slot=99)
        cold = torch.empty(256 << 20, dtype=torch.uint8, device=device)

        def time_cold(fn, reps=10):
            tot = 0.0
            for _ in range(reps):
                cold.fill_(1)                       # evict the inputs from L2
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            return tot / reps

        feat_flat = eb["feat"].view(n_all * hw, pipe.c_feat)
        for _ in range(2):
            pipe.codebook.decode(codes_all, out=feat_flat)
            E.fuse(eb["feat"], aff, pipe.fusion_mode, out=eb["fused"])
        dec_ms = time_cold(lambda: pipe.codebook.decode(codes_all, out=feat_flat))
        fuse_ms = time_cold(lambda: E.fuse(eb["feat"], aff, pipe.fusion_mode, out=eb["fused"]))
        row_bytes = pipe.c_feat * 4
        dec_bytes = n_all * hw * (row_bytes + levels * m)
        fuse_bytes = (n_all + 1) * hw * row_bytes
        # the ego stage's ALGORITHMIC traffic is the codes in (N x 105.6 KB), the tables once (L2) and
        # the fused map out (36.04 MB) -- a fully fused decode+warp+fuse kernel would move no more.  The two kernels
        # of this build materialise the N decoded maps in between; `moved_bytes` is what they actually move.
        alg_bytes = n_all * hw * levels * m + levels * m * pipe.codebook.k[0] * row_bytes + hw * row_bytes
        ego_ms = dec_ms + fuse_ms
        chain = [
            {"kernel": f"three-kernel chain, first two: codebook_decode_kernel + fuse_kernel ({pipe.fusion_mode}, {n_all} agents)",
             "bound": "hbm", "algorithmic_bytes": alg_bytes, "achieved": alg_bytes / (ego_ms * 1e-3) / 1e9,
             "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / (ego_ms * 1e-3) / 1e9 / hbm_peak,
             "us_per_launch": ego_ms * 1e3},
            {"kernel": "codebook_decode_kernel (table gather from shared memory)", "bound": "hbm",
             "moved_bytes": dec_bytes, "moved_gbs": dec_bytes / (dec_ms * 1e-3) / 1e9,
             "moved_frac": dec_bytes / (dec_ms * 1e-3) / 1e9 / hbm_peak, "us_per_launch": dec_ms * 1e3},
            {"kernel": f"fuse_kernel ({pipe.fusion_mode}, warp + fuse, {n_all} agents)", "bound": "hbm",
             "moved_bytes": fuse_bytes, "moved_gbs": fuse_bytes / (fuse_ms * 1e-3) / 1e9,
             "moved_frac": fuse_bytes / (fuse_ms * 1e-3) / 1e9 / hbm_peak, "us_per_launch": fuse_ms * 1e3}]
        if pipe.ego_att is not None:
            # The ego stage the frame actually runs: ONE kernel from the code planes to the head maps over the folded
            # tables (csrc/egostage.cu).  Its HBM traffic is the algorithmic minimum of the stage WITH the heads --
            # codes in, Gram matrix + head table once, 72 head maps out -- and it is bound by the shared-memory pipe
            # (one 288-byte head-table row per agent, tap and code plane), not by HBM: `frac` is reported against the
            # HBM peak for continuity with round 1, `smem_*` against the 128 B/clk/SM shared-memory port.
            out_all = torch.empty((pipe.heads.cout, hw), dtype=torch.float32, device=device)
            for _ in range(2):
                pipe.ego_att.forward(codes_all, aff, n_all, pipe.ho, pipe.wo, out=out_all)
            one_ms = time_cold(lambda: pipe.ego_att.forward(codes_all, aff, n_all, pipe.ho, pipe.wo, out=out_all))
            rows_r = sum(pipe.codebook.k) * m
            one_bytes = n_all * hw * levels * m + rows_r * rows_r * 4 + rows_r * 72 * 4 + pipe.heads.cout * hw * 4
            # shared-memory bytes the kernel must read per pixel: (4 taps x (n-1) agents + 1 ego tap) x planes x 288 B
            smem_bytes = hw * ((4 * (n_all - 1) + 1) * levels * m) * 72 * 4
            smem_peak = (128.0 * torch.cuda.get_device_properties(device).multi_processor_count *
                         (summarize_clocks(samples)["sm_mhz"] or 1965) * 1e6 / 1e9)
            roof["hbm_kernels"] = [
                {"kernel": f"ego_att_kernel<{levels * m}> (decode . warp . attention fusion . heads in one kernel, "
                           f"{n_all} agents; replaces decode + fuse + heads)",
                 "bound": "shared-memory pipe (HBM traffic is at the algorithmic minimum)",
                 "algorithmic_bytes": one_bytes, "achieved": one_bytes / (one_ms * 1e-3) / 1e9, "peak": hbm_peak,
                 "unit": "GB/s", "frac": one_bytes / (one_ms * 1e-3) / 1e9 / hbm_peak, "us_per_launch": one_ms * 1e3,
                 "survey_8d_bytes_without_heads": alg_bytes,
                 "smem_row_bytes": smem_bytes, "smem_gbs": smem_bytes / (one_ms * 1e-3) / 1e9,
                 "smem_peak_gbs": smem_peak, "smem_frac": smem_bytes / (one_ms * 1e-3) / 1e9 / smem_peak,
                 "ncu": "profiles/r2_ncu_ego_att.txt"}] + chain
            del out_all
        else:
            roof["hbm_kernels"] = chain
        roof["hbm_peak_source"] = hbm_src
        del cold, codes_all

    if rank == 0:
        stop.set()
        th.join(timeout=2)

    # ---- parity of this very build at the benchmarked shape, outside the timed region (N = 1 only): one agent's
    # 200 x 704 frame through the plan and the encoder against the oracle (features and codes bit for bit)
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import codebook_oracle as co
        from oracle import int_oracle
        spec = export_spec(q, bev_delta)
        one = tuple(t[: PILLARS] if t.dim() == 1 else t[: PILLARS] for t in pil_pool[0])
        bev1 = pillar.forward(*one, 1)
        codes1 = pipe.encode_agents(bev1).cpu().numpy()
        feat1 = pipe.encode_buffers(1)["feat"].cpu().numpy()
        _, feat_ref = int_oracle.backbone_chain(spec, bev1.cpu().numpy())
        parity = {"checked": "agent 0 of frame 0 at 200x704: uint8 features of the 24-layer plan vs "
                             "oracle/int_oracle.backbone_chain (bit for bit)",
                  "features_bit_exact": bool(np.array_equal(feat1, feat_ref)),
                  "features_sha1": __import__("hashlib").sha1(feat1.tobytes()).hexdigest(),
                  "oracle_features_sha1": __import__("hashlib").sha1(feat_ref.tobytes()).hexdigest(),
                  "codes_sha1": __import__("hashlib").sha1(codes1.tobytes()).hexdigest()}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from quantv2x_b200.export import pillar_spec
        pspec = pillar_spec(q.model.encoder_m1)
        v, cores, sample, _, _ = cpu_reference_frames_per_s(spec, pspec, enc_args, N_AGENTS, 3, 1, budget_s=60.0)
        cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        h2d = in_bytes * world
        n_results = F                               # results read back per step over all ranks
        line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong" if MGPU == "tiles" else "weak", "vs_baseline": None, "dtype": "u8 (int8 tensor cores, int32 accumulate)",
                "data": "synthetic", "preds_sha1": preds_sha1, "parity": parity, "config": config,
                "e2e": {"value": e2e_fps[True], "unit": "frames/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": int(box_bytes) * n_results,
                        "result": f"detections after GPU post-processing (score threshold, box decode, rotated NMS): "
                                  f"{n_cand} candidates above the threshold ranked and run through the rotated-IoU "
                                  f"matrix, {n_boxes} boxes kept in the last frame; buffers of top-1000 boxes are read "
                                  "back; "
                                  f"score threshold {pp_threshold:.4f} = the 300th highest score of frame 0 "
                                  "(random-init heads never reach the yaml's 0.2)"},
                "e2e_head_maps": {"value": e2e_fps[False], "unit": "frames/s", "h2d_bytes_per_step": h2d,
                                  "d2h_bytes_per_step": int(preds_host.numel() * 4) * n_results},
                "gpu_launches": int(launches), "clocks": summarize_clocks(samples),
                "phases_ms_rank0": phases,
                "latency_ms_one_step_at_a_time": serial_ms / args.steps, "roofline": roof,
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
