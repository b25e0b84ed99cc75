"""world_size-2 gloo test (CPU) of the N>1 host logic: agent sharding + code-plane gather ordering."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_agents, hw, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quantv2x_b200.distributed import gather_code_planes, shard_agents

    levels, m = 3, 2
    rng = np.random.default_rng(0)
    full = rng.integers(0, 256, size=(levels, m, n_agents * hw), dtype=np.uint8)      # the 1-process result
    mine = shard_agents(n_agents, world, rank)
    local = torch.from_numpy(np.ascontiguousarray(full[:, :, mine.start * hw:mine.stop * hw]))
    out = gather_code_planes(local, hw)
    if rank == 0:
        ret["ok"] = bool(np.array_equal(out.numpy(), full))
    else:
        ret[f"none{rank}"] = out is None
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process():
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, 8, 37, ret), nprocs=2, join=True)
        assert ret["ok"] and ret["none1"]


def test_shard_agents():
    from quantv2x_b200.distributed import shard_agents

    assert list(shard_agents(8, 4, 1)) == [2, 3]
    assert [list(shard_agents(8, 8, r)) for r in range(8)] == [[r] for r in range(8)]
    try:
        shard_agents(8, 3, 0)
        assert False
    except ValueError:
        pass


def _worker_tiles(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quantv2x_b200.distributed import all_gather_code_planes, gather_pred_tiles, rank_tile

    h, w, cout, hw_codes = 4, 6, 3, 5
    full = np.arange(cout * h * w, dtype=np.float32).reshape(cout, h, w)
    y0, y1, x0, x1 = rank_tile(rank, world, h, w)
    tile = torch.from_numpy(np.ascontiguousarray(full[:, y0:y1, x0:x1]).reshape(cout, -1))
    out = gather_pred_tiles(tile, h, w)
    codes_full = np.arange(2 * 1 * world * hw_codes, dtype=np.uint8).reshape(2, 1, world * hw_codes)
    mine = torch.from_numpy(np.ascontiguousarray(codes_full[:, :, rank * hw_codes:(rank + 1) * hw_codes]))
    allc = all_gather_code_planes(mine, hw_codes)
    ok = bool(np.array_equal(allc.numpy(), codes_full))
    if rank == 0:
        ok = ok and bool(np.array_equal(out.numpy(), full.reshape(cout, -1)))
    ret[rank] = ok
    dist.destroy_process_group()


def test_two_rank_tiles_and_allgather():
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_tiles, args=(2, port, ret), nprocs=2, join=True)
        assert ret[0] and ret[1]


def _worker_a2a(rank, world, port, n_agents, hw, ret):
    """Frame-batched serving: rank r encodes ITS agents for `world` frames; frame f is fused on rank f."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quantv2x_b200.distributed import all_to_all_code_planes, shard_agents

    levels, m = 3, 2
    rng = np.random.default_rng(1)
    frames = rng.integers(0, 256, size=(world, levels, m, n_agents * hw), dtype=np.uint8)   # frame f: all agents' codes
    mine = shard_agents(n_agents, world, rank)
    local = np.concatenate([frames[f][:, :, mine.start * hw:mine.stop * hw] for f in range(world)], axis=2)
    out = all_to_all_code_planes(torch.from_numpy(np.ascontiguousarray(local)))
    ret[f"ok{rank}"] = bool(np.array_equal(out.numpy(), frames[rank]))
    dist.destroy_process_group()


def test_two_rank_all_to_all_gives_each_rank_its_frame():
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_a2a, args=(2, port, 8, 48, ret), nprocs=2, join=True)
        assert ret["ok0"] and ret["ok1"]
