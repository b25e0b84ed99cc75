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
