"""Agent-per-GPU sharding of the cooperative frame (new design -- the reference has no
multi-GPU inference).  One process per GPU; agents are split contiguously over the ranks; the only exchange
is a gather of uint8 code planes (levels*m bytes per BEV cell per agent) to the ego rank.  Integers on the
wire make the G-GPU result bit-identical to the 1-GPU result."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_agents(n_agents: int, world: int, rank: int) -> range:
    """Contiguous block of agents owned by `rank` (agent 0 = ego lives on rank 0)."""
    if n_agents % world != 0:
        raise ValueError(f"{n_agents} agents do not divide over {world} ranks")
    per = n_agents // world
    return range(rank * per, (rank + 1) * per)


def gather_code_planes(codes: torch.Tensor, hw: int, dst: int = 0, recv: torch.Tensor | None = None):
    """codes: uint8 [levels, m, per*hw] of this rank's agents (agent-major rows).
    Returns uint8 [levels, m, N*hw] with all agents in global order on rank `dst`, None elsewhere.
    recv: optional preallocated [world, levels, m, per*hw] buffer on `dst` (keeps the step allocation-free)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return codes
    world, rank = dist.get_world_size(), dist.get_rank()
    levels, m, rows = codes.shape
    assert rows % hw == 0
    if rank == dst:
        if recv is None:
            recv = torch.empty((world, levels, m, rows), dtype=codes.dtype, device=codes.device)
        dist.gather(codes, [recv[i] for i in range(world)], dst=dst)
        return recv.permute(1, 2, 0, 3).reshape(levels, m, world * rows).contiguous()
    dist.gather(codes, None, dst=dst)
    return None


def tile_grid(world: int):
    """How the ego output map is cut over `world` ranks: (rows, cols) of tiles."""
    return {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}[world]


def rank_tile(rank: int, world: int, h: int, w: int):
    """(y0, y1, x0, x1) of the output tile owned by `rank`."""
    gy, gx = tile_grid(world)
    assert h % gy == 0 and w % gx == 0, "feature map must divide over the tile grid"
    ty, tx = rank // gx, rank % gx
    th, tw = h // gy, w // gx
    return (ty * th, (ty + 1) * th, tx * tw, (tx + 1) * tw)


def all_gather_code_planes(codes: torch.Tensor, hw: int, recv: torch.Tensor | None = None) -> torch.Tensor:
    """Every rank ends up with all agents' code planes [levels, m, N*hw] (agent-major)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return codes
    world = dist.get_world_size()
    levels, m, rows = codes.shape
    if recv is None:
        recv = torch.empty((world, levels, m, rows), dtype=codes.dtype, device=codes.device)
    dist.all_gather_into_tensor(recv.view(-1), codes.reshape(-1))
    return recv.permute(1, 2, 0, 3).reshape(levels, m, world * rows).contiguous()


def all_to_all_code_planes(codes: torch.Tensor, recv: torch.Tensor | None = None) -> torch.Tensor:
    """Frame-batched serving: this rank holds the codes of ITS agents for `world` consecutive frames, uint8
    [levels, m, world * rows] with frame-major rows; frame f is fused on rank f.  Returns the codes of ALL agents for
    this rank's frame, uint8 [levels, m, world * rows] with rank-major (= global agent order) rows.  The collective
    form of qv2x_scatter_planes (used when peer-mapped memory is unavailable, and on CPU in the tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return codes
    world = dist.get_world_size()
    levels, m, rows_all = codes.shape
    assert rows_all % world == 0
    rows = rows_all // world
    send = codes.view(levels, m, world, rows).permute(2, 0, 1, 3).contiguous()            # [frame, levels, m, rows]
    if recv is None:
        recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(-1), send.view(-1))
    return recv.permute(1, 2, 0, 3).reshape(levels, m, world * rows).contiguous()        # [levels, m, rank-major rows]


def gather_pred_tiles(preds_tile: torch.Tensor, h: int, w: int, dst: int = 0, recv: torch.Tensor | None = None):
    """preds_tile: [Cout, tile_pixels] of this rank's tile.  Returns [Cout, h*w] on `dst`, None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return preds_tile
    world, rank = dist.get_world_size(), dist.get_rank()
    cout = preds_tile.shape[0]
    if rank != dst:
        dist.gather(preds_tile, None, dst=dst)
        return None
    if recv is None:
        recv = torch.empty((world,) + tuple(preds_tile.shape), dtype=preds_tile.dtype, device=preds_tile.device)
    dist.gather(preds_tile, [recv[i] for i in range(world)], dst=dst)
    gy, gx = tile_grid(world)
    th, tw = h // gy, w // gx
    full = recv.view(gy, gx, cout, th, tw).permute(2, 0, 3, 1, 4).reshape(cout, h * w)
    return full.contiguous()


class PeerExchange:
    """Peer-mapped (symmetric) buffers for the multi-GPU frame: every rank's kernels store their code planes straight
    into every peer's code buffer and their tile of the head maps straight into the ego rank's result, so the step has
    no collective -- only two device-side barriers.  torch's symmetric-memory allocator provides the mapping and the
    barrier (plumbing); the stores are libqv2x kernels (qv2x_push_planes, qv2x_heads_forward_tile).

    One buffer set per in-flight slot: [levels*m*rows_total bytes of codes | cout*hw float32 head maps]."""

    def __init__(self, device, levels: int, m: int, rows_total: int, cout: int, hw: int, slots: int = 1):
        import torch.distributed._symmetric_memory as symm_mem

        group = dist.group.WORLD
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.levels, self.m, self.rows_total, self.cout, self.hw = levels, m, rows_total, cout, hw
        self.code_bytes = (levels * m * rows_total + 255) // 256 * 256
        self.slot_bytes = self.code_bytes + cout * hw * 4
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
        self.buf = symm_mem.empty(slots * self.slot_bytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(self.ptrs) == self.world and self.ptrs[self.rank] == self.buf.data_ptr()

    def codes_full(self, slot: int) -> torch.Tensor:
        """This rank's copy of ALL agents' code planes, uint8 [levels, m, rows_total]."""
        o = slot * self.slot_bytes
        return self.buf[o:o + self.levels * self.m * self.rows_total].view(self.levels, self.m, self.rows_total)

    def preds_full(self, slot: int) -> torch.Tensor:
        """This rank's head-map buffer float32 [cout, hw] (complete on the ego rank after the second barrier)."""
        o = slot * self.slot_bytes + self.code_bytes
        return self.buf[o:o + self.cout * self.hw * 4].view(torch.float32).view(self.cout, self.hw)

    def code_ptrs(self, slot: int):
        return [p + slot * self.slot_bytes for p in self.ptrs]

    def preds_ptr(self, rank: int, slot: int) -> int:
        return self.ptrs[rank] + slot * self.slot_bytes + self.code_bytes

    def barrier(self, slot: int):
        """Device-side barrier of all ranks on the current stream (one signal channel per slot)."""
        self.hdl.barrier(channel=slot)
