"""Multi-GPU use of the lane-NMS op: frames are independent, so a batch is split into contiguous blocks of frames,
one block per rank (one process per GPU), with NO collective on the data path.  The only exchange is the final
collection of the kept-lane results: one all-gather of a compact [frames, top_k + 1] int64 tensor (kept indices
+ count) over NCCL/NVLink.  (The reference shards whole videos across ranks with a DistributedSampler and never
gathers, testOLV3.py:33-40.)
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(F: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of frames owned by `rank`: the first F % world ranks get one extra frame."""
    if world <= 0 or not (0 <= rank < world) or F < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(F, world)
    f0 = rank * base + min(rank, rem)
    return f0, f0 + base + (1 if rank < rem else 0)


def pack_kept(keep: torch.Tensor, num: torch.Tensor, top_k: int) -> torch.Tensor:
    """[F, N] keep + [F] num  ->  compact [F, top_k + 1] (kept indices zero padded, count in the last column)."""
    F = keep.shape[0]
    k = min(top_k, keep.shape[1])
    out = torch.zeros((F, top_k + 1), dtype=torch.int64, device=keep.device)
    out[:, :k] = keep[:, :k]
    out[:, top_k] = num
    return out


def unpack_kept(packed: torch.Tensor):
    return packed[:, :-1], packed[:, -1]


def gather_kept(packed_local: torch.Tensor, F_total: int, group=None) -> torch.Tensor:
    """All-gather the per-rank compact results into frame order.  Every rank passes its block from
    `shard_range`; blocks may differ by one frame, so they are padded to the largest block for the collective."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return packed_local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    width = packed_local.shape[1]
    block = (F_total + world - 1) // world
    send = packed_local
    if send.shape[0] != block:
        send = torch.zeros((block, width), dtype=packed_local.dtype, device=packed_local.device)
        send[: packed_local.shape[0]] = packed_local
    recv = torch.empty((world * block, width), dtype=packed_local.dtype, device=packed_local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    parts = []
    for r in range(world):
        f0, f1 = shard_range(F_total, r, world)
        parts.append(recv[r * block: r * block + (f1 - f0)])
    return torch.cat(parts, dim=0)
