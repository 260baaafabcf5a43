"""Multi-GPU use of the hot path: independent replicas (SURVEY.md 8e -- the path does not shard).

One process per GPU, each holding a full weight copy and its own streams' windows; the only
collective is one broadcast of the weights from ``src`` at init (NCCL over NVLink/NVSwitch on the
GPU box, gloo in the CPU tests).  No per-step traffic.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def flatten_params(model) -> torch.Tensor:
    """All parameters in state-dict order as ONE contiguous fp32 buffer (14,709,260 B for the
    shipped architecture) so the broadcast is a single collective launch."""
    ps = model._ordered_params()
    return torch.cat([p.detach().reshape(-1) for p in ps])


def unflatten_into(model, flat: torch.Tensor) -> None:
    off = 0
    with torch.no_grad():
        for p in model._ordered_params():
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))     # bumps _version -> the C side re-packs
            off += n
    assert off == flat.numel()


def broadcast_weights(model, src: int = 0, group=None) -> int:
    """Broadcast ``model``'s weights from rank ``src`` to every rank; returns bytes moved."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    flat = flatten_params(model)
    dist.broadcast(flat, src=src, group=group)
    if dist.get_rank(group) != src:
        unflatten_into(model, flat)
    return flat.numel() * 4


def stream_owner(stream_index: int, world_size: int) -> int:
    """Stream / batch index -> owning rank (SURVEY.md 8e partitioning)."""
    return stream_index % world_size


def local_streams(n_streams: int, rank: int, world_size: int):
    return [i for i in range(n_streams) if stream_owner(i, world_size) == rank]
