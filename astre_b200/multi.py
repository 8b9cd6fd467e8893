"""Multi-GPU orchestration: one process per GPU, entities sharded by contiguous range (flat scenes) or
by whole worlds (batched worlds); the only exchange on the path is an all-gather of per-shard visible
counts, from which every rank derives the offset of its segment in each global draw list
(SURVEY.md 8e).  Keys never leave their GPU.  Works over NCCL (device tensors) and gloo (CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist

from . import gpu as _gpu


def shard_range(n, rank, world_size, align=4):
    """Contiguous entity range [lo, hi) of `rank`; boundaries are multiples of `align`."""
    per = -(-n // world_size)
    per = (per + align - 1) // align * align
    return min(rank * per, n), min((rank + 1) * per, n)


def shard_worlds(n_worlds, rank, world_size):
    """Whole worlds [lo, hi) of `rank` (batched independent worlds)."""
    per = -(-n_worlds // world_size)
    return min(rank * per, n_worlds), min((rank + 1) * per, n_worlds)


def exchange_counts(counts, group=None):
    """All-gather of this shard's per-view visible counts.  `counts`: 1-D int32 tensor (device tensor
    under NCCL).  Returns a [world_size, n_views] tensor on the same device."""
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    if ws == 1:
        return counts.reshape(1, -1)
    out = torch.empty(ws * counts.numel(), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(out, counts.contiguous(), group=group)
    return out.reshape(ws, -1)


def segment_offsets(all_counts, rank):
    """Offset of `rank`'s segment of every view in the global draw list, which is ordered
    (view, rank, key), and the per-view totals."""
    a = all_counts.detach().cpu().numpy().astype(np.uint32)
    return _gpu.global_offsets(a, rank)
