"""Multi-GPU ray sharding: the tree is replicated (every rank builds the same tree from
the same AABBs -- the builder is deterministic -- or receives it by broadcast), the ray batch is cut
into contiguous shards, every rank traverses its shard, and the per-rank CSR hit lists are
all-gathered over NCCL (NVLink 5 / NVSwitch) into the global CSR in original ray order.

The only exchange step of the path is this all-gather; there is no collective inside the traversal.
The functions work on torch tensors of any device so that the host logic is testable with gloo.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank` (first shards get the remainder)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_csr(offsets: torch.Tensor, hits: torch.Tensor, total: int, group=None):
    """offsets: int64/int32 [n_local+1] local CSR offsets, hits: [>= total] local hit indices.
    Returns (global_offsets int64 [n_global+1], global_hits [H_global]) in rank order == ray order."""
    world = dist.get_world_size(group)
    dev = offsets.device
    n_local = offsets.numel() - 1
    meta = torch.tensor([n_local, int(total)], dtype=torch.int64, device=dev)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    metas = torch.stack(metas).cpu()
    n_each, h_each = metas[:, 0].tolist(), metas[:, 1].tolist()
    max_n, max_h = max(n_each), max(max(h_each), 1)
    # padded all-gathers (one for the counts, one for the hit indices)
    counts = torch.zeros(max_n, dtype=torch.int64, device=dev)
    counts[:n_local] = (offsets[1:] - offsets[:-1]).to(torch.int64)
    all_counts = torch.empty(world * max_n, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_counts, counts, group=group)
    pad_hits = torch.zeros(max_h, dtype=hits.dtype, device=dev)
    pad_hits[: int(total)] = hits[: int(total)]
    all_hits = torch.empty(world * max_h, dtype=hits.dtype, device=dev)
    dist.all_gather_into_tensor(all_hits, pad_hits, group=group)
    g_counts = torch.cat([all_counts[r * max_n: r * max_n + n_each[r]] for r in range(world)])
    g_hits = torch.cat([all_hits[r * max_h: r * max_h + h_each[r]] for r in range(world)])
    g_off = torch.zeros(g_counts.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(g_counts, 0, out=g_off[1:])
    return g_off, g_hits
