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


def allgather_csr(offsets: torch.Tensor, hits: torch.Tensor, total: int | None = None, group=None):
    """offsets: [n_local+1] local CSR offsets (offsets[-1] == number of local hits), hits: [>= n_hits] local hit
    indices.  Returns (global_offsets int64 [n_global+1], global_hits [H_global]) in rank order == ray order.

    One host synchronisation (the gathered totals size the hit all-gather); the local total is read from
    offsets[-1] on the device, so the caller does not need it on the host."""
    world = dist.get_world_size(group)
    dev = offsets.device
    n_local = offsets.numel() - 1
    meta = torch.empty(2, dtype=torch.int64, device=dev)
    meta[0] = n_local
    meta[1] = offsets[-1] if total is None else int(total)
    all_meta = torch.empty(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_meta, meta, group=group)
    all_meta = all_meta.view(world, 2)
    host = all_meta.cpu()                                   # the one sync of the exchange step
    n_each, h_each = host[:, 0].tolist(), host[:, 1].tolist()
    max_n, max_h = max(n_each), max(max(h_each), 1)
    # per-rank offsets (padded only if the shards are uneven)
    if min(n_each) == max_n:
        off_in = offsets
    else:
        off_in = torch.zeros(max_n + 1, dtype=offsets.dtype, device=dev)
        off_in[: n_local + 1] = offsets
    all_off = torch.empty(world * (max_n + 1), dtype=offsets.dtype, device=dev)
    dist.all_gather_into_tensor(all_off, off_in, group=group)
    # hit lists: every rank sends its first max_h entries (entries beyond its own total are ignored by the receiver)
    if hits.numel() >= max_h:
        hit_in = hits[:max_h]
    else:
        hit_in = torch.zeros(max_h, dtype=hits.dtype, device=dev)
        hit_in[: hits.numel()] = hits
    all_hits = torch.empty(world * max_h, dtype=hits.dtype, device=dev)
    dist.all_gather_into_tensor(all_hits, hit_in.contiguous(), group=group)
    # assemble the global CSR: rebase each rank's offsets by the hits of the ranks before it
    base = torch.cumsum(all_meta[:, 1], 0) - all_meta[:, 1]
    g_off = all_off.view(world, max_n + 1)[:, :max_n].to(torch.int64) + base[:, None]
    if min(n_each) == max_n:
        g_off = g_off.reshape(-1)
    else:
        g_off = torch.cat([g_off[r, : n_each[r]] for r in range(world)])
    g_off = torch.cat([g_off, all_meta[:, 1].sum().reshape(1)])
    g_hits = torch.cat([all_hits[r * max_h: r * max_h + h_each[r]] for r in range(world)])
    return g_off, g_hits


class ShardedTraversal:
    """The multi-GPU step with the exchange fused into the traversal (bvhgpu_traverse_sharded_dev_*): every rank ends the step
    with its own copy of the global CSR.  After its walk a rank pushes its per-ray hit COUNTS (1 byte per ray unless a ray has
    more than 255 hits) into all peers' staging buffers and publishes its hit total in their mailboxes; every rank rebuilds the
    global offsets with a local scan; the emit kernel stores the hit lists straight into every rank's hit buffer (P2P stores over
    NVLink).  No NCCL call and no host synchronisation on the data path; torch.distributed is used once, at construction, to
    swap the CUDA IPC handles."""

    def __init__(self, bvh, nrays_local: int, cap: int, group=None, ray_layout: int = 0):
        import ctypes as C

        import numpy as np

        from . import capi

        self.bvh, self.capi, self.C, self.np = bvh, capi, C, np
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > capi.MAX_PEERS:
            raise ValueError(f"at most {capi.MAX_PEERS} ranks")
        sizes = [None] * self.world
        dist.all_gather_object(sizes, int(nrays_local), group=group)
        if min(sizes) <= 0:
            raise ValueError("every rank needs a non-empty shard")
        self.n_each = sizes
        self.nrays_global = sum(sizes)
        self.rays_before = sum(sizes[: self.rank])
        self.cap = int(cap)
        L, ctx = capi.lib(), bvh.ctx._h
        self._own, handles = [], []
        # peer-mapped: count staging, global hit lists, mailbox; local only: the global offsets
        for nbytes in (2 * capi.shard_stage_bytes(self.nrays_global), 4 * self.cap, capi.MAILBOX_BYTES, 4 * (self.nrays_global + 1)):
            ptr, h = C.c_void_p(), (C.c_ubyte * capi.IPC_HANDLE_BYTES)()
            capi.check(L.bvhgpu_peer_alloc(ctx, nbytes, C.byref(ptr), h))
            self._own.append(ptr)
            handles.append(bytes(h))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, handles[:3], group=group)
        self._opened = []
        self.shard = capi.Shard()
        self.shard.rank, self.shard.world = self.rank, self.world
        self.shard.cap, self.shard.seq, self.shard.ray_layout = self.cap, 0, int(ray_layout)
        self.shard.offsets = self._own[3].value
        for r in range(self.world):
            self.shard.shard_rays[r] = sizes[r]
            ptrs = []
            for k in range(3):
                if r == self.rank:
                    ptrs.append(self._own[k].value)
                else:
                    p = C.c_void_p()
                    hb = (C.c_ubyte * capi.IPC_HANDLE_BYTES).from_buffer_copy(everyone[r][k])
                    capi.check(L.bvhgpu_peer_open(ctx, hb, C.byref(p)))
                    self._opened.append(p)
                    ptrs.append(p.value)
            self.shard.peer_counts[r], self.shard.peer_hits[r], self.shard.peer_mailbox[r] = ptrs
        dist.barrier(group=group)
        self._fn = getattr(L, f"bvhgpu_traverse_sharded_dev_{bvh._d['suffix']}")

    def step(self, rays_ptr: int, nrays: int, mode: int = 0):
        """Enqueue one sharded traversal (asynchronous); when the stream reaches the end of it this rank's copy of
        the global CSR is complete (all peers have signalled that their stores landed)."""
        self.shard.seq += 1
        self.capi.check(self._fn(self.bvh._h, mode, self.C.c_void_p(rays_ptr), nrays, self.C.byref(self.shard)))

    def step_host(self, host_rays_ptr: int, dev_rays_ptr: int, nbytes: int, nrays: int, mode: int = 0):
        """The step with this rank's shard still in (pinned) host memory: H2D on the context's stream, then step()."""
        self.capi.check(self.capi.lib().bvhgpu_memcpy_h2d_async(self.bvh.ctx._h, self.C.c_void_p(dev_rays_ptr), self.C.c_void_p(host_rays_ptr), nbytes))
        self.step(dev_rays_ptr, nrays, mode)

    def fetch(self, offsets_out=None, hits_out=None):
        """Global CSR as numpy arrays (offsets u32[n_global+1], hits u32[total]); synchronises and raises if a peer timed out
        (bvhgpu_synchronize), if the u32 offsets overflowed or if the hit lists did not fit `cap`."""
        np, C, L = self.np, self.C, self.capi.lib()
        self.bvh.ctx.synchronize()
        off = offsets_out if offsets_out is not None else np.empty(self.nrays_global + 1, dtype=np.uint32)
        self.capi.check(L.bvhgpu_memcpy_d2h(self.bvh.ctx._h, off.ctypes.data_as(C.c_void_p), self._own[3], 4 * (self.nrays_global + 1)))
        total = int(off[self.nrays_global])
        if total == 0xFFFFFFFF:
            raise self.capi.BvhGpuError(self.capi.ERR_CAPACITY, "sharded traversal: the hit total overflows the u32 CSR offsets")
        if total > self.cap:
            raise self.capi.BvhGpuError(self.capi.ERR_CAPACITY, f"sharded traversal: {total} hits do not fit the global hit buffers (cap {self.cap})")
        hits = hits_out[:total] if hits_out is not None else np.empty(total, dtype=np.uint32)
        self.capi.check(L.bvhgpu_memcpy_d2h(self.bvh.ctx._h, hits.ctypes.data_as(C.c_void_p), self._own[1], 4 * total))
        return off, hits

    def trace(self):
        """Per-step exchange trace of this rank (diagnostics): dict seq -> (ns waited for the peers' totals, ns waited for their
        done flags).  Long waits on one rank mean ANOTHER rank was late; the rank that never waits is the straggler."""
        np, C, L = self.np, self.C, self.capi.lib()
        box = np.empty(self.capi.MAILBOX_BYTES // 8, dtype=np.uint64)
        self.capi.check(L.bvhgpu_memcpy_d2h(self.bvh.ctx._h, box.ctypes.data_as(C.c_void_p), self._own[2], box.nbytes))
        tr = box[self.capi.MB_TRACE_WORD: self.capi.MB_TRACE_WORD + 4 * self.capi.MB_TRACE_LEN].reshape(-1, 4)
        return {int(r[0]): (int(r[2]), int(r[3])) for r in tr if r[0] != 0}

    def close(self):
        L, ctx = self.capi.lib(), self.bvh.ctx._h
        try:
            self.bvh.ctx.synchronize()
        except self.capi.BvhGpuError:
            pass
        dist.barrier(group=self.group)
        for p in self._opened:
            L.bvhgpu_peer_close(ctx, p)
        dist.barrier(group=self.group)
        for p in self._own:
            L.bvhgpu_peer_free(ctx, p)
        self._opened, self._own = [], []
