"""Multi-GPU plumbing for the ray path: one process per GPU over torch.distributed.

Rays are independent and the model (<= 70 MB) is replicated, so rendering shards rays with NO
data-path collective; training adds exactly one exchange step, a sum/avg all-reduce of
the flat packed gradient buffer (12 grids + basis + MLP) over NCCL / NVLink before the host optimiser
sees the gradients.  The reference (tensorf-myc) is single-process; its SimpleSampler (train.py:25-37)
is seeded identically everywhere, so ranks slice one global permutation.
The helpers are backend-agnostic (gloo on CPU tensors in the tests, nccl on the GPU box).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None, device=None):
    """Rendezvous from RANK/WORLD_SIZE/MASTER_* (torchrun).  Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced [start, end) of n items for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_rays(rays, rank: int, world: int):
    s, e = shard_bounds(rays.shape[0], rank, world)
    return rays[s:e]


def render_sharded(rays, tensorf, renderer, group=None, gather=True, **render_kw):
    """Each rank renders its contiguous slice of `rays` with `renderer` (OctreeRender_trilinear_fast);
    with gather=True every rank receives the full (rgb [N,3], depth [N]) -- the only communication,
    after the hot path, 16 bytes per ray."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = rays.shape[0]
    s, e = shard_bounds(n, rank, world)
    rgb, _, depth, _, _ = renderer(rays[s:e], tensorf, **render_kw)
    if world == 1 or not gather:
        return rgb, depth
    # equal-sized slots (shards differ by at most one ray): pad, gather, trim
    slot = -(-n // world)
    out = torch.zeros((slot, 4), dtype=rgb.dtype, device=rgb.device)
    out[:e - s, :3] = rgb
    out[:e - s, 3] = depth
    bufs = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(bufs, out, group=group)
    full = torch.cat([bufs[r][:b - a] for r, (a, b) in enumerate(shard_bounds(n, r, world) for r in range(world))], 0)
    return full[:, :3].contiguous(), full[:, 3].contiguous()


def allreduce_flat_(flat: torch.Tensor, group=None, average=True):
    """In-place all-reduce of one flat gradient buffer (a single collective per step)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(dist.get_world_size(group))
    return flat


class ShardedSampler:
    """SimpleSampler (train.py:25-37) for data-parallel training: every rank draws the SAME
    permutation (same seed) and takes its slice of each global batch."""

    def __init__(self, total: int, global_batch: int, rank: int = 0, world: int = 1, seed: int = 20211202):
        self.total, self.batch, self.rank, self.world = total, global_batch, rank, world
        self.curr = total
        self.ids = None
        self.rng = np.random.Generator(np.random.PCG64(seed))

    def nextids(self):
        self.curr += self.batch
        if self.curr + self.batch > self.total:
            self.ids = self.rng.permutation(self.total)
            self.curr = 0
        glob = self.ids[self.curr:self.curr + self.batch]
        s, e = shard_bounds(glob.shape[0], self.rank, self.world)
        return glob[s:e]
