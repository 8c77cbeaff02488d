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


class PeerComm:
    """Symmetric-memory communicator of tvm_allreduce_sum (include/tvmrender.h): a flat fp32 buffer every rank of the node
    allocates identically (torch.distributed._symmetric_memory: cuMem allocations mapped into every peer, plus the NVLS
    multicast mapping when the fabric has one), a signal pad, and this rank's epoch counter.  Construction is a collective.
    The all-reduce itself is libtvmrender's own kernel over those peer pointers -- no NCCL call on the training path."""

    def __init__(self, n_floats: int, device, group=None, n_ctas: int = 64):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib as L
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > L.AR_MAX_WORLD:
            raise L.TvmError(f"tvm_allreduce_sum handles up to {L.AR_MAX_WORLD} peers of one node")
        n_floats = (int(n_floats) + 3) // 4 * 4
        lib = L.load()
        words = C.c_size_t(0)
        L.check(lib.tvm_allreduce_signal_words(self.world, C.byref(words)), "tvm_allreduce_signal_words")
        self.buf = symm_mem.empty(n_floats, dtype=torch.float32, device=device)
        self.sig = symm_mem.empty(int(words.value), dtype=torch.int32, device=device)
        self.buf.zero_()
        self.sig.zero_()
        self._hdl = symm_mem.rendezvous(self.buf, group)
        self._sig_hdl = symm_mem.rendezvous(self.sig, group)
        self.epoch = torch.zeros(2, dtype=torch.int32, device=device)
        c = L.TvmPeerComm()
        for p in range(self.world):
            c.bufs[p] = int(self._hdl.buffer_ptrs[p])
            c.signals[p] = int(self._sig_hdl.buffer_ptrs[p])
        mc = int(getattr(self._hdl, "multicast_ptr", 0) or 0)
        if os.environ.get("TVM_AR_NO_MULTICAST"):
            mc = 0
        c.multicast = mc or None
        c.epoch_dev = self.epoch.data_ptr()
        c.rank, c.world = self.rank, self.world
        self.struct, self.multicast, self.n_ctas = c, bool(mc), int(n_ctas)
        torch.cuda.synchronize(device)
        dist.barrier(group)          # every pad is zero before anybody signals

    def allreduce_(self, offset_floats=0, n_floats=None):
        """In-place sum over the ranks of buf[offset : offset + n] on the current stream."""
        import ctypes as C
        from . import _lib as L
        n = self.buf.numel() - offset_floats if n_floats is None else n_floats
        L.check(L.load().tvm_allreduce_sum(C.byref(self.struct), int(offset_floats), int(n), self.n_ctas,
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tvm_allreduce_sum")
        return self.buf


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
