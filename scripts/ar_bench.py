#!/usr/bin/env python
"""All-reduce of the flat gradient buffer (3.2 M floats = the 128^3 model, 17.4 M = 300^3) under torchrun: libtvmrender's
peer-memory kernel (NVLS and P2P paths, several grid sizes) against NCCL, back-to-back launches timed with CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import jittor_myc_nerfs_b200 as pkg
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local); dev = torch.device("cuda", local)
rank, world = pkg.dist.init_from_env("nccl", device=dev)


def timed(fn, K=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / K * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


for n in (3_200_000, 17_400_000):
    x = torch.randn(n, device=dev)
    res = {"nccl": timed(lambda: dist.all_reduce(x))}
    for no_mc in ("", "1"):
        if no_mc:
            os.environ["TVM_AR_NO_MULTICAST"] = "1"
        else:
            os.environ.pop("TVM_AR_NO_MULTICAST", None)
        for ctas in (16, 32, 64, 128):
            comm = pkg.dist.PeerComm(n, dev, n_ctas=ctas)
            res[("p2p" if not comm.multicast else "nvls") + f"/{ctas}"] = timed(lambda: comm.allreduce_())
            del comm
    if rank == 0:
        print(f"world {world}, {n * 4 / 1e6:.1f} MB: us per all-reduce:", {k: round(v, 1) for k, v in res.items()}, flush=True)
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
