#!/usr/bin/env python
"""Where does the data-parallel training step lose time?  Wall clock per step() vs CUDA-event time per step(), with and without
the gradient exchange, under torchrun at any world size (diagnostic; prints one line per rank 0)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.distributed as dist
import jittor_myc_nerfs_b200 as pkg
import synthetic as fx
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local); dev = torch.device("cuda", local)
rank, world = pkg.dist.init_from_env("nccl", device=dev)
n, G = 4096, 128
mp = fx.make_model(G, density_shift=0.0)
pool = np.concatenate([fx.subset_rays(n, azimuth=0.7 + v * np.pi / 4) for v in range(8)])
perm = np.random.default_rng(fx.SEED_BASE + 7).permutation(pool.shape[0])
rays = torch.from_numpy(np.ascontiguousarray(pool[perm[(rank % 8) * n:(rank % 8 + 1) * n]])).to(dev)
tgt = torch.from_numpy(fx.target_rgb(n)).to(dev)
S = 443
for mode in sys.argv[1:] or ["none", "peer"]:
    model = pkg.model_from_params(mp, f"cuda:{local}", fx.ball_alpha_volume(128), mp.aabb.copy(), "bf16")
    if world > 1 and mode == "peer":
        model.enable_peer_allreduce()
    elif world > 1 and mode == "nccl":
        model.grad_sync = True
    opt = pkg.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
    g = pkg.TrainStepGraph(model, opt, n, S, white_bg=True, TV_weight_density=2.0, TV_weight_app=2.0)
    for _ in range(10):
        g.step(rays, tgt)
    torch.cuda.synchronize(); 
    if world > 1: dist.barrier()
    K = 300
    evs = []
    t0 = time.perf_counter()
    for _ in range(K):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.step(rays, tgt); b.record(); evs.append((a, b))
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t0
    ev = sum(a.elapsed_time(b) for a, b in evs) / K
    # replay only (no per-step copies / uploads)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1: dist.barrier()
    a.record()
    for _ in range(K):
        g.graph.replay()
    b.record(); torch.cuda.synchronize()
    rep = a.elapsed_time(b) / K
    t = torch.tensor([ev, rep, t_host / K * 1e3, t_wall / K * 1e3], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world {world} exchange {mode}: event ms/step {t[0]:.4f} | back-to-back replay ms {t[1]:.4f} | host enqueue ms/step {t[2]:.4f} | wall ms/step {t[3]:.4f}", flush=True)
    del g
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
sys.stdout.flush()
os._exit(0)
