#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU): (a) ray-sharded render == single-GPU render,
(b) data-parallel training step: flat-buffer NCCL all-reduce of gradients == oracle full-batch gradient."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import jittor_myc_nerfs_b200 as pkg
import synthetic as fx
from oracle import tensorf_oracle as orc

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
rank, world = pkg.dist.init_from_env("nccl", device=dev)
case = fx.make_case(48, 1024, "R2", mask_res=48, train=True)
model = pkg.model_from_params(case["model"], f"cuda:{local}", case["alpha_volume"], case["alpha_aabb"])
rays = torch.from_numpy(case["rays"]).to(dev)
with torch.no_grad():
    full_rgb, _, full_depth, _, _ = pkg.OctreeRender_trilinear_fast(rays, model, N_samples=167, is_train=False)
    rgb, depth = pkg.dist.render_sharded(rays, model, pkg.OctreeRender_trilinear_fast, gather=True, N_samples=167,
                                         is_train=False, white_bg=True)
assert torch.equal(rgb, full_rgb) and torch.equal(depth, full_depth), "sharded render differs"
# DP step: each rank differentiates the MSE of its slice; grads are averaged by ONE flat all-reduce
model.grad_sync = world > 1
s, e = pkg.dist.shard_bounds(1024, rank, world)
# equal shard sizes => mean of per-rank means == global mean
jit = torch.from_numpy(case["jitter"]).to(dev)
tgt = torch.from_numpy(case["target"]).to(dev)
out, _ = model(rays[s:e], is_train=True, N_samples=167, jitter=jit[s:e])
loss = torch.mean((out - tgt[s:e]) ** 2)
loss.backward()
torch.cuda.synchronize()
ref = orc.backward_case(case, N_samples=167)
worst = 0.0
for name, p in (("density_plane.0", model.density_plane[0]), ("app_line.2", model.app_line[2]),
                ("basis_mat.weight", model.basis_mat.weight), ("renderModule.mlp.0.weight", model.renderModule.mlp[0].weight)):
    g, r = p.grad.cpu().numpy().astype(np.float64), ref["grads"][name]
    worst = max(worst, np.abs(g - r).max() / np.abs(r).max())
assert worst <= 1e-4, worst
if rank == 0:
    print(f"dp_check OK: world={world}, sharded render bit-identical, DP grad rel err {worst:.2e}")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
