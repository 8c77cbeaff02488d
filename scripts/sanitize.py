#!/usr/bin/env python
"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck): forward in every MLP mode, backward in
fp32 and on the tensor cores, the maintenance kernels, on a 48^3 model with ragged ray / sample counts."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import jittor_myc_nerfs_b200 as pkg
import synthetic as fx
dev = torch.device("cuda:0")
n, S = int(os.environ.get("SAN_RAYS", "333")), 139
case = fx.make_case(48, n, "R2", mask_res=48, train=True)
rays = torch.from_numpy(case["rays"]).to(dev)
jit = torch.from_numpy(case["jitter"]).to(dev)
tgt = torch.from_numpy(case["target"]).to(dev)
for mode in os.environ.get("SAN_MODES", "fp32,bf16,fp16").split(","):
    m = pkg.model_from_params(case["model"], "cuda:0", case["alpha_volume"], case["alpha_aabb"], mode)
    m.app_planes_bf16 = mode != "fp32"
    with torch.no_grad():
        rgb, _ = m(rays, white_bg=True, is_train=False, N_samples=S)
    rgb, _ = m(rays, is_train=True, white_bg=True, N_samples=S, jitter=jit)
    torch.mean((rgb - tgt) ** 2).backward()
    torch.cuda.synchronize()
    print(mode, "ok", float(rgb.detach().mean()), flush=True)
m = pkg.model_from_params(case["model"], "cuda:0", case["alpha_volume"], case["alpha_aabb"], "fp32")
m.updateAlphaMask((48, 48, 48))
keep = m.filtering_mask(rays, N_samples=S)
tv = pkg.TVLoss()
(m.TV_loss_density(tv) + m.TV_loss_app(tv)).backward()
opt = pkg.Adam(m.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
opt.step()
torch.cuda.synchronize()
print("maintenance ok", int(keep.sum()), flush=True)
