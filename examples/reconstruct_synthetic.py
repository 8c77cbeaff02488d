#!/usr/bin/env python
"""The reference's reconstruction loop (tensorf-myc/train.py:120-360) on a synthetic scene, driven through this package:
coarse-to-fine schedule (updateAlphaMask + shrink, filtering_rays, upsample_volume_grid with a fresh optimiser), TV
regularisers, lr decay, evaluation through OctreeRender_trilinear_fast, checkpoint save / reload.

Scene: a matte sphere (radius 2, colour 0.5 + 0.5 n) on a white background, seen by pinhole cameras on a circle of radius
12 -- analytic ground truth, no dataset needed.  Usage:  python examples/reconstruct_synthetic.py [--iters 600]"""
from __future__ import annotations

import argparse
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jittor_myc_nerfs_b200 as pkg                      # noqa: E402
from jittor_myc_nerfs_b200.utils import N_to_reso, cal_n_samples, SimpleSampler   # noqa: E402


def camera(az, el, radius=12.0):
    """camera-to-world (columns right, up, back) of a camera on a sphere looking at the origin; with get_ray_directions'
    (-x, y, -1) camera-space directions (dataLoader/ray_utils.py:101) the rays leave along -back = towards the origin."""
    eye = radius * np.array([math.cos(el) * math.cos(az), math.cos(el) * math.sin(az), math.sin(el)])
    fwd = -eye / np.linalg.norm(eye)
    right = np.cross(fwd, [0.0, 0.0, 1.0])
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, up, -fwd, eye
    return c2w


def ground_truth(rays, R=2.0):
    """rgb of the analytic scene for rays [n,6] (unit directions)."""
    o, d = rays[:, :3], rays[:, 3:6]
    b = (o * d).sum(-1)
    disc = b * b - ((o * o).sum(-1) - R * R)
    hit = disc > 0
    t = -b - torch.sqrt(disc.clamp_min(0))
    n = (o + d * t[:, None]) / R
    rgb = torch.ones_like(o)
    rgb[hit] = 0.5 + 0.5 * n[hit]
    return rgb


def psnr_of(model, rays, rgb, nSamples):
    with torch.no_grad():
        out, _, _, _, _ = pkg.OctreeRender_trilinear_fast(rays, model, chunk=4096, N_samples=nSamples, white_bg=True, is_train=False)
    return float(-10.0 * torch.log10(torch.mean((out - rgb) ** 2)))


def run(iters=600, res=96, n_views=16, batch=4096, N_voxel_init=32 ** 3, N_voxel_final=64 ** 3, upsamp_list=(250, 400),
        update_AlphaMask_list=(150, 300), lr_init=0.02, lr_basis=1e-3, lr_decay_target_ratio=0.1, TV_weight_density=0.1,
        TV_weight_app=0.01, step_ratio=0.5, mlp_mode="fp32", seed=0, ckpt_path=None, log=print, model_name="TensorVMSplit",
        normal_vector_penalty_weight=0.5, radii=14.0):
    """model_name: TensorVMSplit | REFTensoRF (adds normal_vector_penalty_weight * tensorf.penalty, train.py:253-257) |
    NerfPlusPlus (set_nerfplusplus(2, 2, 3, radii): foreground on black inside the radii-sphere + background MLP)."""
    dev = torch.device("cuda:0")
    torch.manual_seed(seed)
    focal = 0.5 * res / math.tan(0.5 * 0.6911)
    views = [pkg.get_rays_frame(camera(2 * math.pi * i / n_views, 0.45), res, res, focal, blender=False, device=dev) for i in range(n_views)]
    allrays = torch.cat(views)
    allrgbs = ground_truth(allrays)
    test_rays = pkg.get_rays_frame(camera(0.37, 0.6), res, res, focal, blender=False, device=dev)
    test_rgb = ground_truth(test_rays)

    aabb = torch.tensor([[-3.0, -3.0, -3.0], [3.0, 3.0, 3.0]])
    reso_cur = N_to_reso(N_voxel_init, aabb)
    nSamples = min(int(1e6), cal_n_samples(reso_cur, step_ratio))
    tensorf = getattr(pkg, model_name)(aabb, reso_cur, dev, density_n_comp=[16, 16, 16], appearance_n_comp=[48, 48, 48], app_dim=27,
                                near_far=[8.0, 16.0], shadingMode="MLP_Fea", alphaMask_thres=1e-4, density_shift=-10,
                                distance_scale=25, pos_pe=6, view_pe=2, fea_pe=2, featureC=128, step_ratio=step_ratio,
                                fea2denseAct="softplus")
    tensorf.set_nerfplusplus(bg_freq=2, bg_view_freq=2, bg_D=3, radii=radii)        # train.py:173 (a no-op except for NerfPlusPlus)
    tensorf.mlp_mode = mlp_mode
    npp, ref = model_name == "NerfPlusPlus", model_name == "REFTensoRF"
    grad_vars = tensorf.get_optparam_groups(lr_init, lr_basis)
    lr_factor = lr_decay_target_ratio ** (1 / iters)
    optimizer = pkg.Adam(grad_vars, betas=(0.9, 0.99))
    upsamp_list, update_AlphaMask_list = list(upsamp_list), list(update_AlphaMask_list)
    N_voxel_list = torch.round(torch.exp(torch.linspace(math.log(N_voxel_init), math.log(N_voxel_final), len(upsamp_list) + 1))).long().tolist()[1:]
    allrays, allrgbs = tensorf.filtering_rays(allrays, allrgbs, bbox_only=True)
    sampler = SimpleSampler(allrays.shape[0], batch, seed=seed)
    tvreg = pkg.TVLoss()
    hist = dict(psnr_train=[], n_rays=[allrays.shape[0]], aabb=[tensorf.aabb.clone()], reso=[list(reso_cur)], psnr_test=[])
    reso_mask = reso_cur
    for iteration in range(iters):
        optimizer.zero_grad()
        idx = sampler.nextids().to(dev)
        rays_train, rgb_train = allrays[idx], allrgbs[idx]
        rgb_map, _, _, _, _ = pkg.OctreeRender_trilinear_fast(rays_train, tensorf, chunk=batch, N_samples=nSamples, white_bg=True,
                                                            is_train=True)
        loss = torch.mean((rgb_map - rgb_train) ** 2)
        total = loss
        if TV_weight_density > 0:
            TV_weight_density *= lr_factor
            total = total + tensorf.TV_loss_density(tvreg) * TV_weight_density
        if TV_weight_app > 0:
            TV_weight_app *= lr_factor
            total = total + tensorf.TV_loss_app(tvreg) * TV_weight_app
        if ref and normal_vector_penalty_weight > 0:
            total = total + normal_vector_penalty_weight * tensorf.penalty.sum()
        total.backward()
        optimizer.step()
        hist["psnr_train"].append(float(-10.0 * math.log10(max(float(loss), 1e-12))))
        for g in optimizer.param_groups:
            g["lr"] = g["lr"] * lr_factor
        if iteration in update_AlphaMask_list:
            if reso_cur[0] * reso_cur[1] * reso_cur[2] < 256 ** 3:
                reso_mask = reso_cur
            new_aabb = tensorf.updateAlphaMask(tuple(reso_mask))
            if iteration == update_AlphaMask_list[0]:
                tensorf.shrink(new_aabb)
                hist["aabb"].append(tensorf.aabb.clone())
            if iteration == update_AlphaMask_list[1]:
                allrays, allrgbs = tensorf.filtering_rays(allrays, allrgbs)
                sampler = SimpleSampler(allrgbs.shape[0], batch, seed=seed + 1)
                hist["n_rays"].append(allrays.shape[0])
        if iteration in upsamp_list:
            n_voxels = N_voxel_list.pop(0)
            reso_cur = N_to_reso(n_voxels, tensorf.aabb)
            nSamples = min(int(1e6), cal_n_samples(reso_cur, step_ratio))
            tensorf.upsample_volume_grid(reso_cur)
            hist["reso"].append(list(reso_cur))
            lr_scale = lr_decay_target_ratio ** (iteration / iters)
            optimizer = pkg.Adam(tensorf.get_optparam_groups(lr_init * lr_scale, lr_basis * lr_scale), betas=(0.9, 0.99))
        if iteration % 100 == 99 or iteration == iters - 1:
            p = psnr_of(tensorf, test_rays, test_rgb, nSamples)
            hist["psnr_test"].append(p)
            log(f"iter {iteration + 1:5d}: train psnr {np.mean(hist['psnr_train'][-100:]):.2f} dB, held-out view {p:.2f} dB, "
                f"grid {reso_cur}, {allrays.shape[0]} rays, aabb {tensorf.aabb.flatten().tolist()}")
    hist["final_psnr"] = hist["psnr_test"][-1]
    if ckpt_path:
        tensorf.save(ckpt_path, {"global_step": iters - 1, "lr": [g["lr"] for g in optimizer.param_groups]})
        ckpt = pkg.load_checkpoint(ckpt_path)
        kwargs = ckpt["kwargs"]
        kwargs.update({"device": dev})
        bg = [kwargs.pop(k) for k in ("bg_freq", "bg_view_freq", "bg_D", "radii")] if "bg_freq" in kwargs else None
        again = getattr(pkg, model_name)(**kwargs)
        if bg is not None:
            again.set_nerfplusplus(*bg)
        again.load(ckpt)
        again.mlp_mode = mlp_mode
        hist["reload_psnr"] = psnr_of(again, test_rays, test_rgb, nSamples)
    hist["model"] = tensorf
    return hist


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=600)
    ap.add_argument("--mlp", default="fp32", choices=["fp32", "bf16", "fp16"])
    ap.add_argument("--ckpt", default=None)
    ap.add_argument("--model", default="TensorVMSplit", choices=["TensorVMSplit", "REFTensoRF", "NerfPlusPlus"])
    a = ap.parse_args()
    h = run(iters=a.iters, mlp_mode=a.mlp, ckpt_path=a.ckpt, model_name=a.model)
    print("final held-out PSNR %.2f dB" % h["final_psnr"])
