"""OctreeRender_trilinear_fast with the reference's signature (tensorf-myc/renderer.py:12-27).

The reference slices `rays` into `chunk`-sized pieces and synchronises the device after each one
(jt.sync_all + jt.gc).  Rays are independent, so the result does not depend on the chunk size; this
implementation keeps the argument for drop-in compatibility but launches as many rays per kernel as
the workspace budget allows and never synchronises.
"""
from __future__ import annotations

import torch


def OctreeRender_trilinear_fast(rays, tensorf, chunk=4096, N_samples=-1, ndc_ray=False, white_bg=True,
                                is_train=False, device='cuda'):
    if not torch.is_tensor(rays):
        rays = torch.as_tensor(rays, dtype=torch.float32)
    if not rays.is_cuda:
        rays = rays.to(device, non_blocking=True)
    rays = rays.reshape(-1, rays.shape[-1])[:, :6].contiguous()
    rgb_map, depth_map = tensorf(rays, is_train=is_train, white_bg=white_bg, ndc_ray=ndc_ray, N_samples=N_samples)
    return rgb_map, None, depth_map, None, None
