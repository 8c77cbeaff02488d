"""CPU restatement of the callers either side of the ray path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py may import this.
Every function cites the reference lines (relative to tensorf-myc/) it follows; it is pinned by
tests/golden/maint_*.npz, recorded by running the reference's UNMODIFIED python over oracle/jt_shim
(tests/golden/make_golden_maint.py).  Jittor op numerics assumed beyond A1-A6 of tensorf_oracle.py:
  A8  jt.linspace(a, b, n)            = torch.linspace semantics (symmetric evaluation from both ends)
  A9  nn.interpolate(bilinear, align_corners=True): src = dst * (in-1)/(out-1); 4-tap blend as grid_sample
  A10 nn.max_pool3d(k=3, padding=1, stride=1): -inf padding
  A11 jt.optim.Adam.step (jittor/optim.py, as recalled; Jittor itself is absent from the image):
        m = b0 m + (1-b0) g;  v = b1 v + (1-b1) g^2;  p -= m * (lr sqrt(1-b1^n) / (1-b0^n)) / (sqrt(v) + eps)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import tensorf_oracle as orc

MAT_MODE = orc.MAT_MODE if hasattr(orc, "MAT_MODE") else ((0, 1), (0, 2), (1, 2))
VEC_MODE = orc.VEC_MODE if hasattr(orc, "VEC_MODE") else (2, 1, 0)


# ---- §8f-1: getDenseAlpha / updateAlphaMask (models/tensorBase.py:366-409) -----------------------------------
def dense_lattice(aabb, gridSize, dtype=torch.float32):
    """tensorBase.py:371-376: dense_xyz[i,j,k] = aabb0 * (1 - s) + aabb1 * s, s = (lin_x[i], lin_y[j], lin_z[k])."""
    aabb = torch.as_tensor(np.asarray(aabb), dtype=dtype)
    g = [int(x) for x in gridSize]
    samples = torch.stack(torch.meshgrid(torch.linspace(0, 1, g[0], dtype=dtype), torch.linspace(0, 1, g[1], dtype=dtype),
                                         torch.linspace(0, 1, g[2], dtype=dtype), indexing="ij"), -1)
    return aabb[0] * (1 - samples) + aabb[1] * samples


def get_dense_alpha(model, gridSize=None):
    """tensorBase.py:366-384 -> (alpha [Gx,Gy,Gz], dense_xyz [Gx,Gy,Gz,3])."""
    gridSize = [int(g) for g in (model.gridSize if gridSize is None else gridSize)]
    dense_xyz = dense_lattice(model.aabb, gridSize, model.dtype)
    alpha = torch.zeros_like(dense_xyz[..., 0])
    for i in range(gridSize[0]):
        alpha[i] = model.compute_alpha(dense_xyz[i].reshape(-1, 3), float(model.stepSize)).view(gridSize[1], gridSize[2])
    return alpha, dense_xyz


def update_alpha_mask(model, gridSize=(200, 200, 200), thres=0.001):
    """tensorBase.py:386-409 -> dict(volume [Gz,Gy,Gx] {0,1}, pooled (pre-threshold), new_aabb [2,3], alpha_dense [Gz,Gy,Gx])."""
    gridSize = [int(g) for g in gridSize]
    alpha, dense_xyz = get_dense_alpha(model, gridSize)
    dense_xyz = dense_xyz.transpose(0, 2)
    alpha_t = alpha.clamp(0, 1).transpose(0, 2).contiguous()
    pooled = F.max_pool3d(alpha_t[None, None], kernel_size=3, padding=1, stride=1).view(gridSize[::-1])
    vol = torch.where(pooled >= thres, torch.ones_like(pooled), torch.zeros_like(pooled))
    valid = dense_xyz[vol > 0.5]
    new_aabb = torch.stack((valid.amin(0), valid.amax(0)))
    return dict(volume=vol.numpy(), pooled=pooled.numpy(), new_aabb=new_aabb.numpy(), alpha_dense=alpha_t.numpy())


# ---- §8f-3: filtering_rays (models/tensorBase.py:411-441), ray generation (dataLoader/ray_utils.py:81-153) -----
def filtering_rays_mask(model, all_rays, N_samples=256, bbox_only=False):
    """tensorBase.py:411-441 -> bool mask [N] (the reference then returns all_rays[mask], all_rgbs[mask])."""
    rays = torch.as_tensor(np.asarray(all_rays), dtype=model.dtype).reshape(-1, all_rays.shape[-1])
    rays_o, rays_d = rays[..., :3], rays[..., 3:6]
    if bbox_only:
        vec = torch.where(rays_d == 0, torch.full_like(rays_d, 1e-6), rays_d)
        rate_a = (model.aabb[1] - rays_o) / vec
        rate_b = (model.aabb[0] - rays_o) / vec
        t_min = torch.minimum(rate_a, rate_b).amax(-1)
        t_max = torch.maximum(rate_a, rate_b).amin(-1)
        return (t_max > t_min).numpy()
    xyz_sampled, _, _ = model.sample_ray(rays_o, rays_d, N_samples=N_samples, is_train=False)
    a = model.alphaMask.sample_alpha(xyz_sampled.reshape(-1, 3)).view(xyz_sampled.shape[:-1])
    return (a > 0).any(-1).numpy()


def get_ray_directions(H, W, focal, center=None, blender=False):
    """ray_utils.py:81-131: pixel centres (+0.5); (-(i-cx)/fx, (j-cy)/fy, -1), or the *_blender sign convention."""
    i = (torch.arange(W, dtype=torch.float32) + 0.5)[None, :].expand(H, W)
    j = (torch.arange(H, dtype=torch.float32) + 0.5)[:, None].expand(H, W)
    cent = center if center is not None else [W / 2, H / 2]
    if blender:
        return torch.stack([-(i - cent[0]) / focal[0], -(j - cent[1]) / focal[1], torch.ones_like(i)], -1)
    return torch.stack([-(i - cent[0]) / focal[0], (j - cent[1]) / focal[1], -torch.ones_like(i)], -1)


def get_rays(directions, c2w):
    """ray_utils.py:134-153: rays_d = directions @ c2w[:3,:3].T ; rays_o = c2w[:3,3] broadcast."""
    c2w = torch.as_tensor(np.asarray(c2w), dtype=torch.float32)
    rays_d = directions @ c2w[:3, :3].T
    rays_o = c2w[:3, 3].expand(rays_d.shape)
    return rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)


def frame_rays(H, W, focal, c2w, normalize=True, blender=False):
    """blender.py:66-75 style frame: directions normalised BEFORE the rotation -> all_rays [H*W, 6]."""
    d = get_ray_directions(H, W, focal, blender=blender)
    if normalize:
        d = d / torch.norm(d, dim=-1, keepdim=True)
    o, dd = get_rays(d, c2w)
    return torch.cat([o, dd], 1).numpy()


# ---- §8f-2: regularisers (models/tensoRF.py:177-207, utils.py:123-142) and the optimiser step ---------------------
def tv_loss(x, weight=1.0):
    """utils.TVLoss.execute (utils.py:128-139) for x [B,C,H,W]."""
    B, C, H, W = x.shape
    count_h = C * (H - 1) * W
    count_w = C * H * (W - 1)
    h_tv = torch.pow(x[:, :, 1:, :] - x[:, :, :H - 1, :], 2).sum()
    w = h_tv / count_h
    if count_w > 0:
        w = w + torch.pow(x[:, :, :, 1:] - x[:, :, :, :W - 1], 2).sum() / count_w
    return weight * 2 * w / B


def tv_loss_planes(planes):
    """TV_loss_density / TV_loss_app (tensoRF.py:197-207): sum_k reg(plane_k) * 1e-2."""
    total = 0
    for p in planes:
        total = total + tv_loss(p) * 1e-2
    return total


def density_l1(planes, lines):
    """tensoRF.py:191-195."""
    total = 0
    for p, l in zip(planes, lines):
        total = total + torch.mean(torch.abs(p)) + torch.mean(torch.abs(l))
    return total


def vector_diffs(lines):
    """tensoRF.py:177-186: mean |off-diagonal of V V^T| per line tensor [1,C,L,1]."""
    total = 0
    for v in lines:
        n_comp, n_size = v.shape[1:-1]
        m = v.view(n_comp, n_size)
        dotp = m @ m.transpose(-1, -2)
        non_diagonal = dotp.view(-1)[1:].view(n_comp - 1, n_comp + 1)[..., :-1]
        total = total + torch.mean(torch.abs(non_diagonal))
    return total


def adam_step(p, g, m, v, lr, n, betas=(0.9, 0.99), eps=1e-8):
    """A11 (jt.optim.Adam as constructed at train.py:187): in-place update of numpy/torch arrays; n = step count (1-based)."""
    b0, b1 = betas
    m.mul_(b0).add_(g, alpha=1 - b0)
    v.mul_(b1).add_(g * g, alpha=1 - b1)
    step_size = lr * float(np.sqrt(1 - b1 ** n)) / (1 - b0 ** n)
    p.sub_(m * step_size / (torch.sqrt(v) + eps))


# ---- §8f-4: upsample_volume_grid / shrink (models/tensoRF.py:248-314) ------------------------------------------------
def up_sampling_vm(planes, lines, res_target):
    """tensoRF.py:248-262."""
    out_p, out_l = [], []
    for i in range(3):
        m0, m1 = MAT_MODE[i]
        out_p.append(F.interpolate(planes[i], size=(int(res_target[m1]), int(res_target[m0])), mode="bilinear", align_corners=True))
        out_l.append(F.interpolate(lines[i], size=(int(res_target[VEC_MODE[i]]), 1), mode="bilinear", align_corners=True))
    return out_p, out_l


def shrink_indices(aabb, units, gridSize, new_aabb):
    """tensoRF.py:275-280: voxel index box [t_l, b_r) of the crop."""
    aabb = torch.as_tensor(np.asarray(aabb), dtype=torch.float32)
    units = torch.as_tensor(np.asarray(units), dtype=torch.float32)
    new_aabb = torch.as_tensor(np.asarray(new_aabb), dtype=torch.float32)
    t_l, b_r = (new_aabb[0] - aabb[0]) / units, (new_aabb[1] - aabb[0]) / units
    t_l, b_r = torch.round(torch.round(t_l)).long(), torch.round(b_r).long() + 1
    b_r = torch.stack([b_r, torch.as_tensor(np.asarray(gridSize)).long()]).amin(0)
    return t_l, b_r


def shrink(aabb, units, gridSize, new_aabb, planes_d, lines_d, planes_a, lines_a, mask_grid_equal):
    """tensoRF.py:271-314 -> (cropped grids, new_aabb', newSize)."""
    aabb_t = torch.as_tensor(np.asarray(aabb), dtype=torch.float32)
    t_l, b_r = shrink_indices(aabb, units, gridSize, new_aabb)
    out = {"dp": [], "dl": [], "ap": [], "al": []}
    for i in range(3):
        v = VEC_MODE[i]
        out["dl"].append(lines_d[i][..., int(t_l[v]):int(b_r[v]), :])
        out["al"].append(lines_a[i][..., int(t_l[v]):int(b_r[v]), :])
        m0, m1 = MAT_MODE[i]
        out["dp"].append(planes_d[i][..., int(t_l[m1]):int(b_r[m1]), int(t_l[m0]):int(b_r[m0])])
        out["ap"].append(planes_a[i][..., int(t_l[m1]):int(b_r[m1]), int(t_l[m0]):int(b_r[m0])])
    new_aabb = torch.as_tensor(np.asarray(new_aabb), dtype=torch.float32)
    if not mask_grid_equal:
        g = torch.as_tensor(np.asarray(gridSize), dtype=torch.float32)
        t_l_r, b_r_r = t_l / (g - 1), (b_r - 1) / (g - 1)
        correct = torch.zeros_like(new_aabb)
        correct[0] = (1 - t_l_r) * aabb_t[0] + t_l_r * aabb_t[1]
        correct[1] = (1 - b_r_r) * aabb_t[0] + b_r_r * aabb_t[1]
        new_aabb = correct
    return out, new_aabb.numpy(), (b_r - t_l).numpy()
