"""Generates tests/golden/maint_*.npz: the callers either side of the ray path executed from the
reference's UNMODIFIED python (tensorf-myc/models/tensorBase.py, models/tensoRF.py, utils.py,
dataLoader/ray_utils.py, imported from /root/reference) over oracle/jt_shim.  Run in the build container:

    python tests/golden/make_golden_maint.py

utils.py imports plyfile / skimage (absent here) at module level for unrelated mesh/SSIM helpers; empty stand-in
modules are registered for the import only.  /root/reference is NOT needed by any test.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "jt_shim"))
sys.path.insert(0, "/root/reference/tensorf-myc")
sys.path.insert(0, HERE)

import torch
import jittor as jt
from oracle import fixtures as fx

for name in ("plyfile", "skimage", "skimage.measure"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["skimage"].measure = sys.modules["skimage.measure"]

with contextlib.redirect_stdout(io.StringIO()):
    from make_golden import build_reference_model            # same parameter injection as the ray-path goldens
    from utils import TVLoss                                  # the reference, unmodified
    # dataLoader/__init__.py pulls in the dataset classes (jittor.dataset, cv2 ...); load the one file directly
    import importlib.util
    _spec = importlib.util.spec_from_file_location("ref_ray_utils", "/root/reference/tensorf-myc/dataLoader/ray_utils.py")
    _ru = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_ru)
    get_ray_directions, get_ray_directions_blender, get_rays = _ru.get_ray_directions, _ru.get_ray_directions_blender, _ru.get_rays

quiet = lambda: contextlib.redirect_stdout(io.StringIO())


def case_alpha(name, G, mask_res, grid, shift, scale):
    case = fx.make_case(G, 8, "R1" if mask_res else "R0", mask_res=mask_res, grid_scale=scale)
    case["model"].density_shift = shift
    m = build_reference_model(case)
    with quiet(), torch.no_grad():
        alpha, dense_xyz = m.getDenseAlpha(list(grid))
        new_aabb = m.updateAlphaMask(tuple(grid))
    vol = m.alphaMask.alpha_volume.numpy().reshape(grid[::-1])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), alpha=alpha.numpy(), volume=vol, new_aabb=new_aabb.numpy(),
                        stepSize=np.float32(m.stepSize.item()),
                        args=np.array([str(G), str(mask_res), str(tuple(grid)), str(shift), str(scale)]))
    print(name, "set voxels", int(vol.sum()), "of", vol.size, "new_aabb", new_aabb.numpy().round(3).tolist())


def case_filter(name, G, mask_res, n, S):
    case = fx.make_case(G, n, "R1", mask_res=mask_res)
    rays = case["rays"].copy()
    rays[: n // 2, :3] += fx._rng(7).uniform(-6, 6, (n // 2, 3)).astype(np.float32)   # half the rays from moved origins: many miss
    m = build_reference_model(case)
    rgbs = jt.Var(np.zeros((n, 3), np.float32))
    with quiet(), torch.no_grad():
        r_b, _ = m.filtering_rays(jt.Var(rays), rgbs, bbox_only=True)
        r_m, _ = m.filtering_rays(jt.Var(rays), rgbs, N_samples=S, bbox_only=False)
    # recover the masks from the returned (compacted) rays: rows are unique
    key = lambda a: {tuple(x) for x in np.asarray(a).round(6).tolist()}
    kb, km = key(r_b.numpy()), key(r_m.numpy())
    mask_b = np.array([tuple(x) in kb for x in rays.round(6).tolist()])
    mask_m = np.array([tuple(x) in km for x in rays.round(6).tolist()])
    assert mask_b.sum() == r_b.shape[0] and mask_m.sum() == r_m.shape[0]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), rays=rays, mask_bbox=mask_b, mask_alpha=mask_m,
                        args=np.array([str(G), str(mask_res), str(n), str(S)]))
    print(name, "bbox keeps", int(mask_b.sum()), "alpha keeps", int(mask_m.sum()), "of", n)


def case_reg(name, G):
    case = fx.make_case(G, 8, "R0")
    m = build_reference_model(case)
    reg = TVLoss()
    out = {}
    for key, fn in (("tv_density", lambda: m.TV_loss_density(reg)), ("tv_app", lambda: m.TV_loss_app(reg)),
                    ("l1", m.density_L1), ("ortho", m.vector_comp_diffs)):
        for p in m.parameters():
            p.grad = None
        loss = fn()
        loss.backward()
        out[key] = np.float32(loss.item())
        for gname, plist in (("density_plane", m.density_plane), ("density_line", m.density_line),
                             ("app_plane", m.app_plane), ("app_line", m.app_line)):
            for k in range(3):
                if plist[k].grad is not None:
                    out[f"{key}.{gname}.{k}"] = plist[k].grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), args=np.array([str(G)]), **out)
    print(name, {k: float(v) for k, v in out.items() if v.ndim == 0})


def case_resize(name, G, target, new_aabb):
    case = fx.make_case(G, 8, "R1", mask_res=16)
    m = build_reference_model(case)
    out = {}
    with quiet(), torch.no_grad():
        m.upsample_volume_grid(list(target))
        for gname, plist in (("density_plane", m.density_plane), ("density_line", m.density_line),
                             ("app_plane", m.app_plane), ("app_line", m.app_line)):
            for k in range(3):
                out[f"up.{gname}.{k}"] = plist[k].numpy().copy()
        out["up.stepSize"], out["up.nSamples"] = np.float32(m.stepSize.item()), np.int64(m.nSamples)
        m.shrink(jt.Var(np.asarray(new_aabb, np.float32)))
        for gname, plist in (("density_plane", m.density_plane), ("density_line", m.density_line),
                             ("app_plane", m.app_plane), ("app_line", m.app_line)):
            for k in range(3):
                out[f"shrink.{gname}.{k}"] = plist[k].numpy().copy()
        out["shrink.aabb"] = m.aabb.numpy().copy()
        out["shrink.gridSize"] = np.asarray(m.gridSize.numpy(), np.int64)
        out["shrink.stepSize"], out["shrink.nSamples"] = np.float32(m.stepSize.item()), np.int64(m.nSamples)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), new_aabb=np.asarray(new_aabb, np.float32),
                        args=np.array([str(G), str(tuple(target))]), **out)
    print(name, "->", out["up.density_plane.0"].shape, "shrunk to", out["shrink.gridSize"].tolist(), out["shrink.aabb"].round(3).tolist())


def case_rays(name, H, W):
    focal = 0.5 * W / np.tan(0.5 * 0.6911)
    c2w = fx.camera_pose(0.7, 0.5).astype(np.float32)
    d = get_ray_directions(H, W, [focal, focal])
    db = get_ray_directions_blender(H, W, [focal, focal * 1.1], center=[W / 2 - 0.25, H / 2 + 1.5])
    dn = d / jt.norm(d, dim=-1, keepdim=True)                 # blender.py:75
    o, r = get_rays(dn, jt.Var(c2w))
    ob, rb = get_rays(db, jt.Var(c2w))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), c2w=c2w, focal=np.float32(focal), rays_o=o.numpy(), rays_d=r.numpy(),
                        rays_o_blender=ob.numpy(), rays_d_blender=rb.numpy(), args=np.array([str(H), str(W)]))
    print(name, r.numpy()[:2])


if __name__ == "__main__":
    case_alpha("maint_alpha_masked", 40, 32, (36, 33, 30), -9.5, 0.6)
    case_alpha("maint_alpha_nomask", (32, 40, 48), None, (40, 40, 40), -9.0, 0.5)
    case_filter("maint_filter", 40, 32, 768, 96)
    case_reg("maint_reg", (20, 24, 28))
    case_resize("maint_resize", (20, 24, 28), (34, 31, 29), [[-3.1, -2.4, -4.2], [2.2, 4.1, 3.3]])
    case_rays("maint_rays", 37, 53)
