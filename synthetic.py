"""Seeded synthetic inputs for the TensoRF-VM hot path.

The workload generator of bench.py, __graft_entry__.smoke() and the tests (random-init models of the reference's shapes,
pinhole rays, analytic alpha volumes, jitter and target colours): inputs only, no rendering arithmetic.  It depends on numpy
only (no torch, no CUDA) so that the inputs are bit-identical on the build container and on the GPU box.

Shapes/hyper-parameters follow the reference's Scar-style configuration:
  tensorf-myc/configs/Scar.txt:7-10,26-35   bbox +-5, near/far 5/40, n_lamb 16/48, MLP_Fea, pe 2/2
  tensorf-myc/opt.py                         data_dim_color 27, featureC 128, step_ratio 0.5,
                                             distance_scale 25, density_shift -10
  tensorf-myc/models/tensoRF.py:154-164      grid shapes, 0.1*randn init
  tensorf-myc/models/tensorBase.py:62-74     MLP shapes, last bias 0
  tensorf-myc/dataLoader/ray_utils.py:81-104 ray directions; blender.py:66-75 focal + normalise
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SEED_BASE = 20211202  # the reference's own seed, tensorf-myc/train.py:396-397

MAT_MODE = ((0, 1), (0, 2), (1, 2))  # tensorBase.py:168
VEC_MODE = (2, 1, 0)                 # tensorBase.py:169


def _rng(offset: int, seed: int = SEED_BASE) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed + offset))


@dataclass
class ModelParams:
    """Host (numpy, fp32) parameters in the reference's own NCHW shapes and names."""
    aabb: np.ndarray                      # [2,3]
    gridSize: tuple                       # (Gx,Gy,Gz)
    density_plane: list                   # 3 x [1,Cd,G[m1],G[m0]]
    density_line: list                    # 3 x [1,Cd,G[v],1]
    app_plane: list                       # 3 x [1,Ca,G[m1],G[m0]]
    app_line: list                        # 3 x [1,Ca,G[v],1]
    basis_mat: np.ndarray                 # [app_dim, 3*Ca]   (Linear weight, no bias)
    mlp_w: list                           # [[F,in_mlpC],[F,F],[3,F]]
    mlp_b: list                           # [[F],[F],[3]]
    near_far: tuple = (5.0, 40.0)
    density_shift: float = -10.0
    distance_scale: float = 25.0
    rayMarch_weight_thres: float = 1e-4
    step_ratio: float = 0.5
    view_pe: int = 2
    fea_pe: int = 2
    app_dim: int = 27
    featureC: int = 128
    density_n_comp: tuple = (16, 16, 16)
    app_n_comp: tuple = (48, 48, 48)
    fea2denseAct: str = "softplus"
    extra: dict = field(default_factory=dict)


def make_model(G, seed: int = SEED_BASE, density_shift: float = -10.0,
               cd: int = 16, ca: int = 48, app_dim: int = 27, featureC: int = 128,
               view_pe: int = 2, fea_pe: int = 2, bbox: float = 5.0,
               near_far=(5.0, 40.0), grid_scale: float = 0.1, variant: str = "vm") -> ModelParams:
    """Random-init TensorVMSplit parameters (seed+0).  variant="ref" adds REFTensoRF's four heads
    (models/REFTensoRF.py:80-95) and widens the MLP input by the dot-product column (:9)."""
    if isinstance(G, int):
        G = (G, G, G)
    G = tuple(int(g) for g in G)
    r = _rng(0, seed)
    f32 = np.float32
    aabb = np.array([[-bbox] * 3, [bbox] * 3], dtype=f32)

    def grids(c):
        planes, lines = [], []
        for k in range(3):
            m0, m1 = MAT_MODE[k]
            v = VEC_MODE[k]
            planes.append((grid_scale * r.standard_normal((1, c, G[m1], G[m0]))).astype(f32))
            lines.append((grid_scale * r.standard_normal((1, c, G[v], 1))).astype(f32))
        return planes, lines

    dp, dl = grids(cd)
    ap, al = grids(ca)

    def lin(out_c, in_c, bias=True):
        b = 1.0 / math.sqrt(in_c)
        w = r.uniform(-b, b, (out_c, in_c)).astype(f32)
        bb = r.uniform(-b, b, (out_c,)).astype(f32) if bias else None
        return w, bb

    basis, _ = lin(app_dim, 3 * ca, bias=False)
    in_mlpC = 2 * view_pe * 3 + 2 * fea_pe * app_dim + 3 + app_dim + (1 if variant == "ref" else 0)
    w1, b1 = lin(featureC, in_mlpC)
    w2, b2 = lin(featureC, featureC)
    w3, b3 = lin(3, featureC)
    b3 = np.zeros_like(b3)  # tensorBase.py:74
    extra = {"variant": variant}
    if variant == "npp":
        # NerfPlusPlus.set_nerfplusplus (models/nerfplusplus.py:147-160) with configs/Scarf.txt:12-15
        bg_freq, bg_view_freq, bg_D, radii, W = 2, 2, 3, 28.0, 128
        pos_dim, dir_dim = 4 + 4 * bg_freq * 2, 3 + 3 * bg_view_freq * 2
        skips = [int(bg_D / 2)]
        base, dim = [], pos_dim
        for i in range(bg_D):                       # MLPNet.__init__ (models/nerfplusplus.py:84-92)
            base.append(lin(W, dim))
            dim = W
            if i in skips and i != bg_D - 1:
                dim += pos_dim
        extra.update(bg_freq=bg_freq, bg_view_freq=bg_view_freq, bg_D=bg_D, radii=radii, bg_base=base,
                     bg_sigma=lin(1, dim), bg_remap=lin(256, dim), bg_rgb0=lin(W // 2, 256 + dir_dim),
                     bg_rgb1=lin(3, W // 2))
    if variant == "ref":
        for name, oc in (("normal", 3), ("diffuse", 3), ("specular", 1), ("rho", 1)):
            w, b = lin(oc, 3 * ca)
            extra[name + "_w"], extra[name + "_b"] = w, b
    return ModelParams(extra=extra, aabb=aabb, gridSize=G, density_plane=dp, density_line=dl, app_plane=ap, app_line=al,
                       basis_mat=basis, mlp_w=[w1, w2, w3], mlp_b=[b1, b2, b3], near_far=tuple(near_far),
                       density_shift=density_shift, view_pe=view_pe, fea_pe=fea_pe, app_dim=app_dim,
                       featureC=featureC, density_n_comp=(cd,) * 3, app_n_comp=(ca,) * 3)


def camera_pose(azimuth: float, elevation: float, radius: float = 12.0) -> np.ndarray:
    """c2w [3,4]: camera on a sphere looking at the origin, -z forward, +y up (blender/OpenGL style)."""
    ce, se = math.cos(elevation), math.sin(elevation)
    ca, sa = math.cos(azimuth), math.sin(azimuth)
    pos = np.array([radius * ce * ca, radius * ce * sa, radius * se], dtype=np.float64)
    fwd = -pos / np.linalg.norm(pos)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    cam_up = np.cross(right, fwd)
    # camera axes: x=right, y=up, z=-forward
    R = np.stack([right, cam_up, -fwd], axis=1)
    return np.concatenate([R, pos[:, None]], axis=1)


def frame_rays(H: int = 800, W: int = 800, azimuth: float = 0.7, elevation: float = 0.5,
               radius: float = 12.0, camera_angle_x: float = 0.6911) -> np.ndarray:
    """all_rays [H*W,6] fp32 = (origin, unit direction), row-major over (row j, column i).

    Follows get_ray_directions (ray_utils.py:81-104: pixel centre +0.5, (-(i-cx)/f, (j-cy)/f, -1)),
    the focal of blender.py:66-67, normalisation blender.py:75 and get_rays (rays_d = dir @ c2w[:3,:3].T).
    """
    focal = 0.5 * 800 / math.tan(0.5 * camera_angle_x)
    focal *= W / 800
    i = (np.arange(W, dtype=np.float32) + np.float32(0.5))[None, :].repeat(H, 0)
    j = (np.arange(H, dtype=np.float32) + np.float32(0.5))[:, None].repeat(W, 1)
    f = np.float32(focal)
    dirs = np.stack([-(i - np.float32(W / 2)) / f, (j - np.float32(H / 2)) / f, -np.ones_like(i)], -1)
    dirs = dirs / np.linalg.norm(dirs, axis=-1, keepdims=True).astype(np.float32)
    c2w = camera_pose(azimuth, elevation, radius).astype(np.float32)
    rays_d = (dirs.reshape(-1, 3) @ c2w[:3, :3].T).astype(np.float32)
    rays_o = np.broadcast_to(c2w[:3, 3], rays_d.shape).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([rays_o, rays_d], 1), dtype=np.float32)


def subset_rays(n: int, seed: int = SEED_BASE, **frame_kw) -> np.ndarray:
    """n rays drawn from the frame by seeded permutation (seed+1)."""
    rays = frame_rays(**frame_kw)
    idx = _rng(1, seed).permutation(rays.shape[0])[:n]
    return np.ascontiguousarray(rays[idx])


def ball_alpha_volume(res, radius: float = 3.5, bbox: float = 5.0) -> np.ndarray:
    """{0,1} fp32 volume [D(z),H(y),W(x)]: node (x,y,z) set iff its position lies inside the ball.

    Node positions are the align_corners lattice linspace(-bbox,bbox,res) used by
    getDenseAlpha (tensorBase.py:367-376); layout [z,y,x] as after the transpose in tensorBase.py:389-396.
    """
    if isinstance(res, int):
        res = (res, res, res)
    xs = np.linspace(-bbox, bbox, res[0], dtype=np.float32)
    ys = np.linspace(-bbox, bbox, res[1], dtype=np.float32)
    zs = np.linspace(-bbox, bbox, res[2], dtype=np.float32)
    r2 = zs[:, None, None] ** 2 + ys[None, :, None] ** 2 + xs[None, None, :] ** 2
    return (r2 < np.float32(radius * radius)).astype(np.float32)


def jitter(n: int, seed: int = SEED_BASE) -> np.ndarray:
    """U[0,1) per-ray march jitter (seed+2), stands in for jt.rand_like(rng[:, [0]]) tensorBase.py:353."""
    return _rng(2, seed).random(n, dtype=np.float32)


def npp_rand(n: int, S: int, seed: int = SEED_BASE):
    """U[0,1) draws consumed by NerfPlusPlus.perturb_samples (models/nerfplusplus.py:196-205):
    foreground [n,S] (seed+4) and background [n,512] (seed+5); the reference jitters even at eval."""
    return _rng(4, seed).random((n, S), dtype=np.float32), _rng(5, seed).random((n, 512), dtype=np.float32)


def target_rgb(n: int, seed: int = SEED_BASE) -> np.ndarray:
    """U[0,1) training target colours (seed+3) for train.py:228's MSE."""
    return _rng(3, seed).random((n, 3), dtype=np.float32)


# Occupancy regimes of SURVEY.md §8d: reference init never reaches the MLP (weight < 1e-4 everywhere).
REGIMES = {
    "R0": dict(density_shift=-10.0, mask=False),   # reference init exactly as written (config 1)
    "R1": dict(density_shift=0.0, mask=True),      # surface-like: ball mask, dense
    "R2": dict(density_shift=-3.0, mask=True),     # fog: ball mask, many weighted samples per ray
}


def make_case(G: int, n_rays: int, regime: str = "R1", mask_res: int | None = None, seed: int = SEED_BASE,
              train: bool = False, full_frame: bool = False, azimuth: float = 0.7, **model_kw):
    """One seeded test/bench case -> dict(model, rays, alpha_volume|None, alpha_aabb, jitter|None, target|None)."""
    reg = REGIMES[regime]
    model = make_model(G, seed=seed, density_shift=reg["density_shift"], **model_kw)
    if full_frame:
        rays = frame_rays(azimuth=azimuth)
        if n_rays and n_rays < rays.shape[0]:
            rays = np.ascontiguousarray(rays[:n_rays])
    else:
        rays = subset_rays(n_rays, seed=seed, azimuth=azimuth)
    vol = None
    if reg["mask"]:
        vol = ball_alpha_volume(mask_res if mask_res else (128 if max(model.gridSize) <= 128 else 200))
    return dict(model=model, rays=rays, alpha_volume=vol, alpha_aabb=model.aabb.copy(),
                jitter=jitter(rays.shape[0], seed) if train else None,
                target=target_rgb(rays.shape[0], seed) if train else None, regime=regime)
