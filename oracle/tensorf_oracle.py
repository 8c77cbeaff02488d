"""CPU oracle: an op-for-op restatement of tensorf-myc's TensoRF-VM per-ray renderer.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this file.  The product path
(jittor-myc-nerfs_b200/) never imports it and has no CPU fallback.

PARITY STATUS: "parity unpinned" with respect to Jittor itself.  The reference runs on Jittor,
which is absent from this image and cannot be installed (no network), and the reference ships no
tests or golden vectors for this path.  What pins this file instead:
  * tests/golden/make_golden.py executes the reference's UNMODIFIED python modules
    (/root/reference/tensorf-myc/models/tensorBase.py, tensoRF.py) over oracle/jt_shim (a torch-CPU
    stand-in for the ~40 Jittor calls they make) and commits the outputs as tests/golden/*.npz;
    tests/test_oracle_golden.py checks this restatement against those vectors (bit-exact masks,
    <=1e-6 on colour).  That pins the reference's own python logic (op order, indexing, quirks);
  * Jittor's op numerics remain ASSUMED (A1-A3 below, each one a switch);
  * a second, independent evaluation through torch.nn.functional.grid_sample / torch.cumprod /
    F.softplus (OracleOptions.torch_native()) must agree with the explicit restatement.

Every function cites the reference lines it follows (paths relative to /root/reference/tensorf-myc/).

Assumptions about Jittor:
  A1 nn.grid_sample(bilinear, zeros, align_corners=True): u = ((c+1)/2)*(size-1); f = floor(u);
     weights (f+1-u) and (u-f) per axis, products formed first, out-of-range taps read 0.
  A2 jt.cumprod(x, dim) = exp(cumsum(log(x), dim)).
  A3 nn.softplus(x, beta=1, threshold=20) = log(1 + exp(min(x,20))) + max(x-20, 0)
     (older 1.3.x: log(1 + exp(x)); identical for x <= 20).
  A4 nn.Linear: y = x @ W.T + b with W [out,in].
  A5 boolean-mask get/set-item compacts in row-major (ray-major, sample-minor) order.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

MAT_MODE = ((0, 1), (0, 2), (1, 2))  # models/tensorBase.py:168
VEC_MODE = (2, 1, 0)                 # models/tensorBase.py:169


@dataclass
class OracleOptions:
    grid_sample: str = "explicit"   # "explicit" (A1 spelled out) | "torch" (F.grid_sample)
    cumprod: str = "logspace"       # "logspace" (A2) | "product" (running product, torch.cumprod)
    softplus: str = "jittor"        # "jittor" (A3 clamped) | "naive" (log(1+exp x)) | "log1p" (torch F.softplus)

    @staticmethod
    def torch_native() -> "OracleOptions":
        return OracleOptions(grid_sample="torch", cumprod="product", softplus="log1p")


# --------------------------------------------------------------------------------------------
# Jittor op restatements
# --------------------------------------------------------------------------------------------
def _unnormalize(c, size):
    # A1, align_corners=True
    return ((c + 1) / 2) * (size - 1)


def grid_sample_2d(img, gx, gy, impl="explicit"):
    """img [C,H,W]; gx -> W axis, gy -> H axis, both [N] in [-1,1]. Returns [C,N]."""
    C, H, W = img.shape
    if impl == "torch":
        grid = torch.stack([gx, gy], -1).view(1, -1, 1, 2)
        return F.grid_sample(img[None], grid, mode="bilinear", padding_mode="zeros",
                             align_corners=True).view(C, -1)
    ix = _unnormalize(gx, W)
    iy = _unnormalize(gy, H)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    x1 = x0 + 1
    y1 = y0 + 1
    nw = (x1 - ix) * (y1 - iy)
    ne = (ix - x0) * (y1 - iy)
    sw = (x1 - ix) * (iy - y0)
    se = (ix - x0) * (iy - y0)

    def tap(xi, yi):
        inb = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
        v = img[:, yi.clamp(0, H - 1).long(), xi.clamp(0, W - 1).long()]
        return v * inb.to(img.dtype)

    return tap(x0, y0) * nw + tap(x1, y0) * ne + tap(x0, y1) * sw + tap(x1, y1) * se


def grid_sample_3d(vol, gx, gy, gz, impl="explicit"):
    """vol [D,H,W]; gx->W, gy->H, gz->D. Returns [N]."""
    D, H, W = vol.shape
    if impl == "torch":
        grid = torch.stack([gx, gy, gz], -1).view(1, -1, 1, 1, 3)
        return F.grid_sample(vol[None, None], grid, mode="bilinear", padding_mode="zeros",
                             align_corners=True).view(-1)
    ix = _unnormalize(gx, W)
    iy = _unnormalize(gy, H)
    iz = _unnormalize(gz, D)
    x0, y0, z0 = torch.floor(ix), torch.floor(iy), torch.floor(iz)
    out = torch.zeros_like(ix)
    for cz in (0, 1):
        wz = (z0 + 1 - iz) if cz == 0 else (iz - z0)
        for cy in (0, 1):
            wy = (y0 + 1 - iy) if cy == 0 else (iy - y0)
            for cx in (0, 1):
                wx = (x0 + 1 - ix) if cx == 0 else (ix - x0)
                xi, yi, zi = x0 + cx, y0 + cy, z0 + cz
                inb = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1) & (zi >= 0) & (zi <= D - 1)
                v = vol[zi.clamp(0, D - 1).long(), yi.clamp(0, H - 1).long(), xi.clamp(0, W - 1).long()]
                out = out + (v * inb.to(vol.dtype)) * (wx * wy * wz)
    return out


def jt_cumprod(x, dim, impl="logspace"):
    if impl == "logspace":                       # A2
        return torch.exp(torch.cumsum(torch.log(x), dim))
    return torch.cumprod(x, dim)


def jt_softplus(x, impl="jittor"):
    if impl == "jittor":                         # A3 (clamped variant)
        return torch.log(1 + torch.exp(torch.clamp(x, max=20.0))) + torch.clamp(x - 20.0, min=0.0)
    if impl == "naive":
        return torch.log(1 + torch.exp(x))
    return F.softplus(x)


def positional_encoding(positions, freqs):
    """models/tensorBase.py:9-15 -- channel-major, frequency-minor; [sin | cos]."""
    freq_bands = (2 ** torch.arange(freqs)).to(positions.dtype)
    pts = (positions[..., None] * freq_bands).reshape(positions.shape[:-1] + (freqs * positions.shape[-1],))
    return torch.cat([torch.sin(pts), torch.cos(pts)], dim=-1)


def raw2alpha(sigma, dist, cumprod_impl="logspace"):
    """models/tensorBase.py:17-24."""
    alpha = 1. - torch.exp(-sigma * dist)
    T = jt_cumprod(torch.cat([torch.ones((alpha.shape[0], 1), dtype=alpha.dtype), 1. - alpha + 1e-10], -1), -1,
                   cumprod_impl)
    weights = alpha * T[:, :-1]
    return alpha, weights, T[:, -1:]


class AlphaGridMask:
    """models/tensorBase.py:39-59."""

    def __init__(self, aabb, alpha_volume, dtype=torch.float32, impl="explicit"):
        self.aabb = torch.as_tensor(np.asarray(aabb), dtype=dtype)
        self.aabbSize = self.aabb[1] - self.aabb[0]
        self.invgridSize = 1.0 / self.aabbSize * 2
        self.alpha_volume = torch.as_tensor(np.asarray(alpha_volume), dtype=dtype)
        self.alpha_volume = self.alpha_volume.view(*self.alpha_volume.shape[-3:])
        self.gridSize = [self.alpha_volume.shape[-1], self.alpha_volume.shape[-2], self.alpha_volume.shape[-3]]
        self.impl = impl

    def normalize_coord(self, xyz):
        return (xyz - self.aabb[0]) * self.invgridSize - 1

    def sample_alpha(self, xyz):
        c = self.normalize_coord(xyz)
        if c.shape[0] == 0:
            return torch.zeros((0,), dtype=c.dtype)
        return grid_sample_3d(self.alpha_volume, c[:, 0], c[:, 1], c[:, 2], self.impl)


class OracleTensorVMSplit:
    """TensorVMSplit + TensorBase.execute (models/tensoRF.py:141-244, models/tensorBase.py:140-224,340-360,444-536)."""

    def __init__(self, params, alpha_volume=None, alpha_aabb=None, dtype=torch.float32,
                 opts: OracleOptions | None = None, requires_grad=False):
        self.opts = opts or OracleOptions()
        self.dtype = dtype
        p = params
        t = lambda a: torch.tensor(np.asarray(a), dtype=dtype)
        self.aabb = t(p.aabb)
        self.near_far = p.near_far
        self.density_shift = p.density_shift
        self.distance_scale = p.distance_scale
        self.rayMarch_weight_thres = p.rayMarch_weight_thres
        self.step_ratio = p.step_ratio
        self.view_pe, self.fea_pe = p.view_pe, p.fea_pe
        self.fea2denseAct = p.fea2denseAct
        self.density_plane = [t(a) for a in p.density_plane]
        self.density_line = [t(a) for a in p.density_line]
        self.app_plane = [t(a) for a in p.app_plane]
        self.app_line = [t(a) for a in p.app_line]
        self.basis_mat = t(p.basis_mat)
        self.mlp_w = [t(a) for a in p.mlp_w]
        self.mlp_b = [t(a) for a in p.mlp_b]
        if requires_grad:
            for x in self.parameters():
                x.requires_grad_(True)
        self.update_stepSize(p.gridSize)
        self.alphaMask = None
        if alpha_volume is not None:
            self.alphaMask = AlphaGridMask(alpha_aabb if alpha_aabb is not None else p.aabb, alpha_volume,
                                           dtype=dtype, impl=self.opts.grid_sample)

    # -- bookkeeping ------------------------------------------------------------------------
    def named_parameters(self):
        out = {}
        for k in range(3):
            out[f"density_plane.{k}"] = self.density_plane[k]
            out[f"density_line.{k}"] = self.density_line[k]
            out[f"app_plane.{k}"] = self.app_plane[k]
            out[f"app_line.{k}"] = self.app_line[k]
        out["basis_mat.weight"] = self.basis_mat
        for i, li in enumerate((0, 2, 4)):
            out[f"renderModule.mlp.{li}.weight"] = self.mlp_w[i]
            out[f"renderModule.mlp.{li}.bias"] = self.mlp_b[i]
        return out

    def parameters(self):
        return list(self.named_parameters().values())

    def update_stepSize(self, gridSize):
        """models/tensorBase.py:197-209."""
        self.aabbSize = self.aabb[1] - self.aabb[0]
        self.invaabbSize = 2.0 / self.aabbSize
        self.gridSize = torch.tensor([int(g) for g in gridSize], dtype=torch.int32)
        self.units = self.aabbSize / (self.gridSize - 1)
        self.stepSize = torch.mean(self.units) * self.step_ratio
        self.aabbDiag = torch.sqrt(torch.sum(torch.pow(self.aabbSize, 2)))
        self.nSamples = int((self.aabbDiag / self.stepSize).item()) + 1

    def normalize_coord(self, xyz):
        """models/tensorBase.py:223-224."""
        return (xyz - self.aabb[0]) * self.invaabbSize - 1

    # -- sampling ---------------------------------------------------------------------------
    def sample_ray(self, rays_o, rays_d, is_train=True, N_samples=-1, jitter=None):
        """models/tensorBase.py:340-360. `jitter` [n] replaces jt.rand_like(rng[:, [0]])."""
        N_samples = N_samples if N_samples > 0 else self.nSamples
        stepsize = self.stepSize
        near, far = self.near_far
        vec = torch.where(rays_d == 0, torch.full_like(rays_d, 1e-6), rays_d)
        rate_a = (self.aabb[1] - rays_o) / vec
        rate_b = (self.aabb[0] - rays_o) / vec
        t_min = torch.minimum(rate_a, rate_b).amax(-1).clamp(min=near, max=far)
        rng = torch.arange(N_samples)[None].to(self.dtype)
        if is_train:
            rng = rng.repeat(rays_d.shape[-2], 1)
            rng = rng + jitter.to(self.dtype).view(-1, 1)
        step = stepsize * rng
        interpx = (t_min[..., None] + step)
        rays_pts = rays_o[..., None, :] + rays_d[..., None, :] * interpx[..., None]
        mask_outbbox = ((self.aabb[0] > rays_pts) | (rays_pts > self.aabb[1])).any(dim=-1)
        return rays_pts, interpx, ~mask_outbbox

    # -- factor-grid gathers ----------------------------------------------------------------
    def _coords(self, xyz):
        cp = [(xyz[..., MAT_MODE[k][0]], xyz[..., MAT_MODE[k][1]]) for k in range(3)]
        cl = [xyz[..., VEC_MODE[k]] for k in range(3)]
        return cp, cl

    def compute_densityfeature(self, xyz):
        """models/tensoRF.py:209-225."""
        cp, cl = self._coords(xyz.detach())
        g = self.opts.grid_sample
        sigma_feature = torch.zeros((xyz.shape[0],), dtype=self.dtype)
        for k in range(3):
            plane = grid_sample_2d(self.density_plane[k][0], cp[k][0], cp[k][1], g)
            line = grid_sample_2d(self.density_line[k][0], torch.zeros_like(cl[k]), cl[k], g)
            sigma_feature = sigma_feature + torch.sum(plane * line, dim=0)
        return sigma_feature

    def compute_app_vector(self, xyz):
        """The [M,144] product vector fed to basis_mat (models/tensoRF.py:228-243)."""
        cp, cl = self._coords(xyz.detach())
        g = self.opts.grid_sample
        planes, lines = [], []
        for k in range(3):
            planes.append(grid_sample_2d(self.app_plane[k][0], cp[k][0], cp[k][1], g))
            lines.append(grid_sample_2d(self.app_line[k][0], torch.zeros_like(cl[k]), cl[k], g))
        return (torch.cat(planes) * torch.cat(lines)).T

    def compute_appfeature(self, xyz):
        """models/tensoRF.py:228-244; basis_mat = Linear(144, 27, bias=False) (A4)."""
        return self.compute_app_vector(xyz) @ self.basis_mat.T

    def feature2density(self, f):
        """models/tensorBase.py:444-448."""
        if self.fea2denseAct == "softplus":
            return jt_softplus(f + self.density_shift, self.opts.softplus)
        return torch.relu(f)

    def renderModule(self, pts, viewdirs, features):
        """MLPRender_Fea.execute, models/tensorBase.py:76-86."""
        indata = [features, viewdirs]
        if self.fea_pe > 0:
            indata += [positional_encoding(features, self.fea_pe)]
        if self.view_pe > 0:
            indata += [positional_encoding(viewdirs, self.view_pe)]
        x = torch.cat(indata, dim=-1)
        x = torch.relu(x @ self.mlp_w[0].T + self.mlp_b[0])
        x = torch.relu(x @ self.mlp_w[1].T + self.mlp_b[1])
        x = x @ self.mlp_w[2].T + self.mlp_b[2]
        return torch.sigmoid(x)

    # -- the per-chunk pipeline -------------------------------------------------------------
    def execute(self, rays_chunk, white_bg=True, is_train=False, ndc_ray=False, N_samples=-1,
                jitter=None, stages=None):
        """models/tensorBase.py:476-536 (ndc_ray=False branch). `stages` (dict) collects intermediates."""
        assert not ndc_ray
        rays_chunk = rays_chunk.to(self.dtype)
        viewdirs = rays_chunk[:, 3:6]
        xyz_sampled, z_vals, ray_valid = self.sample_ray(rays_chunk[:, :3], viewdirs, is_train=is_train,
                                                         N_samples=N_samples, jitter=jitter)
        dists = torch.cat((z_vals[:, 1:] - z_vals[:, :-1], torch.zeros_like(z_vals[:, :1])), dim=-1)
        if dists.shape[0] != xyz_sampled.shape[0]:
            dists = dists.expand(xyz_sampled.shape[0], -1)
        viewdirs = viewdirs.view(-1, 1, 3).expand(xyz_sampled.shape)
        if stages is not None:
            stages["bbox_valid"] = ray_valid.clone()
            stages["z_vals"] = z_vals.expand(xyz_sampled.shape[0], -1).clone()

        if self.alphaMask is not None:
            alphas = self.alphaMask.sample_alpha(xyz_sampled[ray_valid])
            alpha_mask = alphas > 0
            ray_invalid = ~ray_valid
            tmp = ray_invalid[ray_valid]
            tmp |= (~alpha_mask)
            ray_invalid[ray_valid] = tmp
            ray_valid = ~ray_invalid

        sigma = torch.zeros(xyz_sampled.shape[:-1], dtype=self.dtype)
        rgb = torch.zeros((*xyz_sampled.shape[:2], 3), dtype=self.dtype)

        if ray_valid.any():
            xyz_sampled = self.normalize_coord(xyz_sampled)
            sigma_feature = self.compute_densityfeature(xyz_sampled[ray_valid])
            validsigma = self.feature2density(sigma_feature)
            sigma = sigma.index_put((ray_valid,), validsigma)

        alpha, weight, bg_weight = raw2alpha(sigma, dists * self.distance_scale, self.opts.cumprod)
        app_mask = weight > self.rayMarch_weight_thres

        if app_mask.any():
            valid_rgbs = self.shade(xyz_sampled[app_mask], viewdirs[app_mask], weight[app_mask])
            rgb = rgb.index_put((app_mask,), valid_rgbs)

        acc_map = torch.sum(weight, -1)
        rgb_map_raw = torch.sum(weight[..., None] * rgb, -2)
        if white_bg:
            rgb_map_raw = rgb_map_raw + (1. - acc_map[..., None])
        rgb_map = rgb_map_raw.clamp(0, 1)

        with torch.no_grad():
            depth_map = torch.sum(weight * z_vals, -1)
            depth_map = depth_map + (1. - acc_map) * rays_chunk[..., -1]   # column 5 = d_z: reference quirk

        self._alpha_live = alpha        # the reference hands the live Var to NerfPlusPlus.execute (additional_output, :533-534)
        if stages is not None:
            stages.update(ray_valid=ray_valid, sigma=sigma.detach(), alpha=alpha.detach(), weight=weight.detach(),
                          bg_weight=bg_weight.detach(), app_mask=app_mask, rgb=rgb.detach(),
                          acc_map=acc_map.detach(), rgb_map_raw=rgb_map_raw.detach())
        return rgb_map, depth_map

    def __call__(self, *a, **k):
        return self.execute(*a, **k)

    def shade(self, xyz, viewdirs, weight):
        """models/tensorBase.py:516-517: appearance features -> MLPRender_Fea."""
        app_features = self.compute_appfeature(xyz)
        return self.renderModule(xyz, viewdirs, app_features)

    # -- §8f-1: dense alpha on arbitrary points ------------------------------------------------
    def compute_alpha(self, xyz_locs, length=1.0):
        """models/tensorBase.py:451-473."""
        xyz_locs = xyz_locs.to(self.dtype)
        if self.alphaMask is not None:
            alpha_mask = self.alphaMask.sample_alpha(xyz_locs) > 0
        else:
            alpha_mask = torch.ones_like(xyz_locs[:, 0]).bool()
        sigma = torch.zeros(xyz_locs.shape[:-1], dtype=self.dtype)
        if alpha_mask.any():
            xyz_sampled = self.normalize_coord(xyz_locs[alpha_mask])
            sigma[alpha_mask] = self.feature2density(self.compute_densityfeature(xyz_sampled))
        return 1 - torch.exp(-sigma * length).view(xyz_locs.shape[:-1])


def jt_normalize(x, dim=-1, eps=1e-30):
    """A7: jt.normalize(x, p=2, dim) = x / sqrt(max(sum(x^2, dim), eps)) (jittor/misc.py, as recalled)."""
    return x / torch.sqrt(torch.clamp(torch.sum(x * x, dim, keepdim=True), min=eps))


class OracleREFTensoRF(OracleTensorVMSplit):
    """REFTensoRF (models/REFTensoRF.py:64-256): Ref-NeRF style heads on the 144-vector, reflected view
    direction into MLPRender_Fea_Ref (:5-29), rgb = tint * clamp(rgb_s, 0) + rgb_d, normal penalty."""

    def __init__(self, params, *a, **k):
        super().__init__(params, *a, **k)
        t = lambda a_: torch.tensor(np.asarray(a_), dtype=self.dtype)
        e = params.extra
        self.heads = {n: (t(e[n + "_w"]), t(e[n + "_b"])) for n in ("normal", "diffuse", "specular", "rho")}
        if self.basis_mat.requires_grad:
            for w, b in self.heads.values():
                w.requires_grad_(True)
                b.requires_grad_(True)
        self.penalty = torch.zeros((), dtype=self.dtype)

    def named_parameters(self):
        out = super().named_parameters()
        for n, (w, b) in getattr(self, "heads", {}).items():
            out[f"{n}_linear.weight"], out[f"{n}_linear.bias"] = w, b
        return out

    def shade(self, xyz, viewdirs, weight):
        """REFTensoRF.compute_appfeature (:107-133) + execute (:216-238)."""
        h = self.compute_app_vector(xyz)
        lin = lambda n: h @ self.heads[n][0].T + self.heads[n][1]
        app_features = h @ self.basis_mat.T
        normal_vector, rgb_d = lin("normal"), lin("diffuse")
        specular_tint = torch.relu(lin("specular"))
        rho = torch.relu(lin("rho"))                     # only ever passed on as k = 1/rho, which the MLP ignores
        normal_vector = jt_normalize(normal_vector, dim=-1)
        d = -viewdirs
        dot_product = (d * normal_vector).sum(dim=1)[:, None]
        reflection = 2 * dot_product * normal_vector - d
        # MLPRender_Fea_Ref.execute(pts, viewdirs=reflection, features, dot_product=-dot, k)
        indata = [-dot_product, app_features, reflection]
        if self.fea_pe > 0:
            indata += [positional_encoding(app_features, self.fea_pe)]
        if self.view_pe > 0:
            indata += [positional_encoding(reflection, self.view_pe)]
        x = torch.cat(indata, dim=-1)
        x = torch.relu(x @ self.mlp_w[0].T + self.mlp_b[0])
        x = torch.relu(x @ self.mlp_w[1].T + self.mlp_b[1])
        rgb_s = torch.sigmoid(x @ self.mlp_w[2].T + self.mlp_b[2])
        valid_rgbs = specular_tint * torch.clamp(rgb_s, min=0) + rgb_d
        penalty = torch.relu(-dot_product) ** 2
        self.penalty = torch.sum(weight * penalty.squeeze(-1), -1)
        return valid_rgbs


HUGE_NUMBER = 1e10     # models/nerfplusplus.py:3-4
TINY_NUMBER = 1e-6


class OracleNerfPlusPlus(OracleTensorVMSplit):
    """NerfPlusPlus (models/nerfplusplus.py:143-318): sphere-bounded, always-jittered foreground sampling and a
    512-sample inverted-sphere background MLP (Embedder :7-56, MLPNet :66-140).  The U[0,1) draws of
    perturb_samples are injected: `fg_rand` [n,S] and `bg_rand` [n,512]."""

    def __init__(self, params, *a, **k):
        super().__init__(params, *a, **k)
        e = params.extra
        t = lambda x: torch.tensor(np.asarray(x), dtype=self.dtype)
        self.radii, self.bg_freq, self.bg_view_freq, self.bg_D = e["radii"], e["bg_freq"], e["bg_view_freq"], e["bg_D"]
        self.bg = {k_: [(t(w), t(b)) for w, b in (v if isinstance(v, list) else [v])]
                   for k_, v in (("base", e["bg_base"]), ("sigma", e["bg_sigma"]), ("remap", e["bg_remap"]),
                                 ("rgb0", e["bg_rgb0"]), ("rgb1", e["bg_rgb1"]))}
        self.skips = [int(self.bg_D / 2)]
        self.fg_rand = self.bg_rand = None
        if k.get("requires_grad") or (len(a) > 4 and a[4]):
            for x in self.bg_parameters().values():
                x.requires_grad_(True)

    def bg_parameters(self):
        """bg_net parameters under the reference's state_dict names (MLPNet, nerfplusplus.py:86-113)."""
        out = {}
        for i, (w, b) in enumerate(self.bg["base"]):
            out[f"bg_net.base_layers.{i}.0.weight"], out[f"bg_net.base_layers.{i}.0.bias"] = w, b
        for name, key, idx in (("sigma_layers", "sigma", 0), ("base_remap_layers", "remap", 0), ("rgb_layers", "rgb0", 0),
                               ("rgb_layers", "rgb1", 2)):
            w, b = self.bg[key][0]
            out[f"bg_net.{name}.{idx}.weight"], out[f"bg_net.{name}.{idx}.bias"] = w, b
        return out

    def named_parameters(self):
        out = super().named_parameters()
        if hasattr(self, "bg"):
            out.update(self.bg_parameters())
        return out

    @staticmethod
    def embed(x, n_freqs):
        """Embedder.execute (:40-56): [x, sin(x f0), cos(x f0), sin(x f1), ...] with f = 2**linspace(0, N-1, N)."""
        out = [x]
        for f in (2.0 ** torch.linspace(0.0, n_freqs - 1, n_freqs)).tolist():
            out += [torch.sin(x * f), torch.cos(x * f)]
        return torch.cat(out, -1)

    def intersect_sphere(self, ray_o, ray_d, radii):
        """:178-194 (radii is passed squared by the caller, :241)."""
        d1 = -torch.sum(ray_d * ray_o, dim=-1) / torch.sum(ray_d * ray_d, dim=-1)
        p = ray_o + d1.unsqueeze(-1) * ray_d
        ray_d_cos = 1. / torch.norm(ray_d, dim=-1)
        p_norm_sq = torch.sum(p * p, dim=-1)
        d2 = torch.sqrt(radii - p_norm_sq) * ray_d_cos
        return d1 + d2

    @staticmethod
    def perturb_samples(z_vals, t_rand):
        """:196-205."""
        mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat([mids, z_vals[..., -1:]], dim=-1)
        lower = torch.cat([z_vals[..., 0:1], mids], dim=-1)
        return lower + (upper - lower) * t_rand

    def depth2pts_outside(self, ray_o, ray_d, depth, radii):
        """:207-237."""
        d1 = -torch.sum(ray_d * ray_o, dim=-1) / torch.sum(ray_d * ray_d, dim=-1)
        p_mid = ray_o + d1.unsqueeze(-1) * ray_d
        p_mid_norm = torch.norm(p_mid, dim=-1)
        ray_d_cos = 1. / torch.norm(ray_d, dim=-1)
        d2 = torch.sqrt(radii * radii - p_mid_norm * p_mid_norm) * ray_d_cos
        p_sphere = ray_o + (d1 + d2).unsqueeze(-1) * ray_d
        rot_axis = torch.cross(ray_o, p_sphere, dim=-1)
        rot_axis = rot_axis / torch.norm(rot_axis, dim=-1, keepdim=True)
        phi = torch.asin(p_mid_norm / radii)
        theta = torch.asin(p_mid_norm * depth / (radii * radii))
        rot_angle = (phi - theta).unsqueeze(-1)
        p_sphere_new = p_sphere * torch.cos(rot_angle) + \
            torch.cross(rot_axis, p_sphere, dim=-1) * torch.sin(rot_angle) + \
            rot_axis * torch.sum(rot_axis * p_sphere, dim=-1, keepdim=True) * (1. - torch.cos(rot_angle))
        return torch.cat((p_sphere_new, depth.unsqueeze(-1)), dim=-1)

    def sample_ray(self, rays_o, rays_d, is_train=True, N_samples=-1, jitter=None):
        """:239-269: depths linear from `near` to the sphere exit, always stratified-jittered."""
        N_samples = N_samples if N_samples > 0 else self.nSamples
        fg_far_depth = self.intersect_sphere(rays_o, rays_d, radii=self.radii * self.radii)
        near, far = self.near_far
        step = (fg_far_depth - near) / (N_samples - 1)
        fg_depth = torch.stack([near + i * step for i in range(N_samples)], dim=-1)
        interpx = self.perturb_samples(fg_depth, self.fg_rand.to(self.dtype))
        rays_pts = rays_o[..., None, :] + rays_d[..., None, :] * interpx[..., None]
        mask_outbbox = ((self.aabb[0] > rays_pts) | (rays_pts > self.aabb[1])).any(dim=-1)
        return rays_pts, interpx, ~mask_outbbox

    def bg_net(self, inp, pos_dim, dir_dim):
        """MLPNet.execute (:115-140)."""
        input_pts = inp[..., :pos_dim]
        lin = lambda x, wb: x @ wb[0].T + wb[1]
        base = torch.relu(lin(input_pts, self.bg["base"][0]))
        for i in range(len(self.bg["base"]) - 1):
            if i in self.skips:
                base = torch.cat((input_pts, base), dim=-1)
            base = torch.relu(lin(base, self.bg["base"][i + 1]))
        sigma = torch.abs(lin(base, self.bg["sigma"][0])).squeeze(-1)
        base_remap = lin(base, self.bg["remap"][0])
        x = torch.relu(lin(torch.cat((base_remap, inp[..., -dir_dim:]), dim=-1), self.bg["rgb0"][0]))
        return torch.sigmoid(lin(x, self.bg["rgb1"][0])), sigma

    def execute(self, rays_chunk, white_bg=False, is_train=False, ndc_ray=False, N_samples=-1, jitter=None,
                stages=None, fg_rand=None, bg_rand=None):
        """:272-318.  white_bg is ignored: the foreground always renders on black (:274)."""
        N_samples = N_samples if N_samples > 0 else self.nSamples
        self.fg_rand, bg_rand = fg_rand, bg_rand.to(self.dtype)
        st = {} if stages is None else stages
        rgb_map, depth_map = super().execute(rays_chunk, False, is_train, ndc_ray, N_samples, stages=st)
        alpha = self._alpha_live        # not detached: bg_lambda carries gradient into the density grids (:277-278)
        bg_lambda = jt_cumprod(1. - alpha + TINY_NUMBER, -1, self.opts.cumprod)[..., -1]
        rays_chunk = rays_chunk.to(self.dtype)
        ray_o, ray_d = rays_chunk[:, :3], rays_chunk[:, 3:6]
        viewdirs = ray_d / torch.norm(ray_d, dim=-1, keepdim=True)
        n, NB = ray_d.shape[0], 512
        bg_z_vals = torch.linspace(0., self.radii, NB).to(self.dtype).view(1, NB).expand(n, NB)
        bg_z_vals = self.perturb_samples(bg_z_vals, bg_rand)
        bg_pts = self.depth2pts_outside(ray_o.unsqueeze(-2).expand(n, NB, 3), ray_d.unsqueeze(-2).expand(n, NB, 3),
                                        bg_z_vals, radii=self.radii)
        pos = self.embed(bg_pts, self.bg_freq)
        dirs = self.embed(viewdirs.unsqueeze(-2).expand(n, NB, 3), self.bg_view_freq)
        inp = torch.flip(torch.cat((pos, dirs), dim=-1), [-2])
        bg_z_vals = torch.flip(bg_z_vals, [-1])
        bg_dists = bg_z_vals[..., :-1] - bg_z_vals[..., 1:]
        bg_dists = torch.cat((bg_dists, HUGE_NUMBER * torch.ones_like(bg_dists[..., 0:1])), dim=-1)
        bg_rgb, bg_sigma = self.bg_net(inp, pos.shape[-1], dirs.shape[-1])
        bg_alpha = 1. - torch.exp(-bg_sigma * bg_dists)
        T = jt_cumprod(1. - bg_alpha + TINY_NUMBER, -1, self.opts.cumprod)[..., :-1]
        T = torch.cat((torch.ones_like(T[..., 0:1]), T), dim=-1)
        bg_weights = bg_alpha * T
        bg_rgb_map = torch.sum(bg_weights.unsqueeze(-1) * bg_rgb, dim=-2)
        bg_lambda = torch.where(bg_lambda > 0.1, bg_lambda, torch.zeros_like(bg_lambda))
        st.update(bg_lambda=bg_lambda.detach(), bg_rgb_map=bg_rgb_map.detach(), fg_rgb_map=rgb_map.detach())
        return rgb_map + bg_lambda.unsqueeze(-1) * bg_rgb_map, depth_map


def make_oracle(case, dtype=torch.float32, opts=None, requires_grad=False):
    v = case["model"].extra.get("variant")
    cls = {"ref": OracleREFTensoRF, "npp": OracleNerfPlusPlus}.get(v, OracleTensorVMSplit)
    return cls(case["model"], case["alpha_volume"], case["alpha_aabb"], dtype=dtype, opts=opts,
               requires_grad=requires_grad)


def OctreeRender_trilinear_fast(rays, tensorf, chunk=4096, N_samples=-1, ndc_ray=False, white_bg=True,
                                is_train=False, device="cpu", jitter=None):
    """renderer.py:12-27 (the jt.sync_all()/jt.gc() per chunk have no CPU counterpart)."""
    rgbs, depth_maps = [], []
    N_rays_all = rays.shape[0]
    for chunk_idx in range(N_rays_all // chunk + int(N_rays_all % chunk > 0)):
        sl = slice(chunk_idx * chunk, (chunk_idx + 1) * chunk)
        rgb_map, depth_map = tensorf(rays[sl], is_train=is_train, white_bg=white_bg, ndc_ray=ndc_ray,
                                     N_samples=N_samples, jitter=None if jitter is None else jitter[sl])
        rgbs.append(rgb_map)
        depth_maps.append(depth_map)
    return torch.cat(rgbs), None, torch.cat(depth_maps), None, None


# --------------------------------------------------------------------------------------------
# helpers used by tests / bench
# --------------------------------------------------------------------------------------------
def run_case(case, dtype=torch.float32, opts=None, N_samples=-1, white_bg=True, want_stages=True,
             chunk=4096):
    """Forward a fixtures.make_case() dict; returns dict of numpy outputs (+ per-sample stages)."""
    m = make_oracle(case, dtype=dtype, opts=opts)
    rays = torch.from_numpy(case["rays"])
    jit = None if case.get("jitter") is None else torch.from_numpy(case["jitter"])
    is_train = jit is not None
    outs, st_all = [], []
    penalty = 0.0
    for s in range(0, rays.shape[0], chunk):
        st = {} if want_stages else None
        kw = {}
        if case.get("fg_rand") is not None:
            kw = dict(fg_rand=torch.from_numpy(case["fg_rand"][s:s + chunk]),
                      bg_rand=torch.from_numpy(case["bg_rand"][s:s + chunk]))
        with torch.no_grad():
            rgb, depth = m(rays[s:s + chunk], white_bg=white_bg, is_train=is_train, N_samples=N_samples,
                           jitter=None if jit is None else jit[s:s + chunk], stages=st, **kw)
        outs.append((rgb, depth))
        st_all.append(st)
        penalty += float(getattr(m, "penalty", 0.0))
    res = dict(penalty=penalty, rgb_map=torch.cat([o[0] for o in outs]).numpy(), depth_map=torch.cat([o[1] for o in outs]).numpy(),
               nSamples=m.nSamples, stepSize=float(m.stepSize))
    if want_stages:
        for k in st_all[0]:
            res[k] = torch.cat([s[k] for s in st_all]).numpy()
    return res


def work_counts(stages):
    """M_in / M_v / M_a of SURVEY.md §8d from a stages dict."""
    return dict(n=int(stages["bbox_valid"].shape[0]), M_in=int(stages["bbox_valid"].sum()),
                M_v=int(stages["ray_valid"].sum()), M_a=int(stages["app_mask"].sum()))


def backward_case(case, d_rgb_map=None, dtype=torch.float64, opts=None, N_samples=-1, white_bg=True, penalty_weight=0.0):
    """Gradients of sum(rgb_map * d_rgb_map) (or of train.py:228's MSE against case['target'] when
    d_rgb_map is None) w.r.t. every parameter, in the reference's NCHW shapes (row a12)."""
    m = make_oracle(case, dtype=dtype, opts=opts, requires_grad=True)
    rays = torch.from_numpy(case["rays"])
    jit = None if case.get("jitter") is None else torch.from_numpy(case["jitter"])
    kw = {}
    if case.get("fg_rand") is not None:      # NerfPlusPlus: injected draws of perturb_samples
        kw = dict(fg_rand=torch.from_numpy(case["fg_rand"]), bg_rand=torch.from_numpy(case["bg_rand"]))
    rgb, depth = m(rays, white_bg=white_bg, is_train=jit is not None, N_samples=N_samples, jitter=jit, **kw)
    if d_rgb_map is None:
        tgt = torch.from_numpy(case["target"]).to(dtype)
        loss = torch.mean((rgb - tgt) ** 2)
    else:
        loss = torch.sum(rgb * torch.as_tensor(d_rgb_map, dtype=dtype))
    penalty = None
    if penalty_weight:
        # train.py:253-255: total_loss += normal_vector_penalty_weight * tensorf.penalty (REFTensoRF only)
        penalty = m.penalty
        loss = loss + penalty_weight * penalty
    loss.backward()
    grads = {k: (v.grad.detach().numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in
             m.named_parameters().items()}
    return dict(loss=float(loss), rgb_map=rgb.detach().numpy(), grads=grads,
                penalty=None if penalty is None else float(penalty.detach()))
