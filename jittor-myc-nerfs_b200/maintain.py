"""Host mirror of the reference methods either side of the ray path: updateAlphaMask / getDenseAlpha,
filtering_rays, upsample_volume_grid / shrink (tensorf-myc/models/tensorBase.py:366-441, models/tensoRF.py:248-314) and
ray generation (dataLoader/ray_utils.py:81-153).  Same names, arguments and return values as the reference; the arithmetic
runs in libtvmrender.so (csrc/tvm_maintain.cu); torch only owns the memory and does the index plumbing (slicing,
boolean compaction)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _i3(v):
    return (C.c_int32 * 3)(*[int(x) for x in v])


def _linspace01(i, n):
    """torch.linspace(0, 1, n)[i] in fp32 (evaluated from both ends), as csrc/tvm_maintain.cu::linspace01."""
    f32 = np.float32
    if n <= 1:
        return f32(0)
    step = f32(1) / f32(n - 1)
    return f32(step * f32(i)) if i < n // 2 else f32(f32(1) - f32(step * f32(n - 1 - i)))


def lattice_point(aabb, grid, idx):
    """dense_xyz[idx] = aabb[0] * (1 - s) + aabb[1] * s (tensorBase.py:376) in fp32."""
    f32 = np.float32
    a = np.asarray(aabb, dtype=f32).reshape(2, 3)
    out = np.zeros(3, f32)
    for k in range(3):
        s = _linspace01(int(idx[k]), int(grid[k]))
        out[k] = f32(f32(a[0, k] * f32(f32(1) - s)) + f32(a[1, k] * s))
    return out


class MaintainMixin:
    """Mixed into TensorVMSplit (tensorf.py)."""

    # ---- §8f-1 ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _dense_alpha_zyx(self, gridSize):
        g = [int(x) for x in gridSize]
        alpha = torch.empty((g[2], g[1], g[0]), dtype=torch.float32, device=self.device)
        model = self._model()
        L.check(L.load().tvm_dense_alpha(C.byref(model), _i3(g), float(self.stepSize), _ptr(alpha), _stream_ptr()),
                "tvm_dense_alpha")
        return alpha

    @torch.no_grad()
    def getDenseAlpha(self, gridSize=None):
        """tensorBase.py:366-384 -> (alpha [Gx,Gy,Gz], dense_xyz [Gx,Gy,Gz,3])."""
        gridSize = self.gridSize if gridSize is None else gridSize
        g = [int(x) for x in (gridSize.tolist() if torch.is_tensor(gridSize) else gridSize)]
        alpha = self._dense_alpha_zyx(g).permute(2, 1, 0)
        lin = [torch.linspace(0, 1, n, device=self.device) for n in g]
        samples = torch.stack(torch.meshgrid(*lin, indexing="ij"), -1)
        aabb = self.aabb.to(self.device)
        return alpha, aabb[0] * (1 - samples) + aabb[1] * samples

    @torch.no_grad()
    def updateAlphaMask(self, gridSize=(200, 200, 200)):
        """tensorBase.py:386-409: rebuilds self.alphaMask from the current density field and returns the tight new_aabb [2,3]."""
        from .tensorf import AlphaGridMask
        lib = L.load()
        g = [int(x) for x in gridSize]
        alpha = self._dense_alpha_zyx(g)
        n_vox = g[0] * g[1] * g[2]
        volume = torch.empty((1, 1, g[2], g[1], g[0]), dtype=torch.float32, device=self.device)
        bits = torch.zeros((n_vox + 31) // 32 + 8, dtype=torch.int32, device=self.device)
        stats = torch.zeros(8, dtype=torch.int32, device=self.device)          # bbox idx [6] + n_set (uint64)
        L.check(lib.tvm_alpha_mask_from_dense(_ptr(alpha), _i3(g), float(self.alphaMask_thres), _ptr(volume), _ptr(bits),
                                              _ptr(stats), C.c_void_p(stats.data_ptr() + 24), _stream_ptr()),
                "tvm_alpha_mask_from_dense")
        self.alphaMask = AlphaGridMask(self.device, self.aabb, volume, packed_bits=bits)
        self._model_struct = None
        st = stats.cpu().numpy()
        lo, hi = st[:3], st[3:6]
        self.alpha_rest = int(st[6:8].view(np.uint64)[0]) / float(n_vox)       # the reference prints this ratio (:407)
        if hi[0] < 0:
            raise RuntimeError("updateAlphaMask: no voxel passed alphaMask_thres")  # the reference fails on the empty min()
        a = self.aabb.numpy()
        new_aabb = np.stack([lattice_point(a, g, lo), lattice_point(a, g, hi)])
        return torch.from_numpy(new_aabb)

    # ---- §8f-3 ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def filtering_mask(self, all_rays, N_samples=256, bbox_only=False):
        rays = all_rays.reshape(-1, all_rays.shape[-1])[:, :6].to(self.device, torch.float32).contiguous()
        mask = torch.empty(rays.shape[0], dtype=torch.uint8, device=self.device)
        model = self._model()
        fg_rand = None
        if not bbox_only and model.sampling == L.SAMPLING_NPP:
            # NerfPlusPlus.sample_ray perturbs its samples even with is_train=False (nerfplusplus.py:251)
            fg_rand = torch.rand((rays.shape[0], int(N_samples)), dtype=torch.float32, device=self.device)
        L.check(L.load().tvm_filter_rays(C.byref(model), _ptr(rays), rays.shape[0], int(N_samples), 1 if bbox_only else 0,
                                         _ptr(fg_rand), _ptr(mask), _stream_ptr()), "tvm_filter_rays")
        return mask.bool()

    @torch.no_grad()
    def filtering_rays(self, all_rays, all_rgbs, N_samples=256, chunk=10240 * 5, bbox_only=False):
        """tensorBase.py:411-441 -> (all_rays[mask], all_rgbs[mask]); `chunk` is accepted and ignored (one launch)."""
        mask = self.filtering_mask(all_rays, N_samples, bbox_only).view(all_rgbs.shape[:-1])
        m_r = mask.to(all_rays.device)
        return all_rays[m_r], all_rgbs[mask.to(all_rgbs.device)]

    # ---- §8f-4 ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _upsample_grids(self, pairs, res_target):
        """Bilinear (align_corners=True) resize of several (plane_coef, line_coef) halves in ONE launch (tvm_upsample_grids)."""
        from .tensorf import MAT_MODE, VEC_MODE
        lib, st = L.load(), _stream_ptr()
        srcs, dsts, chw, hw = [], [], [], []
        for plane_coef, line_coef in pairs:
            for i in range(3):
                m0, m1 = MAT_MODE[i]
                for src, (H2, W2) in ((plane_coef[i], (int(res_target[m1]), int(res_target[m0]))),
                                      (line_coef[i], (int(res_target[VEC_MODE[i]]), 1))):
                    s = src.detach().contiguous()
                    _, Cc, H, W = s.shape
                    srcs.append(s)
                    dsts.append(torch.empty((1, Cc, H2, W2), dtype=torch.float32, device=s.device))
                    chw += [Cc, H, W]
                    hw += [H2, W2]
        n = len(srcs)
        L.check(lib.tvm_upsample_grids(n, (C.c_void_p * n)(*[t.data_ptr() for t in srcs]), (C.c_int32 * (3 * n))(*chw),
                                       (C.c_void_p * n)(*[t.data_ptr() for t in dsts]), (C.c_int32 * (2 * n))(*hw), st),
                "tvm_upsample_grids")
        out = []
        for h in range(len(pairs)):
            g = dsts[6 * h:6 * h + 6]
            out.append((torch.nn.ParameterList([torch.nn.Parameter(g[2 * i]) for i in range(3)]),
                        torch.nn.ParameterList([torch.nn.Parameter(g[2 * i + 1]) for i in range(3)])))
        return out

    @torch.no_grad()
    def up_sampling_VM(self, plane_coef, line_coef, res_target):
        """tensoRF.py:248-262."""
        return self._upsample_grids([(plane_coef, line_coef)], res_target)[0]

    @torch.no_grad()
    def upsample_volume_grid(self, res_target):
        """tensoRF.py:264-269; all twelve grids in one launch."""
        (self.app_plane, self.app_line), (self.density_plane, self.density_line) = self._upsample_grids(
            [(self.app_plane, self.app_line), (self.density_plane, self.density_line)], res_target)
        self.update_stepSize(res_target)
        self._invalidate_packed()

    @torch.no_grad()
    def shrink(self, new_aabb):
        """tensoRF.py:271-314: crop every grid to the voxel box of new_aabb (index plumbing only, no arithmetic on the grids)."""
        from .tensorf import MAT_MODE, VEC_MODE
        f32 = np.float32
        new_aabb = np.asarray(new_aabb.detach().cpu() if torch.is_tensor(new_aabb) else new_aabb, dtype=f32).reshape(2, 3)
        aabb = self.aabb.numpy().astype(f32)
        units = self.units.numpy().astype(f32)
        G = self.gridSize.numpy().astype(np.int64)
        t_l = np.round(np.round((new_aabb[0] - aabb[0]) / units)).astype(np.int64)      # np.round == jt.round: half to even
        b_r = np.minimum(np.round((new_aabb[1] - aabb[0]) / units).astype(np.int64) + 1, G)
        crop = lambda p, *sl: torch.nn.Parameter(p.detach()[(Ellipsis, *sl)].contiguous())
        for i in range(3):
            v = VEC_MODE[i]
            self.density_line[i] = crop(self.density_line[i], slice(int(t_l[v]), int(b_r[v])), slice(None))
            self.app_line[i] = crop(self.app_line[i], slice(int(t_l[v]), int(b_r[v])), slice(None))
            m0, m1 = MAT_MODE[i]
            self.density_plane[i] = crop(self.density_plane[i], slice(int(t_l[m1]), int(b_r[m1])), slice(int(t_l[m0]), int(b_r[m0])))
            self.app_plane[i] = crop(self.app_plane[i], slice(int(t_l[m1]), int(b_r[m1])), slice(int(t_l[m0]), int(b_r[m0])))
        if not np.array_equal(self.alphaMask.gridSize.numpy().astype(np.int64), G):
            t_l_r, b_r_r = t_l.astype(f32) / (G - 1).astype(f32), (b_r - 1).astype(f32) / (G - 1).astype(f32)
            correct = np.zeros_like(new_aabb)
            correct[0] = (f32(1) - t_l_r) * aabb[0] + t_l_r * aabb[1]
            correct[1] = (f32(1) - b_r_r) * aabb[0] + b_r_r * aabb[1]
            new_aabb = correct.astype(f32)
        newSize = b_r - t_l
        self.aabb = torch.from_numpy(new_aabb.copy())
        self.update_stepSize((int(newSize[0]), int(newSize[1]), int(newSize[2])))
        self._invalidate_packed()


def get_rays_frame(c2w, H, W, focal, center=None, blender=False, normalize=True, device="cuda:0"):
    """get_ray_directions[_blender] + (normalise, blender.py:75) + get_rays (ray_utils.py:81-153) for one camera, on the
    device: returns all_rays [H*W, 6] = (origin, direction) without the 15 MB/frame host upload."""
    fx, fy = (focal, focal) if np.isscalar(focal) else (focal[0], focal[1])
    cx, cy = (W / 2, H / 2) if center is None else center
    dev = torch.device(device)
    rays = torch.empty((H * W, 6), dtype=torch.float32, device=dev)
    m = (C.c_float * 12)(*np.asarray(c2w, dtype=np.float32).reshape(-1)[:12].tolist())
    with torch.cuda.device(dev):
        L.check(L.load().tvm_generate_rays(m, int(H), int(W), float(fx), float(fy), float(cx), float(cy), 1 if blender else 0,
                                           1 if normalize else 0, _ptr(rays), _stream_ptr()), "tvm_generate_rays")
    return rays
