"""Host-side (numpy) restatement of the packed layouts of include/tvmrender.h, used to drive
tests/host_emul (the kernels' per-sample math compiled for the CPU).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def build_emul():
    src = os.path.join(HERE, "host_emul", "emul.cu")
    out_dir = os.path.join(HERE, "host_emul", "_build")
    out = os.path.join(out_dir, "libtvm_emul.so")
    deps = [src, os.path.join(ROOT, "jittor-myc-nerfs_b200", "csrc", "tvm_math.cuh"),
            os.path.join(ROOT, "include", "tvmrender.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler",
                               "-fPIC,-ffp-contract=off", "-shared", "-o", out, src])
    return C.CDLL(out)


def pack_bits(volume):
    flat = (np.asarray(volume).reshape(-1) > 0).astype(np.uint8)
    pad = (-flat.size) % 32
    flat = np.concatenate([flat, np.zeros(pad + 256, np.uint8)])
    return np.packbits(flat, bitorder="little").view(np.uint32).copy()


def pack_dilated(volume):
    """numpy restatement of tvm_pack_alpha_dilated: bit (z,y,x) = OR of the 2x2x2 corner bits, clipped at the upper faces."""
    v = np.asarray(volume).reshape(np.asarray(volume).shape[-3:]) > 0
    pad = np.zeros(tuple(d + 1 for d in v.shape), bool)
    pad[:-1, :-1, :-1] = v
    d = np.zeros_like(v)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                d |= pad[dz:dz + v.shape[0], dy:dy + v.shape[1], dx:dx + v.shape[2]]
    return pack_bits(d)


def host_model(pkg, params, alpha_volume=None, alpha_aabb=None, dilated=True):
    """TvmModel whose pointers are numpy host buffers (kept alive in the returned list)."""
    L = pkg._lib
    s = pkg.derive_march_scalars(params.aabb, params.gridSize, params.step_ratio)
    m = L.TvmModel()
    keep = []

    def buf(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        keep.append(a)
        return a.ctypes.data

    for i in range(3):
        m.aabb[i] = float(params.aabb[0, i])
        m.aabb[3 + i] = float(params.aabb[1, i])
        m.inv_aabb_size[i] = float(s["invaabbSize"][i])
        m.grid[i] = int(params.gridSize[i])
    m.step_size = float(s["stepSize"])
    m.near_, m.far_ = params.near_far
    m.density_shift, m.distance_scale = params.density_shift, params.distance_scale
    m.weight_thres = params.rayMarch_weight_thres
    m.act = 0
    m.n_density, m.n_app, m.app_dim = params.density_n_comp[0], params.app_n_comp[0], params.app_dim
    m.view_pe, m.fea_pe, m.feature_c = params.view_pe, params.fea_pe, params.featureC
    for k in range(3):
        m.density_plane[k] = buf(params.density_plane[k][0].transpose(1, 2, 0))   # [H][W][C]
        m.density_line[k] = buf(params.density_line[k][0, :, :, 0].T)             # [L][C]
        m.app_plane[k] = buf(params.app_plane[k][0].transpose(1, 2, 0))
        m.app_line[k] = buf(params.app_line[k][0, :, :, 0].T)
    bt = np.zeros((params.basis_mat.shape[1], 32), np.float32)
    bt[:, :params.app_dim] = params.basis_mat.T
    m.basis_t = buf(bt)
    m.w1_t = buf(params.mlp_w[0].T)
    m.b1 = buf(params.mlp_b[0])
    m.w2_t = buf(params.mlp_w[1].T)
    m.b2 = buf(params.mlp_b[1])
    m.w3 = buf(params.mlp_w[2])
    m.b3 = buf(params.mlp_b[2])
    if alpha_volume is not None:
        bits = pack_bits(alpha_volume)
        keep.append(bits)
        m.alpha_bits = bits.ctypes.data
        if dilated:                                  # the one-lookup fast path of alpha_mask_test
            dil = pack_dilated(alpha_volume)
            keep.append(dil)
            m.alpha_dilated = dil.ctypes.data
        D, H, W = alpha_volume.shape[-3:]
        a = np.asarray(alpha_aabb, np.float32).reshape(2, 3)
        inv = (np.float32(1.0) / (a[1] - a[0]) * np.float32(2)).astype(np.float32)
        for i, g in enumerate((W, H, D)):
            m.alpha_grid[i] = g
            m.alpha_aabb_min[i] = float(a[0, i])
            m.alpha_inv_size[i] = float(inv[i])
    return m, keep, s


def emul_forward(pkg, case, S=-1, white_bg=True):
    lib = build_emul()
    m, keep, s = host_model(pkg, case["model"], case["alpha_volume"], case["alpha_aabb"])
    S = s["nSamples"] if S <= 0 else S
    rays = np.ascontiguousarray(case["rays"], np.float32)
    n = rays.shape[0]
    jit = case.get("jitter")
    if case.get("fg_rand") is not None:            # NerfPlusPlus sampling: `jitter` is the [n][S] draw array
        m.sampling, m.radii = pkg._lib.SAMPLING_NPP, float(case["model"].extra["radii"])
        jit = np.ascontiguousarray(case["fg_rand"], np.float32)
        assert jit.shape == (n, S)
    u8 = lambda: np.zeros((n, S), np.uint8)
    f = lambda *sh: np.zeros(sh, np.float32)
    out = dict(bbox=u8(), valid=u8(), app=u8(), sigma=f(n, S), weight=f(n, S), rgb=f(n, S, 3),
               rgb_map=f(n, 3), depth_map=f(n))
    p = lambda a: C.c_void_p(a.ctypes.data)
    lib.emul_forward.argtypes = [C.c_void_p] * 2 + [C.c_int] * 2 + [C.c_void_p, C.c_uint32] + [C.c_void_p] * 8
    rc = lib.emul_forward(C.byref(m), p(rays), n, S, p(jit) if jit is not None else None,
                          pkg._lib.WHITE_BG if white_bg else 0, *(p(out[k]) for k in
                          ("bbox", "valid", "app", "sigma", "weight", "rgb", "rgb_map", "depth_map")))
    assert rc == 0
    out["S"] = S
    return out


def pack_bricks(volume):
    """numpy restatement of tvm_pack_alpha_bricks: one bit per 8x8x8 brick, set iff any voxel is set."""
    v = np.asarray(volume) > 0
    D, H, W = v.shape
    BD, BH, BW = (D + 7) // 8, (H + 7) // 8, (W + 7) // 8
    pad = np.zeros((BD * 8, BH * 8, BW * 8), bool)
    pad[:D, :H, :W] = v
    b = pad.reshape(BD, 8, BH, 8, BW, 8).any(axis=(1, 3, 5)).reshape(-1).astype(np.uint8)
    b = np.concatenate([b, np.zeros((-b.size) % 32 + 256, np.uint8)])
    return np.packbits(b, bitorder="little").view(np.uint32).copy()


def pack_bricks3(volume):
    """numpy restatement of tvm_pack_alpha_bricks3: one uint32 per brick, bit (dz+1)*9 + (dy+1)*3 + (dx+1) = the neighbour
    brick at that offset holds a set voxel (outside the grid: 0)."""
    v = np.asarray(volume) > 0
    D, H, W = v.shape
    BD, BH, BW = (D + 7) // 8, (H + 7) // 8, (W + 7) // 8
    pad = np.zeros((BD * 8, BH * 8, BW * 8), bool)
    pad[:D, :H, :W] = v
    b = pad.reshape(BD, 8, BH, 8, BW, 8).any(axis=(1, 3, 5))
    p = np.pad(b, 1)
    d = np.zeros(b.shape, np.uint32)
    for dz in range(3):
        for dy in range(3):
            for dx in range(3):
                d |= p[dz:dz + BD, dy:dy + BH, dx:dx + BW].astype(np.uint32) << np.uint32(dz * 9 + dy * 3 + dx)
    return np.concatenate([d.reshape(-1), np.zeros(8, np.uint32)]).copy()


def emul_block_maybe(pkg, case, S=-1, bricks3=True):
    lib = build_emul()
    m, keep, s = host_model(pkg, case["model"], case["alpha_volume"], case["alpha_aabb"])
    if case["alpha_volume"] is not None:
        bricks = pack_bricks(case["alpha_volume"])
        keep.append(bricks)
        m.alpha_bricks = bricks.ctypes.data
        if bricks3:                                    # the one-lookup rejection in front of the exact brick loop
            b3 = pack_bricks3(case["alpha_volume"])
            keep.append(b3)
            m.alpha_bricks3 = b3.ctypes.data
    S = s["nSamples"] if S <= 0 else S
    rays = np.ascontiguousarray(case["rays"], np.float32)
    n, NB = rays.shape[0], (S + 31) // 32
    jit = case.get("jitter")
    visit = np.zeros((n, NB), np.uint8)
    lib.emul_block_maybe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    rc = lib.emul_block_maybe(C.byref(m), C.c_void_p(rays.ctypes.data), n, S,
                              C.c_void_p(jit.ctypes.data) if jit is not None else None, C.c_void_p(visit.ctypes.data))
    assert rc == 0
    return visit.astype(bool), S
