"""Pins oracle/maintain_oracle.py (updateAlphaMask, filtering_rays, ray generation, regularisers,
upsample / shrink) against tests/golden/maint_*.npz, recorded from the reference's UNMODIFIED python over
oracle/jt_shim (tests/golden/make_golden_maint.py).  Reads only the committed .npz files."""
import ast
import os

import numpy as np
import torch

from oracle import fixtures as fx, tensorf_oracle as orc, maintain_oracle as mo

GD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def alpha_case(name):
    g = np.load(os.path.join(GD, name + ".npz"))
    G, mask_res, grid, shift, scale = [ast.literal_eval(str(x)) for x in g["args"]]
    case = fx.make_case(G, 8, "R1" if mask_res else "R0", mask_res=mask_res, grid_scale=scale)
    case["model"].density_shift = shift
    return g, case, grid


def test_update_alpha_mask_matches_reference():
    for name in ("maint_alpha_masked", "maint_alpha_nomask"):
        g, case, grid = alpha_case(name)
        m = orc.make_oracle(case)
        alpha, _ = mo.get_dense_alpha(m, grid)
        assert np.allclose(alpha.numpy(), g["alpha"], rtol=2e-5, atol=1.5e-7)      # 1 - exp(-x): one ulp of 1.0
        r = mo.update_alpha_mask(m, grid, thres=0.001)
        # a voxel may only differ where the pooled alpha sits within float noise of the threshold
        diff = r["volume"] != g["volume"]
        assert np.all(np.abs(r["pooled"][diff] - 0.001) < 2e-7)
        assert diff.sum() <= 2
        if diff.sum() == 0:
            assert np.array_equal(r["new_aabb"], g["new_aabb"])


def test_filtering_rays_matches_reference():
    g = np.load(os.path.join(GD, "maint_filter.npz"))
    G, mask_res, n, S = [ast.literal_eval(str(x)) for x in g["args"]]
    case = fx.make_case(G, n, "R1", mask_res=mask_res)
    m = orc.make_oracle(case)
    assert np.array_equal(mo.filtering_rays_mask(m, g["rays"], bbox_only=True), g["mask_bbox"])
    assert np.array_equal(mo.filtering_rays_mask(m, g["rays"], N_samples=S, bbox_only=False), g["mask_alpha"])


def test_ray_generation_matches_reference():
    g = np.load(os.path.join(GD, "maint_rays.npz"))
    H, W = [int(x) for x in g["args"]]
    f = float(g["focal"])
    d = mo.get_ray_directions(H, W, [f, f])
    d = d / torch.norm(d, dim=-1, keepdim=True)
    o, r = mo.get_rays(d, g["c2w"])
    assert np.array_equal(o.numpy(), g["rays_o"]) and np.abs(r.numpy() - g["rays_d"]).max() <= 1e-7
    db = mo.get_ray_directions(H, W, [f, f * 1.1], center=[W / 2 - 0.25, H / 2 + 1.5], blender=True)
    ob, rb = mo.get_rays(db, g["c2w"])
    assert np.abs(rb.numpy() - g["rays_d_blender"]).max() <= 1e-6
    # fixtures.frame_rays (the bench input) is the same construction
    fr = mo.frame_rays(H, W, [f, f], g["c2w"])
    assert np.abs(fr[:, 3:] - g["rays_d"]).max() <= 1e-7


def reg_model():
    g = np.load(os.path.join(GD, "maint_reg.npz"))
    G = ast.literal_eval(str(g["args"][0]))
    p = fx.make_case(G, 8, "R0")["model"]
    t = lambda a: torch.tensor(a, requires_grad=True)
    return g, dict(density_plane=[t(a) for a in p.density_plane], density_line=[t(a) for a in p.density_line],
                   app_plane=[t(a) for a in p.app_plane], app_line=[t(a) for a in p.app_line])


def test_regularisers_match_reference():
    g, P = reg_model()
    fns = {"tv_density": lambda: mo.tv_loss_planes(P["density_plane"]), "tv_app": lambda: mo.tv_loss_planes(P["app_plane"]),
           "l1": lambda: mo.density_l1(P["density_plane"], P["density_line"]),
           "ortho": lambda: mo.vector_diffs(P["density_line"]) + mo.vector_diffs(P["app_line"])}
    for key, fn in fns.items():
        for lst in P.values():
            for x in lst:
                x.grad = None
        loss = fn()
        loss.backward()
        assert abs(float(loss) - float(g[key])) <= 1e-6 * max(1.0, abs(float(g[key])))
        for name, lst in P.items():
            for k in range(3):
                gk = f"{key}.{name}.{k}"
                if gk in g.files:
                    assert np.allclose(lst[k].grad.numpy(), g[gk], rtol=1e-5, atol=1e-9), gk
                else:
                    assert lst[k].grad is None or float(lst[k].grad.abs().max()) == 0.0


def test_upsample_and_shrink_match_reference():
    g = np.load(os.path.join(GD, "maint_resize.npz"))
    G, target = [ast.literal_eval(str(x)) for x in g["args"]]
    p = fx.make_case(G, 8, "R1", mask_res=16)["model"]
    t = lambda lst: [torch.tensor(a) for a in lst]
    dp, dl = mo.up_sampling_vm(t(p.density_plane), t(p.density_line), target)
    ap, al = mo.up_sampling_vm(t(p.app_plane), t(p.app_line), target)
    for name, lst in (("density_plane", dp), ("density_line", dl), ("app_plane", ap), ("app_line", al)):
        for k in range(3):
            assert np.allclose(lst[k].numpy(), g[f"up.{name}.{k}"], rtol=1e-6, atol=1e-7)
    from jittor_myc_nerfs_b200 import derive_march_scalars
    s = derive_march_scalars(p.aabb, target, p.step_ratio)
    assert np.float32(s["stepSize"]) == g["up.stepSize"] and s["nSamples"] == int(g["up.nSamples"])
    out, new_aabb, new_size = mo.shrink(p.aabb, s["units"], target, g["new_aabb"], dp, dl, ap, al, mask_grid_equal=False)
    for name, key in (("density_plane", "dp"), ("density_line", "dl"), ("app_plane", "ap"), ("app_line", "al")):
        for k in range(3):
            assert np.array_equal(out[key][k].numpy(), g[f"shrink.{name}.{k}"])
    assert np.allclose(new_aabb, g["shrink.aabb"], atol=1e-6)
    assert new_size.tolist() == g["shrink.gridSize"].tolist()
    s2 = derive_march_scalars(new_aabb, new_size, p.step_ratio)
    assert abs(float(s2["stepSize"]) - float(g["shrink.stepSize"])) <= 1e-7 and s2["nSamples"] == int(g["shrink.nSamples"])


def test_adam_step_formula():
    """A11 against an independent float64 evaluation of the same recurrences (Jittor itself cannot run here)."""
    rng = np.random.default_rng(0)
    p0, g1, g2 = (rng.standard_normal(1000).astype(np.float32) for _ in range(3))
    p, m, v = torch.tensor(p0), torch.zeros(1000), torch.zeros(1000)
    mo.adam_step(p, torch.tensor(g1), m, v, lr=0.02, n=1)
    mo.adam_step(p, torch.tensor(g2), m, v, lr=0.02, n=2)
    P, M, V = p0.astype(np.float64), np.zeros(1000), np.zeros(1000)
    for n, gg in ((1, g1), (2, g2)):
        M = 0.9 * M + 0.1 * gg
        V = 0.99 * V + 0.01 * gg.astype(np.float64) ** 2
        P = P - M * (0.02 * np.sqrt(1 - 0.99 ** n) / (1 - 0.9 ** n)) / (np.sqrt(V) + 1e-8)
    assert np.abs(p.numpy() - P).max() <= 1e-5
