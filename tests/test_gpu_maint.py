"""GPU parity of the callers either side of the ray path: libtvmrender's kernels, called through the
reference-named host methods, against oracle/maintain_oracle.py and the committed golden vectors."""
import ast
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def env(built_lib):
    import torch
    from oracle import fixtures as fx, tensorf_oracle as orc, maintain_oracle as mo
    built_lib._lib.require_cuda()
    return built_lib, torch, fx, orc, mo


def test_update_alpha_mask(env):
    """getDenseAlpha + updateAlphaMask (tensorBase.py:366-409) vs the golden vectors of the unmodified reference."""
    pkg, torch, fx, orc, mo = env
    from util import gpu_model
    for name in ("maint_alpha_masked", "maint_alpha_nomask"):
        g = np.load(os.path.join(GD, name + ".npz"))
        G, mask_res, grid, shift, scale = [ast.literal_eval(str(x)) for x in g["args"]]
        case = fx.make_case(G, 8, "R1" if mask_res else "R0", mask_res=mask_res, grid_scale=scale)
        case["model"].density_shift = shift
        model = gpu_model(pkg, case)
        alpha, dense_xyz = model.getDenseAlpha(list(grid))
        assert tuple(alpha.shape) == tuple(grid) and tuple(dense_xyz.shape) == (*grid, 3)
        assert np.allclose(alpha.cpu().numpy(), g["alpha"], rtol=2e-5, atol=1.5e-7)
        new_aabb = model.updateAlphaMask(tuple(grid))
        vol = model.alphaMask.alpha_volume.cpu().numpy().reshape(grid[::-1])
        ref = mo.update_alpha_mask(orc.make_oracle(case), grid, thres=0.001)
        diff = vol != g["volume"]
        assert diff.sum() <= 2 and np.all(np.abs(ref["pooled"][diff] - 0.001) < 2e-7)
        if diff.sum() == 0:
            assert np.array_equal(new_aabb.numpy(), g["new_aabb"])
        # the bit stream and the brick index the ray path consumes agree with the float volume
        bits = model.alphaMask.bits.cpu().numpy().view(np.uint32)
        unpacked = np.unpackbits(bits.view(np.uint8), bitorder="little")[: vol.size].reshape(vol.shape)
        assert np.array_equal(unpacked.astype(bool), vol > 0.5)
        assert abs(model.alpha_rest - vol.mean()) < 1e-9
        # and the new mask drives the renderer: masks bit-exact against the oracle holding the same volume
        case2 = dict(case, alpha_volume=vol, alpha_aabb=case["model"].aabb.copy(), rays=fx.subset_rays(64))
        out = model.forward_with_aux(torch.from_numpy(case2["rays"]).cuda())
        r = orc.run_case(case2)
        assert np.array_equal(pkg.unpack_bits(out["valid_bits"], model.nSamples), r["ray_valid"])


def test_filtering_rays(env):
    pkg, torch, fx, orc, mo = env
    from util import gpu_model
    g = np.load(os.path.join(GD, "maint_filter.npz"))
    G, mask_res, n, S = [ast.literal_eval(str(x)) for x in g["args"]]
    case = fx.make_case(G, n, "R1", mask_res=mask_res)
    model = gpu_model(pkg, case)
    rays = torch.from_numpy(g["rays"]).cuda()
    rgbs = torch.arange(n * 3, dtype=torch.float32).view(n, 3).cuda()
    assert np.array_equal(model.filtering_mask(rays, bbox_only=True).cpu().numpy(), g["mask_bbox"])        # bit-exact
    assert np.array_equal(model.filtering_mask(rays, N_samples=S).cpu().numpy(), g["mask_alpha"])          # bit-exact
    r2, c2 = model.filtering_rays(rays, rgbs, N_samples=S)
    assert r2.shape[0] == int(g["mask_alpha"].sum()) and np.array_equal(r2.cpu().numpy(), g["rays"][g["mask_alpha"]])
    assert np.array_equal(c2.cpu().numpy(), rgbs.cpu().numpy()[g["mask_alpha"]])
    # empty-space index off: same decisions
    model.empty_space_skipping = False
    assert np.array_equal(model.filtering_mask(rays, N_samples=S).cpu().numpy(), g["mask_alpha"])
    # full-size property: a frame of rays through the 300^3 / 200^3 configuration against the oracle on a slice
    case = fx.make_case(300, 0, "R1", full_frame=True)
    model = gpu_model(pkg, case)
    rays = torch.from_numpy(case["rays"]).cuda()
    m_gpu = model.filtering_mask(rays, N_samples=256).cpu().numpy()
    sl = slice(320000, 320000 + 2048)
    assert np.array_equal(m_gpu[sl], mo.filtering_rays_mask(orc.make_oracle(case), case["rays"][sl], N_samples=256))
    assert 0.3 < m_gpu.mean() < 0.9


def test_ray_generation(env):
    pkg, torch, fx, orc, mo = env
    g = np.load(os.path.join(GD, "maint_rays.npz"))
    H, W = [int(x) for x in g["args"]]
    f = float(g["focal"])
    rays = pkg.get_rays_frame(g["c2w"], H, W, f).cpu().numpy()
    assert np.array_equal(rays[:, :3], g["rays_o"]) and np.abs(rays[:, 3:] - g["rays_d"]).max() <= 2e-7
    rb = pkg.get_rays_frame(g["c2w"], H, W, [f, f * 1.1], center=[W / 2 - 0.25, H / 2 + 1.5], blender=True,
                            normalize=False).cpu().numpy()
    assert np.abs(rb[:, 3:] - g["rays_d_blender"]).max() <= 1e-6
    # the bench's 800x800 frame, generated on the device, equals the host fixture
    full = pkg.get_rays_frame(fx.camera_pose(0.7, 0.5), 800, 800, 0.5 * 800 / np.tan(0.5 * 0.6911)).cpu().numpy()
    assert np.abs(full - fx.frame_rays()).max() <= 2e-6


def test_regularisers(env):
    pkg, torch, fx, orc, mo = env
    from util import gpu_model
    g = np.load(os.path.join(GD, "maint_reg.npz"))
    G = ast.literal_eval(str(g["args"][0]))
    model = gpu_model(pkg, fx.make_case(G, 8, "R0"))
    reg = pkg.TVLoss()
    fns = {"tv_density": lambda: model.TV_loss_density(reg), "tv_app": lambda: model.TV_loss_app(reg),
           "l1": model.density_L1, "ortho": model.vector_comp_diffs}
    groups = dict(density_plane=model.density_plane, density_line=model.density_line, app_plane=model.app_plane,
                  app_line=model.app_line)
    for key, fn in fns.items():
        for p in model.parameters():
            p.grad = None
        loss = fn() * 3.0                                  # the incoming gradient scale must reach the grids
        loss.backward()
        assert abs(float(loss.detach()) / 3.0 - float(g[key])) <= 2e-6 * max(1.0, abs(float(g[key])))
        for name, lst in groups.items():
            for k in range(3):
                gk = f"{key}.{name}.{k}"
                if gk in g.files:
                    assert np.allclose(lst[k].grad.cpu().numpy() / 3.0, g[gk], rtol=2e-5, atol=1e-9), gk
                else:
                    assert lst[k].grad is None
    # a lone plane through TVLoss itself (utils.py:128-139)
    x = model.app_plane[1].detach().clone().requires_grad_(True)
    l = reg(x)
    l.backward()
    xr = x.detach().cpu().clone().requires_grad_(True)
    lr = mo.tv_loss(xr)
    lr.backward()
    assert abs(float(l.detach()) - float(lr.detach())) <= 1e-6 * float(lr.detach()) + 1e-9
    assert np.allclose(x.grad.cpu().numpy(), xr.grad.numpy(), rtol=2e-5, atol=1e-10)


def test_adam_matches_oracle(env):
    pkg, torch, fx, orc, mo = env
    rng = np.random.default_rng(3)
    shapes = [(1, 16, 40, 33), (1, 16, 40, 1), (27, 144), (128,), (5,), (1, 48, 7, 9)]
    params = [torch.nn.Parameter(torch.from_numpy(rng.standard_normal(s).astype(np.float32)).cuda()) for s in shapes]
    ref = [p.detach().cpu().clone() for p in params]
    rm, rv = [torch.zeros_like(r) for r in ref], [torch.zeros_like(r) for r in ref]
    groups = [{"params": params[:2], "lr": 0.02}, {"params": params[2:], "lr": 0.001}]
    opt = pkg.Adam(groups, lr=0.001, betas=(0.9, 0.99))
    for step in range(1, 4):
        gs = [torch.from_numpy(rng.standard_normal(s).astype(np.float32)) for s in shapes]
        for p, gg in zip(params, gs):
            p.grad = gg.cuda()
        v0 = [p._version for p in params]
        opt.step()
        assert all(p._version > v for p, v in zip(params, v0))          # the packed device image will be rebuilt
        for i, (r, gg) in enumerate(zip(ref, gs)):
            mo.adam_step(r, gg, rm[i], rv[i], lr=0.02 if i < 2 else 0.001, n=step)
        for p, r in zip(params, ref):
            assert np.allclose(p.detach().cpu().numpy(), r.numpy(), rtol=1e-5, atol=2e-6)
    opt.param_groups[0]["lr"] = 0.5                                      # train.py:263-264 rescales lr in place
    for p in params:
        p.grad = torch.ones_like(p)
    before = params[0].detach().clone()
    opt.step()
    assert float((params[0].detach() - before).abs().max()) > 0.01


def test_upsample_and_shrink(env):
    pkg, torch, fx, orc, mo = env
    from util import gpu_model
    g = np.load(os.path.join(GD, "maint_resize.npz"))
    G, target = [ast.literal_eval(str(x)) for x in g["args"]]
    case = fx.make_case(G, 64, "R1", mask_res=16)
    model = gpu_model(pkg, case)
    model.upsample_volume_grid(list(target))
    groups = dict(density_plane=model.density_plane, density_line=model.density_line, app_plane=model.app_plane,
                  app_line=model.app_line)
    for name, lst in groups.items():
        for k in range(3):
            assert np.allclose(lst[k].detach().cpu().numpy(), g[f"up.{name}.{k}"], rtol=1e-6, atol=1e-7), name
    assert np.float32(model.stepSize) == g["up.stepSize"] and model.nSamples == int(g["up.nSamples"])
    model.shrink(torch.from_numpy(g["new_aabb"]))
    groups = dict(density_plane=model.density_plane, density_line=model.density_line, app_plane=model.app_plane,
                  app_line=model.app_line)
    for name, lst in groups.items():
        for k in range(3):
            assert np.allclose(lst[k].detach().cpu().numpy(), g[f"shrink.{name}.{k}"], rtol=1e-6, atol=1e-7), name
    assert np.allclose(model.aabb.numpy(), g["shrink.aabb"], atol=1e-6)
    assert model.gridSize.tolist() == g["shrink.gridSize"].tolist()
    assert abs(float(model.stepSize) - float(g["shrink.stepSize"])) <= 1e-7 and model.nSamples == int(g["shrink.nSamples"])
    # the resized model renders: compare with the oracle built from the resized parameters
    import dataclasses
    host = lambda lst: [t.detach().cpu().numpy() for t in lst]
    p2 = dataclasses.replace(case["model"], gridSize=tuple(int(x) for x in model.gridSize), aabb=model.aabb.numpy().copy(),
                             density_plane=host(model.density_plane), density_line=host(model.density_line),
                             app_plane=host(model.app_plane), app_line=host(model.app_line))
    case2 = dict(case, model=p2, alpha_aabb=case["model"].aabb.copy())
    r = orc.run_case(case2)
    with torch.no_grad():
        rgb, depth = model(torch.from_numpy(case["rays"]).cuda(), white_bg=True, is_train=False)
    assert r["nSamples"] == model.nSamples
    assert np.abs(rgb.cpu().numpy() - r["rgb_map"]).max() <= 1e-4
