"""GPU parity of tvm_forward (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): valid-sample indices and alpha-mask decisions bit-exact;
rgb/depth within 1e-4 abs with the fp32 appearance head; PSNR delta < 0.01 dB."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RGB_TOL = 1e-4      # north_star: rgb within 1e-4 abs in fp32
DEPTH_TOL = 1e-4    # north_star: depth within 1e-4 abs


@pytest.fixture(scope="module")
def env(built_lib):
    import torch
    from oracle import fixtures as fx, tensorf_oracle as orc
    built_lib._lib.require_cuda()
    return built_lib, torch, fx, orc


def _check_aux(pkg, torch, orc, case, white_bg=True, S=-1):
    from util import gpu_model
    model = gpu_model(pkg, case)
    rays = torch.from_numpy(case["rays"]).cuda()
    jit = None if case["jitter"] is None else torch.from_numpy(case["jitter"]).cuda()
    ref = orc.run_case(case, N_samples=S, white_bg=white_bg)
    Sx = ref["nSamples"] if S <= 0 else S
    out = model.forward_with_aux(rays, white_bg=white_bg, N_samples=S, jitter=jit)
    torch.cuda.synchronize()
    bbox = pkg.unpack_bits(out["bbox_bits"], Sx)
    valid = pkg.unpack_bits(out["valid_bits"], Sx)
    app = pkg.unpack_bits(out["app_bits"], Sx)
    assert np.array_equal(bbox, ref["bbox_valid"]), "in-bbox mask differs"
    assert np.array_equal(valid, ref["ray_valid"]), "ray_valid (alpha-mask decisions) differs"
    n_app_mis = int((app != ref["app_mask"]).sum())
    # app_mask is a float threshold (weight > 1e-4), not in the bit-exact set: allow ulp-level flips
    assert n_app_mis <= max(2, int(2e-4 * max(1, ref["app_mask"].sum()))), f"{n_app_mis} app_mask mismatches"
    sig = out["sigma"].cpu().numpy()
    # log(1+exp(x)) (A3) quantises sigma to 2^-23 steps near x ~ -10: one expf ulp can move it a step
    assert np.allclose(sig, ref["sigma"], rtol=2e-5, atol=2.5e-7)
    assert np.abs(out["weight"].cpu().numpy() - ref["weight"]).max() <= 2e-6
    both = app & ref["app_mask"]
    rgb_s = out["rgb"].cpu().numpy()
    assert np.abs(rgb_s[both] - ref["rgb"][both]).max(initial=0) <= 2e-5
    assert np.abs(out["rgb_map"].cpu().numpy() - ref["rgb_map"]).max() <= RGB_TOL
    d_err = np.abs(out["depth_map"].cpu().numpy() - ref["depth_map"]).max()
    assert d_err <= DEPTH_TOL, d_err
    return model, rays, jit, ref, out


@pytest.mark.parametrize("regime", ["R0", "R1", "R2"])
def test_forward_aux_parity_128(env, regime):
    pkg, torch, fx, orc = env
    case = fx.make_case(128, 2048, regime)
    _check_aux(pkg, torch, orc, case)


def test_forward_train_jitter_and_black_bg(env):
    pkg, torch, fx, orc = env
    case = fx.make_case(128, 1024, "R2", train=True)
    _check_aux(pkg, torch, orc, case, white_bg=False, S=443)   # cal_n_samples value used in training


def test_forward_ert_path_matches(env):
    """The production path (no aux, early ray termination on) against the oracle, and chunk invariance."""
    pkg, torch, fx, orc = env
    from util import gpu_model, psnr
    for regime in ("R1", "R2"):
        case = fx.make_case(128, 4096, regime)
        ref = orc.run_case(case, want_stages=False)
        model = gpu_model(pkg, case)
        rays = torch.from_numpy(case["rays"]).cuda()
        with torch.no_grad():
            rgb, _, depth, _, _ = pkg.OctreeRender_trilinear_fast(rays, model, chunk=4096, N_samples=-1,
                                                                  white_bg=True, is_train=False)
            model.early_termination = False
            rgb2, depth2 = model(rays)
            model.early_termination = True
            rgb3 = torch.cat([model(rays[s:s + 1000])[0] for s in range(0, 4096, 1000)])
        rgb, depth, rgb2 = rgb.cpu().numpy(), depth.cpu().numpy(), rgb2.cpu().numpy()
        assert np.abs(rgb - ref["rgb_map"]).max() <= RGB_TOL
        assert np.abs(rgb2 - ref["rgb_map"]).max() <= RGB_TOL
        assert np.abs(depth - ref["depth_map"]).max() <= DEPTH_TOL
        assert np.array_equal(rgb3.cpu().numpy(), rgb), "result depends on the chunking"
        # PSNR delta < 0.01 dB against a random target image
        tgt = fx.target_rgb(4096)
        assert abs(psnr(rgb, tgt) - psnr(ref["rgb_map"], tgt)) < 0.01


def test_forward_300_grid_mask200(env):
    pkg, torch, fx, orc = env
    case = fx.make_case(300, 1024, "R1")
    _check_aux(pkg, torch, orc, case)


def test_edge_cases(env):
    """Ragged ray counts, rays that miss the box, axis-aligned rays (d component == 0 -> 1e-6 rule),
    a ray starting inside the box, S not a multiple of 32, non-cubic grid."""
    pkg, torch, fx, orc = env
    case = fx.make_case((64, 80, 100), 37, "R2", mask_res=(50, 60, 70))
    rays = case["rays"].copy()
    rays[0] = [0.3, 0.2, 12.0, 0, 0, -1]          # straight down the z axis: two zero components
    rays[1] = [12.0, 0.1, -0.2, -1, 0, 0]
    rays[2] = [0.0, 0.0, 0.0, 0.6, 0.8, 0.0]      # starts inside the box (t_min clamps to near)
    rays[3] = [20.0, 20.0, 20.0, 0.0, 0.0, 1.0]   # misses the box
    rays[4] = [5.0, 0.0, 12.0, 0, 0, -1]          # grazes the x = +5 face exactly (strict > keeps it inside)
    rays[5] = [-12.0, 5.0, 5.0, 1, 0, 0]          # runs along an edge
    case["rays"] = rays
    for S in (-1, 33, 95, 1):
        _check_aux(pkg, torch, orc, case, S=S)
    one = dict(case, rays=rays[:1].copy())
    _check_aux(pkg, torch, orc, one)


def test_no_cpu_fallback(env):
    pkg, torch, fx, orc = env
    from util import gpu_model
    case = fx.make_case(32, 8, "R0")
    model = gpu_model(pkg, case)
    with pytest.raises(Exception):
        model(torch.from_numpy(case["rays"]))     # host tensor: must not be silently computed on the CPU


def test_compute_alpha(env):
    pkg, torch, fx, orc = env
    from util import gpu_model
    case = fx.make_case(64, 8, "R2", mask_res=64)
    model = gpu_model(pkg, case)
    rng = np.random.default_rng(5)
    xyz = rng.uniform(-5, 5, (5000, 3)).astype(np.float32)
    o = orc.OracleTensorVMSplit(case["model"], case["alpha_volume"], case["alpha_aabb"])
    ref = o.compute_alpha(torch.from_numpy(xyz), float(o.stepSize)).numpy()
    got = model.compute_alpha(torch.from_numpy(xyz).cuda(), float(model.stepSize)).cpu().numpy()
    assert np.array_equal(got > 0, ref > 0)
    assert np.allclose(got, ref, rtol=2e-5, atol=2.5e-7)


def test_tensor_core_mlp_bf16(env):
    """TVM_MLP_BF16: tcgen05 appearance head; north_star tolerance 1e-2 on rgb, PSNR delta < 0.01 dB."""
    pkg, torch, fx, orc = env
    from util import gpu_model, psnr
    for regime, G in (("R1", 128), ("R2", 128), ("R1", 300)):
        case = fx.make_case(G, 2048, regime)
        ref = orc.run_case(case, want_stages=False)
        model = gpu_model(pkg, case, mlp_mode="bf16")
        rays = torch.from_numpy(case["rays"]).cuda()
        with torch.no_grad():
            rgb, depth = model(rays)
            model.mlp_mode = "fp32"
            rgb32, _ = model(rays)
        torch.cuda.synchronize()
        rgb, rgb32 = rgb.cpu().numpy(), rgb32.cpu().numpy()
        err = np.abs(rgb - ref["rgb_map"]).max()
        print(f"bf16 MLP {regime} G={G}: max|rgb-oracle|={err:.3e} max|rgb-fp32 path|={np.abs(rgb - rgb32).max():.3e}")
        assert err <= 1e-2
        assert np.abs(depth.cpu().numpy() - ref["depth_map"]).max() <= DEPTH_TOL
        tgt = fx.target_rgb(2048)
        assert abs(psnr(rgb, tgt) - psnr(ref["rgb_map"], tgt)) < 0.01
    # ragged tile (entries not a multiple of 128) and a single ray
    case = fx.make_case(64, 3, "R2", mask_res=64)
    ref = orc.run_case(case, want_stages=False)
    model = gpu_model(pkg, case, mlp_mode="bf16")
    with torch.no_grad():
        rgb, _ = model(torch.from_numpy(case["rays"]).cuda())
    assert np.abs(rgb.cpu().numpy() - ref["rgb_map"]).max() <= 1e-2


def test_bf16_appearance_planes(env):
    """app_planes_bf16 (TvmModel.app_plane_pair / app_line_pair): the tensor-core head gathers 16-bit pair records of the appearance grids.
    Same north_star bound as the bf16 head (1e-2 on rgb, PSNR delta < 0.01 dB); masks, depth and the work counters do
    not depend on it.  REFTensoRF uses the same gather."""
    pkg, torch, fx, orc = env
    from util import gpu_model, psnr
    for regime, G, variant in (("R1", 128, None), ("R2", 128, None), ("R1", 300, None), ("R2", 64, "ref")):
        kw = dict(variant=variant, mask_res=G) if variant else {}
        case = fx.make_case(G, 2048, regime, **kw)
        ref = orc.run_case(case, want_stages=False)
        model = gpu_model(pkg, case, mlp_mode="bf16")
        rays = torch.from_numpy(case["rays"]).cuda()
        model.collect_counters = True
        with torch.no_grad():
            rgb_a, depth_a = model(rays)
            cnt_a = model.counters.clone()
            model.counters.zero_()
            model.app_planes_bf16 = True
            rgb_b, depth_b = model(rays)
            cnt_b = model.counters.clone()
        torch.cuda.synchronize()
        assert model._model().app_plane_pair[0] and model._model().app_line_pair[0]
        assert torch.equal(depth_a, depth_b) and torch.equal(cnt_a, cnt_b)
        a, b = rgb_a.cpu().numpy(), rgb_b.cpu().numpy()
        err = np.abs(b - ref["rgb_map"]).max()
        print(f"bf16 planes {regime} G={G} {variant or 'vm'}: max|rgb-oracle|={err:.3e} (fp32 planes: {np.abs(a - ref['rgb_map']).max():.3e}), "
              f"max|bf16 planes - fp32 planes|={np.abs(a - b).max():.3e}")
        assert err <= 1e-2
        tgt = fx.target_rgb(2048)
        assert abs(psnr(b, tgt) - psnr(ref["rgb_map"], tgt)) < 0.01


def test_tensor_core_mlp_fp16_meets_fp32_tolerance(env):
    """TVM_MLP_FP16: the same tcgen05 head with fp16 operands (11 significant bits instead of bf16's 8, fp32 accumulation):
    rgb within the north_star's FP32 tolerance (1e-4), with fp32 or fp16 appearance-plane copies; REFTensoRF included."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    for regime, G, variant in (("R1", 128, None), ("R2", 128, None), ("R1", 300, None), ("R2", 64, "ref")):
        kw = dict(variant=variant, mask_res=G) if variant else {}
        case = fx.make_case(G, 2048, regime, **kw)
        ref = orc.run_case(case, want_stages=False)
        model = gpu_model(pkg, case, mlp_mode="fp16")
        rays = torch.from_numpy(case["rays"]).cuda()
        errs = []
        for planes16 in (False, True):
            model.app_planes_bf16 = planes16
            with torch.no_grad():
                rgb, depth = model(rays)
            torch.cuda.synchronize()
            errs.append(float(np.abs(rgb.cpu().numpy() - ref["rgb_map"]).max()))
            assert np.abs(depth.cpu().numpy() - ref["depth_map"]).max() <= DEPTH_TOL
        print(f"fp16 MLP {regime} G={G} {variant or 'vm'}: max|rgb-oracle| = {errs[0]:.3e} (fp32 planes), {errs[1]:.3e} (fp16 planes)")
        assert max(errs) <= RGB_TOL
    # training in this mode: fp16 forward, bf16 tensor-core backward (TvmModel.tc_weights_bwd): the bf16 gradient bounds
    case = fx.make_case(48, 384, "R2", mask_res=48, train=True)
    S = 167
    d_rgb = (fx.target_rgb(384, seed=7) - 0.5).astype(np.float32)
    ref = orc.backward_case(case, d_rgb_map=d_rgb.astype(np.float64), N_samples=S, white_bg=True)
    model = gpu_model(pkg, case, mlp_mode="fp16")
    rgb, _ = model(torch.from_numpy(case["rays"]).cuda(), is_train=True, N_samples=S, jitter=torch.from_numpy(case["jitter"]).cuda())
    assert np.abs(rgb.detach().cpu().numpy() - ref["rgb_map"]).max() <= RGB_TOL
    pkg._lib.profile_enable(True); pkg._lib.profile_collect()
    (rgb * torch.from_numpy(d_rgb).cuda()).sum().backward()
    torch.cuda.synchronize()
    pkg._lib.profile_enable(False)
    for name, p in (("app_plane.0", model.app_plane[0]), ("density_plane.0", model.density_plane[0]),
                    ("renderModule.mlp.0.weight", model.renderModule.mlp[0].weight), ("basis_mat.weight", model.basis_mat.weight)):
        g, r = p.grad.detach().cpu().numpy().astype(np.float64), ref["grads"][name]
        assert np.linalg.norm(g - r) <= 8e-2 * np.linalg.norm(r), name


@pytest.mark.parametrize("mode,tol", [("fp32", RGB_TOL), ("bf16", 1e-2)])
def test_reftensorf_variant(env, mode, tol):
    """REFTensoRF (models/REFTensoRF.py): extra heads, reflected direction, tint*rgb_s + rgb_d, penalty."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    for regime, train in (("R2", False), ("R1", True)):
        case = fx.make_case(64, 1024, regime, mask_res=64, train=train, variant="ref")
        ref = orc.run_case(case)
        model = gpu_model(pkg, case, mlp_mode=mode)
        assert isinstance(model, pkg.REFTensoRF)
        rays = torch.from_numpy(case["rays"]).cuda()
        jit = None if not train else torch.from_numpy(case["jitter"]).cuda()
        with torch.no_grad():
            rgb, depth = model(rays, is_train=train, jitter=jit)
        torch.cuda.synchronize()
        err = np.abs(rgb.cpu().numpy() - ref["rgb_map"]).max()
        pen = float(model.penalty.item())
        print(f"REF {mode} {regime}: max|rgb-oracle|={err:.3e} penalty {pen:.5f} vs {ref['penalty']:.5f}")
        assert err <= tol
        assert np.abs(depth.cpu().numpy() - ref["depth_map"]).max() <= DEPTH_TOL
        assert abs(pen - ref["penalty"]) <= (1e-4 if mode == "fp32" else 2e-2) * max(1.0, abs(ref["penalty"]))
        if mode == "fp32":
            out = model.forward_with_aux(rays, jitter=jit)
            S = ref["nSamples"]
            assert np.array_equal(pkg.unpack_bits(out["valid_bits"], S), ref["ray_valid"])
            both = pkg.unpack_bits(out["app_bits"], S) & ref["app_mask"]
            assert np.abs(out["rgb"].cpu().numpy()[both] - ref["rgb"][both]).max() <= 2e-5


def test_nerfplusplus_variant(env):
    """NerfPlusPlus (models/nerfplusplus.py): sphere-bounded jittered fg sampling, bg_lambda gate, 512-sample bg MLP."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    for regime, S, G in (("R2", 150, 64), ("R1", 217, 64), ("R0", 100, 32)):
        n = 512
        case = fx.make_case(G, n, regime, mask_res=G, variant="npp")
        case["fg_rand"], case["bg_rand"] = fx.npp_rand(n, S)
        ref = orc.run_case(case, N_samples=S, white_bg=False)
        model = gpu_model(pkg, case)
        assert isinstance(model, pkg.NerfPlusPlus)
        rays = torch.from_numpy(case["rays"]).cuda()
        fg, bg = torch.from_numpy(case["fg_rand"]).cuda(), torch.from_numpy(case["bg_rand"]).cuda()
        out = model.forward_with_aux(rays, N_samples=S, fg_rand=fg, bg_rand=bg)
        torch.cuda.synchronize()
        assert np.array_equal(pkg.unpack_bits(out["bbox_bits"], S), ref["bbox_valid"])
        assert np.array_equal(pkg.unpack_bits(out["valid_bits"], S), ref["ray_valid"])
        lam = out["bg_lambda"].cpu().numpy()
        gate = (lam > 0) != (ref["bg_lambda"] > 0)
        assert gate.sum() <= 1                                     # 0.1 threshold is a float compare
        assert np.abs(lam - ref["bg_lambda"])[~gate].max() <= 1e-5
        act = (lam > 0) & ~gate
        assert np.abs(out["bg_rgb_map"].cpu().numpy()[act] - ref["bg_rgb_map"][act]).max(initial=0) <= RGB_TOL
        err = np.abs(out["rgb_map"].cpu().numpy() - ref["rgb_map"])[~gate].max()
        print(f"NeRF++ {regime}: active bg rays {int(act.sum())}/{n}, max|rgb-oracle|={err:.3e}")
        assert err <= RGB_TOL
        assert np.abs(out["depth_map"].cpu().numpy() - ref["depth_map"]).max() <= DEPTH_TOL
        # production path (ERT on, device-side random draws replaced by the injected ones)
        with torch.no_grad():
            rgb, _ = model(rays, N_samples=S, fg_rand=fg, bg_rand=bg)
        assert np.abs(rgb.cpu().numpy() - ref["rgb_map"])[~gate].max() <= RGB_TOL
        # tensor-core background + appearance head (bf16 operands, fp32 accumulate): north_star tolerance 1e-2
        model.mlp_mode = "bf16"
        out16 = model.forward_with_aux(rays, N_samples=S, fg_rand=fg, bg_rand=bg)
        torch.cuda.synchronize()
        e_bg = np.abs(out16["bg_rgb_map"].cpu().numpy()[act] - ref["bg_rgb_map"][act]).max(initial=0)
        e_rgb = np.abs(out16["rgb_map"].cpu().numpy() - ref["rgb_map"])[~gate].max()
        print(f"NeRF++ {regime} bf16/tcgen05: max|bg_rgb-oracle|={e_bg:.3e}, max|rgb-oracle|={e_rgb:.3e}")
        assert e_bg <= 1e-2 and e_rgb <= 1e-2
        model.mlp_mode = "fp32"


def test_umma_descriptor_conventions(env):
    """tcgen05 known-answer test: K-major and MN-major ("transposed") reads of one core-matrix image."""
    pkg, torch, fx, orc = env
    import ctypes as C
    lib = pkg._lib.load()
    g = torch.Generator().manual_seed(5)
    mk = lambda *s: (torch.randint(-8, 9, s, generator=g).float() / 8.0).cuda()      # exactly representable in bf16
    P, Q, W = mk(128, 128), mk(128, 160), mk(128, 128)
    D1, D2, D3 = torch.zeros(128, 160).cuda(), torch.zeros(128, 128).cuda(), torch.zeros(128, 128).cuda()
    p = lambda t: C.c_void_p(t.data_ptr())
    pkg._lib.check(lib.tvm_selftest_umma(p(P), p(Q), p(W), p(D1), p(D2), p(D3),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tvm_selftest_umma")
    torch.cuda.synchronize()
    assert torch.equal(D3, P @ W.T), "K-major (forward) convention"
    assert torch.equal(D2, P @ W), "MN-major B operand"
    assert torch.equal(D1, P.T @ Q), "MN-major A and B operands"


def _pair_records(src):
    """numpy restatement of the record layout of tvmrender.h (TvmModel.app_plane_pair): src [rows, W, C] ->
    [rows, W, C/4, 2, 4]: record x = texels x and min(x + 1, W - 1), interleaved by groups of 4 channels."""
    rows, W, C = src.shape
    nxt = src[:, np.minimum(np.arange(W) + 1, W - 1), :]
    return np.stack([src.reshape(rows, W, C // 4, 4), nxt.reshape(rows, W, C // 4, 4)], axis=3)


def test_pack_pair16_layout(env):
    """tvm_pack_pair16 known-answer test: record layout and round-to-nearest-even conversion, bit for bit, for both 16-bit
    formats; a plane (rows > 1), a line (rows = 1) and a single-texel row (the record repeats the texel); bad arguments fail."""
    pkg, torch, fx, orc = env
    import ctypes as C
    lib = pkg._lib.load()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator().manual_seed(11)
    for rows, W, Cn in ((5, 7, 48), (1, 300, 48), (3, 1, 16), (1, 2, 4)):
        src = torch.randn(rows, W, Cn, generator=g) * 0.1
        src_d = src.cuda()
        for flag, dt in ((pkg._lib.MLP_BF16, torch.bfloat16), (pkg._lib.MLP_FP16, torch.float16)):
            dst = torch.zeros(rows * W * 2 * Cn, dtype=dt, device="cuda")
            pkg._lib.check(lib.tvm_pack_pair16(C.c_void_p(src_d.data_ptr()), rows, W, Cn, C.c_void_p(dst.data_ptr()), flag, stream),
                           "tvm_pack_pair16")
            torch.cuda.synchronize()
            want = torch.from_numpy(_pair_records(src.numpy())).to(dt).reshape(-1)       # torch rounds to nearest even as well
            assert torch.equal(dst.cpu().view(torch.int16), want.view(torch.int16)), (rows, W, Cn, dt)
    src_d = torch.zeros(64, device="cuda")
    dst = torch.zeros(128, dtype=torch.float16, device="cuda")
    assert lib.tvm_pack_pair16(C.c_void_p(src_d.data_ptr()), 1, 4, 6, C.c_void_p(dst.data_ptr()), pkg._lib.MLP_FP16, stream) != 0   # C % 4
    assert lib.tvm_pack_pair16(C.c_void_p(src_d.data_ptr()), 1, 4, 4, C.c_void_p(dst.data_ptr()), 0, stream) != 0                  # fp32 mode
    assert lib.tvm_pack_pair16(C.c_void_p(src_d.data_ptr()), 1, 4, 4, C.c_void_p(dst.data_ptr() + 2), pkg._lib.MLP_FP16, stream) != 0   # alignment


def test_pack_alpha_bricks3_words(env):
    """tvm_pack_alpha_bricks3 (27-bit neighbourhood word per 8^3 brick) against its numpy restatement, cubic and ragged masks;
    that the words never change which blocks are visited is tests/test_host_emul.py::test_coarse_block_skip_is_conservative."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    import emul_util as eu
    for G, mres in ((40, 40), ((24, 40, 32), (20, 30, 25)), (32, 64)):
        case = fx.make_case(G, 16, "R1", mask_res=mres)
        model = gpu_model(pkg, case)
        want = eu.pack_bricks3(case["alpha_volume"])
        got = model.alphaMask.bricks3.cpu().numpy().view(np.uint32)
        n = want.size - 8
        assert np.array_equal(got[:n], want[:n]) and want[:n].any()


def test_streamed_host_render_matches_resident(env):
    """OctreeRender_trilinear_fast(pinned host rays, out_host=...) pipelines upload / render / download per chunk; the
    pixels are those of the device-resident call, bit for bit (compositing is deterministic)."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    case = fx.make_case(64, 0, "R1", mask_res=64, full_frame=True)
    rays = case["rays"][:300000]
    model = gpu_model(pkg, case, mlp_mode="bf16")
    rays_host = torch.from_numpy(rays).pin_memory()
    rgb_host = torch.empty((rays.shape[0], 3)).pin_memory()
    depth_host = torch.empty((rays.shape[0],)).pin_memory()
    with torch.no_grad():
        rgb, _, depth, _, _ = pkg.OctreeRender_trilinear_fast(rays_host.cuda(), model, white_bg=True, is_train=False)
    out = pkg.OctreeRender_trilinear_fast(rays_host, model, white_bg=True, is_train=False, out_host=(rgb_host, depth_host))
    torch.cuda.synchronize()
    assert out[0] is rgb_host and out[2] is depth_host
    assert torch.equal(rgb_host, rgb.cpu()) and torch.equal(depth_host, depth.cpu())
    assert float(rgb_host.min()) < 0.99      # the frame is not empty


def test_bounded_workspace_overflow_and_relaunch(env):
    """tvmrender.h "Bounded workspaces": a workspace below the worst case bounds the entry list; samples beyond its capacity
    are counted (tvm_forward_entries) but not stored, nothing is written outside the workspace, and the host mirror renders
    the overflowed ranges again -- the pixels equal those of the worst-case workspace bit for bit."""
    import ctypes as C
    pkg, torch, fx, orc = env
    from util import gpu_model
    L = pkg._lib
    lib = L.load()
    case = fx.make_case(64, 0, "R2", mask_res=64, full_frame=True)
    rays = torch.from_numpy(case["rays"][100000:160000]).cuda()
    n = rays.shape[0]
    model = gpu_model(pkg, case, mlp_mode="bf16")
    S = model.nSamples
    model.ws_budget_bytes = 4 * model.workspace_bytes(n, S)
    with torch.no_grad():
        rgb0, dep0 = model(rays, white_bg=True, is_train=False)
    wanted = model.workspace_view(n, S)["n_entries"]
    assert wanted > 4 * n                                   # the fog regime: many weighted samples per ray

    # (a) the C ABI: capacity / bytes are inverse to each other; a launch into half the needed capacity reports the full count,
    #     writes nothing outside, and a launch into exactly the needed capacity reproduces the pixels
    def bounded_bytes(entries):
        out = C.c_size_t(0)
        L.check(lib.tvm_workspace_bytes_bounded(n, S, entries, C.byref(out)), "tvm_workspace_bytes_bounded")
        cap = C.c_uint32(0)
        L.check(lib.tvm_workspace_capacity(n, S, out.value, C.byref(cap)), "tvm_workspace_capacity")
        assert entries <= cap.value <= entries + 64
        return out.value
    for entries, fits in ((wanted // 2, False), (wanted, True)):
        nbytes = bounded_bytes(entries)
        assert nbytes < model.workspace_bytes(n, S) // 4
        ws = torch.full((nbytes + 65536,), 0xA5, dtype=torch.uint8, device="cuda")
        model._ws = ws
        rgb1, dep1 = model._forward_raw(rays, None, model._flags(True) | L.EVAL_ONLY, S, ws_bytes=nbytes)      # as model(rays) above
        got = C.c_uint32(0)
        L.check(lib.tvm_forward_entries(ws.data_ptr(), torch.cuda.current_stream().cuda_stream, C.byref(got)), "tvm_forward_entries")
        assert got.value == wanted
        assert bool((ws[nbytes:] == 0xA5).all())
        assert bool(torch.isfinite(rgb1).all())
        assert torch.equal(rgb1, rgb0) == fits and (not fits or torch.equal(dep1, dep0))
    model._ws = None
    # a workspace that cannot hold one entry per ray is refused, and so is a bounded workspace in the backward pass
    tiny = torch.empty(1 << 16, dtype=torch.uint8, device="cuda")
    model._ws = tiny
    with pytest.raises(L.TvmError, match="workspace too small"):
        model._forward_raw(rays, None, model._flags(True), S, ws_bytes=tiny.numel())
    model._ws = None

    # (b) the host mirror: a budget far below what the frame needs and a hint that is far too low -> overflows, re-renders,
    #     identical pixels; the next render starts from the corrected hint and does not overflow
    model.ws_budget_bytes = 2 * bounded_bytes(wanted // 3)
    model._epr_hint = 1.0
    model.ws_overflows = 0
    with torch.no_grad():
        rgb2, dep2 = model(rays, white_bg=True, is_train=False)
    assert model.ws_overflows >= 1
    assert torch.equal(rgb2, rgb0) and torch.equal(dep2, dep0)
    seen = model.ws_overflows
    with torch.no_grad():
        rgb3, dep3 = model(rays, white_bg=True, is_train=False)
    assert model.ws_overflows == seen and torch.equal(rgb3, rgb0)
    # deferred check (frame sequences): the render returns unverified, verify_renders() repairs the SAME tensors; outputs the
    # caller has dropped are not rendered again
    model.defer_overflow_check = True
    model._epr_hint = 1.0
    with torch.no_grad():
        rgb4, dep4 = model(rays, white_bg=True, is_train=False)
        model(rays, white_bg=True, is_train=False)            # result dropped
    torch.cuda.synchronize()
    assert not torch.equal(rgb4, rgb0)
    assert model.verify_renders() >= 2 and model.verify_renders() == 0
    assert torch.equal(rgb4, rgb0) and torch.equal(dep4, dep0)
    model.defer_overflow_check = False
    seen = model.ws_overflows
    # ... and through the streamed host renderer
    rays_host = rays.cpu().pin_memory()
    rgb_host, depth_host = torch.empty((n, 3)).pin_memory(), torch.empty((n,)).pin_memory()
    model._epr_hint = 1.0
    pkg.OctreeRender_trilinear_fast(rays_host, model, white_bg=True, is_train=False, out_host=(rgb_host, depth_host))
    torch.cuda.synchronize()
    assert model.ws_overflows > seen
    assert torch.equal(rgb_host, rgb0.cpu()) and torch.equal(depth_host, dep0.cpu())
    # streamed + deferred: two frames through the same staging buffers, the first one repaired afterwards (re-uploaded)
    rgb_h2, depth_h2 = torch.empty((n, 3)).pin_memory(), torch.empty((n,)).pin_memory()
    rays_b = rays_host.clone().pin_memory()
    rays_b[:, 0] += 0.01
    model.defer_overflow_check = True
    model._epr_hint = 1.0
    rgb_host.zero_()
    pkg.OctreeRender_trilinear_fast(rays_host, model, white_bg=True, is_train=False, out_host=(rgb_host, depth_host))
    pkg.OctreeRender_trilinear_fast(rays_b, model, white_bg=True, is_train=False, out_host=(rgb_h2, depth_h2))
    assert model.verify_renders() >= 2
    model.defer_overflow_check = False
    assert torch.equal(rgb_host, rgb0.cpu()) and torch.equal(depth_host, dep0.cpu())
    with torch.no_grad():
        rgb_b, dep_b = model(rays_b.cuda(), white_bg=True, is_train=False)
    assert torch.equal(rgb_h2, rgb_b.cpu()) and torch.equal(depth_h2, dep_b.cpu()) and not torch.equal(rgb_h2, rgb_host)


def test_full_frame_properties(env):
    """BASELINE configs[1] at full size (800x800 rays, 300^3 grid, 200^3 mask, S = 1036) through size-independent
    properties: chunking invariance, early-termination error bound, ray-order invariance, background rays, work counters
    identical between the fp32 and tensor-core heads, colour agreement (north_star: 1e-2 abs, PSNR delta < 0.01 dB), and a
    slice of the frame against the oracle."""
    pkg, torch, fx, orc = env
    from util import gpu_model, psnr
    case = fx.make_case(300, 0, "R1", full_frame=True)
    rays_np = case["rays"]
    n = rays_np.shape[0]
    assert n == 640000
    rays = torch.from_numpy(rays_np).cuda()
    model = gpu_model(pkg, case, mlp_mode="fp32")
    assert model.nSamples == 1036
    model.collect_counters = True

    def render(m, r, **kw):
        m.counters.zero_()
        with torch.no_grad():
            rgb, depth = m(r, white_bg=True, is_train=False, **kw)
        torch.cuda.synchronize()
        return rgb, depth, m.counters.cpu().numpy().copy()

    rgb32, dep32, cnt32 = render(model, rays)
    # (a) chunking: the default 3-chunk render equals a 7-chunk render bit for bit
    model.ws_budget_bytes //= 3
    rgb_c, dep_c, cnt_c = render(model, rays)
    model.ws_budget_bytes *= 3
    assert torch.equal(rgb_c, rgb32) and torch.equal(dep_c, dep32) and np.array_equal(cnt_c[:3], cnt32[:3])
    # (b) early ray termination changes colours by far less than the fp32 tolerance
    model.early_termination = False
    rgb_n, dep_n, cnt_n = render(model, rays)
    model.early_termination = True
    assert float((rgb_n - rgb32).abs().max()) <= 2e-6 and cnt_n[pkg._lib.CNT_M_V] >= cnt32[pkg._lib.CNT_M_V]
    # (c) ray order does not matter
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1)).cuda()
    rgb_p, dep_p, _ = render(model, rays[perm].contiguous())
    assert torch.equal(rgb_p, rgb32[perm]) and torch.equal(dep_p, dep32[perm])
    # (d) rays that never meet the mask see the white background exactly; everything stays in [0, 1]
    miss = ~model.filtering_mask(rays, N_samples=1036)
    assert int(miss.sum()) > 10000 and torch.all(rgb32[miss] == 1.0)
    assert float(rgb32.min()) >= 0.0 and float(rgb32.max()) <= 1.0
    # (e) tensor-core head: same samples, colours within the north_star tolerance
    m16 = gpu_model(pkg, case, mlp_mode="bf16")
    m16.collect_counters = True
    rgb16, dep16, cnt16 = render(m16, rays)
    assert np.array_equal(cnt16[:4], cnt32[:4])                      # M_in, M_v, M_a, rays: identical masks
    assert torch.equal(dep16, dep32)
    err = float((rgb16 - rgb32).abs().max())
    ps = psnr(rgb16.cpu().numpy(), rgb32.cpu().numpy())
    # (f) a slice against the oracle, and the PSNR of both heads against it
    sl = slice(n // 2 + 80, n // 2 + 80 + 640)
    ref = orc.run_case(dict(case, rays=rays_np[sl]), want_stages=False)
    e32 = float(np.abs(rgb32[sl].cpu().numpy() - ref["rgb_map"]).max())
    e16 = float(np.abs(rgb16[sl].cpu().numpy() - ref["rgb_map"]).max())
    p32, p16 = psnr(rgb32[sl].cpu().numpy(), ref["rgb_map"]), psnr(rgb16[sl].cpu().numpy(), ref["rgb_map"])
    print(f"full frame: M_in={cnt32[0]} M_v={cnt32[1]} M_a={cnt32[2]}; bf16 vs fp32 max {err:.2e}, PSNR {ps:.1f} dB; "
          f"vs oracle fp32 {e32:.2e} ({p32:.1f} dB) bf16 {e16:.2e} ({p16:.1f} dB)")
    assert err <= 1e-2 and ps >= 60.0
    assert e32 <= 1e-4 and e16 <= 1e-2
    assert np.abs(dep32[sl].cpu().numpy() - ref["depth_map"]).max() <= 1e-3


def test_no_write_outside_caller_buffers(env):
    """Canaries around the caller-owned buffers (workspace, rgb_map, depth_map): the kernels write nothing outside the sizes
    tvm_workspace_bytes / the signatures promise, for ragged ray counts and sample counts that are not multiples of 32
    (compute-sanitizer is not available on this pool)."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    for n, S, mode, variant in ((1, 33, "fp32", None), (777, 139, "bf16", None), (1025, 64, "fp16", None), (130, 97, "bf16", "ref")):
        kw = dict(variant=variant) if variant else {}
        case = fx.make_case(40, n, "R2", mask_res=40, train=True, **kw)
        model = gpu_model(pkg, case, mlp_mode=mode)
        model.app_planes_bf16 = mode != "fp32"
        need = model.workspace_bytes(n, S)
        pad = 1 << 16
        ws = torch.full((need + pad,), 0xA5, dtype=torch.uint8, device="cuda")
        model._ws = ws
        out = torch.full((n * 4 + 64,), float("nan"), dtype=torch.float32, device="cuda")
        rgb, depth = out[16:16 + 3 * n].view(n, 3), out[32 + 3 * n:32 + 4 * n]
        rays = torch.from_numpy(case["rays"]).cuda()
        jit = torch.from_numpy(case["jitter"]).cuda()
        model._forward_raw(rays, jit, model._flags(True) | pkg._lib.EVAL_ONLY, S, out=(rgb, depth))     # compositing inside the head
        torch.cuda.synchronize()
        assert bool((ws[need:] == 0xA5).all()), ("eval-only", n, S, mode)
        assert bool(torch.isnan(out[:16]).all() and torch.isnan(out[16 + 3 * n:32 + 3 * n]).all() and torch.isnan(out[32 + 4 * n:]).all())
        rgb_f = rgb.clone()
        model._forward_raw(rays, jit, model._flags(True), S, out=(rgb, depth))
        torch.cuda.synchronize()
        assert float((rgb - rgb_f).abs().max()) <= 1e-6, (n, S, mode)
        assert model._ws is ws
        assert bool((ws[need:] == 0xA5).all()), (n, S, mode)
        assert bool(torch.isnan(out[:16]).all() and torch.isnan(out[16 + 3 * n:32 + 3 * n]).all() and torch.isnan(out[32 + 4 * n:]).all())
        assert bool(torch.isfinite(rgb).all() and torch.isfinite(depth).all())
        # ... and the backward pass on the same workspace
        d_rgb = torch.ones_like(rgb)
        model._backward_raw(rays, jit, model._flags(True), S, rgb.contiguous(), d_rgb)
        torch.cuda.synchronize()
        assert bool((ws[need:] == 0xA5).all()), ("backward", n, S, mode)


def _params_from_model(model, fx):
    """A trained host model -> synthetic.ModelParams (numpy, reference shapes), so that the oracle can render it."""
    g = lambda t: t.detach().cpu().numpy().astype(np.float32).copy()
    mlp = model.renderModule.mlp
    return fx.ModelParams(aabb=model.aabb.numpy().astype(np.float32).copy(), gridSize=tuple(int(x) for x in model.gridSize),
                          density_plane=[g(p) for p in model.density_plane], density_line=[g(p) for p in model.density_line],
                          app_plane=[g(p) for p in model.app_plane], app_line=[g(p) for p in model.app_line],
                          basis_mat=g(model.basis_mat.weight), mlp_w=[g(mlp[i].weight) for i in (0, 2, 4)],
                          mlp_b=[g(mlp[i].bias) for i in (0, 2, 4)], near_far=tuple(model.near_far),
                          density_shift=float(model.density_shift), distance_scale=float(model.distance_scale),
                          step_ratio=float(model.step_ratio))


def test_sixteen_bit_operands_on_trained_and_scaled_models(env):
    """The 16-bit modes away from the 0.1 N(0,1) synthetic grids they were tuned on.
    (i) A TRAINED model (the reconstruction schedule of examples/reconstruct_synthetic.py: features and activations have the
        magnitudes training produces): the fp16 head, with fp32 planes and with fp16 pair records, measures 0.8e-4 from the
        oracle there -- AT the fp32 tolerance, not comfortably inside it as on the synthetic grids (1.7e-5); asserted <= 2e-4
        (training is not bit-reproducible: float atomics).  bf16 head 0.8e-3, asserted at its 1e-2.  The fp32 head: 1e-6.
    (ii) Random grids scaled x3 / x10 / x30: appearance features grow with the square of the scale, and the positional
        encoding sin(2 f) turns an absolute feature error of 2^-11 |f| into a colour error -- the 16-bit modes are accurate
        to 1e-4 only while |features| = O(1).  What is asserted is the contract: fp32 head <= 1e-4 at every scale; every
        16-bit mode <= 1e-2 up to x10; the measured errors are printed (and quoted in DESIGN.md) so that the limit of
        the fp16-within-1e-4 claim is on record; fp16 (range 65504) does not overflow at x30."""
    pkg, torch, fx, orc = env
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    import reconstruct_synthetic as ex
    h = ex.run(iters=400, res=48, n_views=12, upsamp_list=(200,), update_AlphaMask_list=(150, 300), mlp_mode="fp32", log=lambda s: None)
    model = h["model"]
    assert h["final_psnr"] > 22.0
    mp = _params_from_model(model, fx)
    am = model.alphaMask
    rays = pkg.get_rays_frame(ex.camera(0.9, 0.5), 40, 40, 0.5 * 40 / np.tan(0.5 * 0.6911), blender=False, device=torch.device("cuda:0"))
    case = dict(model=mp, rays=rays.cpu().numpy(), alpha_volume=am.alpha_volume[0, 0].cpu().numpy(), alpha_aabb=am.aabb.numpy().copy(),
                jitter=None, target=None)
    ref = orc.run_case(case, want_stages=False)
    assert float(np.abs(ref["rgb_map"] - 1.0).max()) > 0.3            # the view shows the object
    errs = {}
    for mode, planes16 in (("fp32", False), ("fp16", False), ("fp16", True), ("bf16", True)):
        model.mlp_mode, model.app_planes_bf16 = mode, planes16
        with torch.no_grad():
            rgb, _ = model(rays, white_bg=True, is_train=False)
        errs[(mode, planes16)] = float(np.abs(rgb.cpu().numpy() - ref["rgb_map"]).max())
    print("trained model, max |rgb - oracle|:", {f"{m}{'+16-bit planes' if p else ''}": f"{e:.2e}" for (m, p), e in errs.items()})
    assert errs[("fp32", False)] <= RGB_TOL and errs[("fp16", False)] <= 2e-4 and errs[("fp16", True)] <= 2e-4
    assert errs[("bf16", True)] <= 1e-2
    assert abs(errs[("fp16", True)] - errs[("fp16", False)]) <= 5e-5      # the 16-bit plane copies are not what costs accuracy
    # (ii) scaled random grids
    from util import gpu_model
    table = {}
    for scale in (1.0, 3.0, 10.0, 30.0):
        case = fx.make_case(128, 1024, "R1", grid_scale=0.1 * scale)
        ref = orc.run_case(case, want_stages=False)
        rays = torch.from_numpy(case["rays"]).cuda()
        m = gpu_model(pkg, case)
        for mode, planes16 in (("fp32", False), ("fp16", False), ("fp16", True), ("bf16", True)):
            m.mlp_mode, m.app_planes_bf16 = mode, planes16
            with torch.no_grad():
                rgb, _ = m(rays, white_bg=True, is_train=False)
            assert bool(torch.isfinite(rgb).all()), (scale, mode)
            table[(scale, mode, planes16)] = float(np.abs(rgb.cpu().numpy() - ref["rgb_map"]).max())
        print(f"grid scale x{scale:g}:", {f"{mo}{'+16' if p else ''}": f"{e:.1e}" for (s_, mo, p), e in table.items() if s_ == scale})
        assert table[(scale, "fp32", False)] <= RGB_TOL
        if scale <= 10.0:
            assert max(table[(scale, "fp16", False)], table[(scale, "fp16", True)], table[(scale, "bf16", True)]) <= 1e-2
    assert table[(1.0, "fp16", True)] <= RGB_TOL


def test_degenerate_inputs_production_path(env):
    """The production launch (no aux) on inputs the reference's tests never see, in every head mode:
    (a) a render in which NO sample reaches the appearance head (regime R0: every weight < 1e-4; the tensor-core kernel runs
        with zero tiles) -- equal to the oracle (white background);
    (b) non-finite rays (NaN / inf origins and directions, zero direction) between good ones: the call returns, the good rays
        keep their pixels bit for bit, nothing is written outside the outputs (canaries);
    (c) bad arguments fail loudly with an error string instead of launching."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    import ctypes as C
    # (a)
    case = fx.make_case(64, 600, "R0", mask_res=64)
    case["model"].density_shift = -16.0          # sigma ~ 1e-7: every weight far below the 1e-4 threshold
    ref = orc.run_case(case)
    assert ref["app_mask"].sum() == 0 and ref["ray_valid"].sum() > 0 and ref["weight"].max() > 0
    rays = torch.from_numpy(case["rays"]).cuda()
    for mode in ("fp32", "bf16", "fp16"):
        model = gpu_model(pkg, case, mlp_mode=mode)
        model.app_planes_bf16 = mode != "fp32"
        with torch.no_grad():
            rgb, depth = model(rays)
        torch.cuda.synchronize()
        assert np.abs(rgb.cpu().numpy() - ref["rgb_map"]).max() <= RGB_TOL
        assert np.abs(depth.cpu().numpy() - ref["depth_map"]).max() <= DEPTH_TOL
    # (b)
    case = fx.make_case(64, 512, "R2", mask_res=64)
    good = torch.from_numpy(case["rays"]).cuda()
    bad = good.clone()
    nan, inf = float("nan"), float("inf")
    rows = {3: [nan, 0, 0, 0, 0, 1], 17: [0, 0, 12, nan, nan, nan], 64: [inf, 0, 0, 0, 0, -1], 65: [0, 0, 12, 0, 0, -inf],
            130: [0, 0, 12, 0, 0, 0], 255: [1e30, -1e30, 1e30, 1, 0, 0], 256: [0, 0, 0, 1e-38, 1e-38, 1e-38], 511: [-inf, inf, nan, inf, -inf, nan]}
    for i, v in rows.items():
        bad[i] = torch.tensor(v, dtype=torch.float32)
    keep = np.array([i not in rows for i in range(512)])
    for mode in ("fp32", "fp16"):
        model = gpu_model(pkg, case, mlp_mode=mode)
        model.app_planes_bf16 = mode != "fp32"
        with torch.no_grad():
            rgb0, depth0 = model(good)
            out = torch.full((512 * 4 + 64,), -7.0, device="cuda")          # outputs carved from one buffer with canaries around them
            rgb1, depth1 = out[16:16 + 1536].view(512, 3), out[16 + 1536 + 16:16 + 1536 + 16 + 512]
            model._forward_raw(bad, None, model._flags(True) | pkg._lib.EVAL_ONLY, model.nSamples, out=(rgb1, depth1))      # as model(good) above
        torch.cuda.synchronize()
        o = out.cpu().numpy()
        assert (o[:16] == -7.0).all() and (o[16 + 1536:16 + 1536 + 16] == -7.0).all() and (o[16 + 1536 + 16 + 512:] == -7.0).all()
        assert np.array_equal(rgb1.cpu().numpy()[keep], rgb0.cpu().numpy()[keep])
        assert np.array_equal(depth1.cpu().numpy()[keep], depth0.cpu().numpy()[keep])
    # (c)
    lib, Lb = pkg._lib.load(), pkg._lib
    m = model._model()
    ws = model._workspace(512, model.nSamples)
    args = lambda **kw: [kw.get("model", C.byref(m)), kw.get("rays", C.c_void_p(good.data_ptr())), kw.get("n", 512), kw.get("S", int(model.nSamples)),
                         None, model._flags(True), C.c_void_p(rgb0.data_ptr()), C.c_void_p(depth0.data_ptr()), None, None,
                         kw.get("ws", C.c_void_p(ws.data_ptr())), kw.get("ws_bytes", ws.numel()), None]
    for kw in (dict(n=0), dict(S=0), dict(rays=None), dict(ws=None), dict(ws_bytes=1024), dict(ws=C.c_void_p(ws.data_ptr() + 4)), dict(n=1 << 24, S=1 << 10)):
        rc = lib.tvm_forward(*args(**kw))
        assert rc != 0, kw
        assert len(lib.tvm_last_error()) > 0
    torch.cuda.synchronize()
    with pytest.raises(Lb.TvmError, match="TVM_EVAL_ONLY"):   # the backward pass refuses the flag of a stash-less forward
        model._backward_raw(good, None, model._flags(True) | Lb.EVAL_ONLY, model.nSamples, rgb0, torch.ones_like(rgb0))
    with torch.no_grad():                                   # and the library still renders afterwards
        rgb2, _ = model(good)
    assert np.array_equal(rgb2.cpu().numpy(), rgb0.cpu().numpy())
