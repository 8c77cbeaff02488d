"""GPU parity of tvm_backward (through the C ABI / the autograd node) against the oracle's fp64
autograd gradients (row a12).  Float atomics make the sums order-dependent, so gradients are compared
relative to each tensor's largest entry: max|g - g_ref| <= 1e-4 * max|g_ref|."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GRAD_RTOL = 1e-4


@pytest.fixture(scope="module")
def env(built_lib):
    import torch
    from oracle import fixtures as fx, tensorf_oracle as orc
    built_lib._lib.require_cuda()
    return built_lib, torch, fx, orc


def _names(model):
    n = []
    for k in range(3):
        n.append((f"density_plane.{k}", model.density_plane[k]))
    for k in range(3):
        n.append((f"density_line.{k}", model.density_line[k]))
    for k in range(3):
        n.append((f"app_plane.{k}", model.app_plane[k]))
    for k in range(3):
        n.append((f"app_line.{k}", model.app_line[k]))
    n.append(("basis_mat.weight", model.basis_mat.weight))
    for li in (0, 2, 4):
        n.append((f"renderModule.mlp.{li}.weight", model.renderModule.mlp[li].weight))
        n.append((f"renderModule.mlp.{li}.bias", model.renderModule.mlp[li].bias))
    return n


def _compare(model, ref_grads, rtol=GRAD_RTOL):
    worst = {}
    for name, p in _names(model):
        g = p.grad.detach().cpu().numpy().astype(np.float64)
        r = ref_grads[name]
        scale = np.abs(r).max()
        err = np.abs(g - r).max()
        worst[name] = err / max(scale, 1e-30)
        assert scale > 0, name
        assert err <= rtol * scale, f"{name}: max err {err:.3e} vs scale {scale:.3e}"
    return worst


@pytest.mark.parametrize("regime,white_bg,ert", [("R2", True, False), ("R1", True, True), ("R2", False, True)])
def test_backward_matches_oracle(env, regime, white_bg, ert):
    pkg, torch, fx, orc = env
    from util import gpu_model
    case = fx.make_case(48, 384, regime, mask_res=48, train=True)
    S = 167
    d_rgb = (fx.target_rgb(384, seed=7) - 0.5).astype(np.float32)
    ref = orc.backward_case(case, d_rgb_map=d_rgb.astype(np.float64), N_samples=S, white_bg=white_bg)
    model = gpu_model(pkg, case)
    model.early_termination = ert
    rays = torch.from_numpy(case["rays"]).cuda()
    jit = torch.from_numpy(case["jitter"]).cuda()
    rgb, depth = model(rays, is_train=True, white_bg=white_bg, N_samples=S, jitter=jit)
    assert rgb.requires_grad and not depth.requires_grad
    (rgb * torch.from_numpy(d_rgb).cuda()).sum().backward()
    torch.cuda.synchronize()
    assert np.abs(rgb.detach().cpu().numpy() - ref["rgb_map"]).max() <= 1e-4
    worst = _compare(model, ref["grads"])
    print("worst relative gradient errors:", {k: f"{v:.2e}" for k, v in sorted(worst.items(), key=lambda kv: -kv[1])[:4]})


@pytest.mark.parametrize("regime", ["R1", "R2"])
def test_backward_matches_oracle_config2(env, regime):
    """The same check at the size of BASELINE configs[2]: 4096 rays, 128^3 grid, 128^3 mask, S = cal_n_samples = 443
    (utils.py:61-62), per-ray jitter, white background; fp32 kernels against the oracle's fp64 autograd gradients of every
    parameter, production march (skipping + ERT on).  Bound: 5e-4 of each tensor's largest entry -- ten times more fp32
    terms of mixed sign meet in one texel than in the 384-ray case, in an order the float atomics choose anew every run
    (measured over several runs: 1.2e-4 .. 2.4e-4 on the density planes in the fog regime, <= 2.4e-5 in regime R1)."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    n, S = 4096, 443
    case = fx.make_case(128, n, regime, train=True)
    d_rgb = (fx.target_rgb(n, seed=7) - 0.5).astype(np.float32)
    ref = orc.backward_case(case, d_rgb_map=d_rgb.astype(np.float64), N_samples=S, white_bg=True)
    model = gpu_model(pkg, case)
    rays = torch.from_numpy(case["rays"]).cuda()
    jit = torch.from_numpy(case["jitter"]).cuda()
    rgb, _ = model(rays, is_train=True, white_bg=True, N_samples=S, jitter=jit)
    (rgb * torch.from_numpy(d_rgb).cuda()).sum().backward()
    torch.cuda.synchronize()
    assert np.abs(rgb.detach().cpu().numpy() - ref["rgb_map"]).max() <= 1e-4
    # The oracle runs in fp64; app_mask = weight > 1e-4 (tensorBase.py:513) is a float threshold, and in the fog regime a
    # handful of the 536 k weighted samples sit within fp32 rounding of it: whether such a sample reaches the appearance head
    # is decided differently in fp32 and fp64, and its whole gradient contribution (a few 1e-3 of the largest entry of an
    # appearance plane) appears or not.  The reference computes in fp32, so the same restatement evaluated in fp32 is the
    # second witness: every gradient element must agree with the fp64 OR the fp32 evaluation (measured: the fp32 evaluation
    # differs from fp64 by 4.1e-3 / 3.2e-3 / 2.2e-3 on app_plane 1 / 0 / 2 -- the kernels reproduce exactly those digits
    # against fp64 and sit at ~1e-5 from the fp32 one there).
    ref32 = orc.backward_case(case, d_rgb_map=d_rgb, dtype=torch.float32, N_samples=S, white_bg=True) if regime == "R2" else None
    worst, l2 = {}, {}
    for name, p in _names(model):
        g, r = p.grad.detach().cpu().numpy().astype(np.float64), ref["grads"][name]
        err = np.abs(g - r)
        if ref32 is not None and not name.startswith("density"):
            err = np.minimum(err, np.abs(g - ref32["grads"][name].astype(np.float64)) + 2e-5 * np.abs(r).max())
        worst[name] = float(err.max() / np.abs(r).max())
        l2[name] = float(np.linalg.norm(g - r) / np.linalg.norm(r))
    print(f"configs[2] {regime}: fp32 max error / largest entry:", {k: f"{v:.1e}" for k, v in sorted(worst.items(), key=lambda kv: -kv[1])[:6]},
          "relative L2 vs fp64:", {k: f"{v:.1e}" for k, v in sorted(l2.items(), key=lambda kv: -kv[1])[:4]})
    for name in worst:
        assert worst[name] <= 5e-4, f"{name}: max err {worst[name]:.3e} of the largest entry"
        assert l2[name] <= 1e-3, f"{name}: relative L2 {l2[name]:.3e}"
    # the tensor-core step (bf16 forward + backward) on the same batch, per tensor: relative L2 against the fp64 oracle.
    # Bounds are per tensor class (measured on B200, see the print): density grids are untouched by the 16-bit head
    # except through d rgb; the last layer sees one rounding; hidden layers / basis / appearance grids carry the ReLU
    # sign flips of bf16 pre-activations.
    m16 = gpu_model(pkg, case, mlp_mode="bf16")
    rgb16, _ = m16(rays, is_train=True, white_bg=True, N_samples=S, jitter=jit)
    (rgb16 * torch.from_numpy(d_rgb).cuda()).sum().backward()
    torch.cuda.synchronize()
    assert np.abs(rgb16.detach().cpu().numpy() - ref["rgb_map"]).max() <= 1e-2
    l2 = {}
    for name, p in _names(m16):
        g, r = p.grad.detach().cpu().numpy().astype(np.float64), ref["grads"][name]
        l2[name] = float(np.linalg.norm(g - r) / max(np.linalg.norm(r), 1e-30))
    print(f"configs[2] {regime} bf16 relative L2:", {k: f"{v:.1e}" for k, v in sorted(l2.items(), key=lambda kv: -kv[1])})
    for name, v in l2.items():
        bound = 2e-3 if name.startswith("density") else 1e-2 if name.startswith("renderModule.mlp.4") else 8e-2
        assert v <= bound, f"{name}: relative L2 error {v:.3e} > {bound}"


def test_mse_training_step(env):
    """train.py:228: loss = mean((rgb_map - target)^2); one Adam step changes the render."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    case = fx.make_case(48, 512, "R2", mask_res=48, train=True)
    ref = orc.backward_case(case, N_samples=167)
    model = gpu_model(pkg, case)
    rays = torch.from_numpy(case["rays"]).cuda()
    jit = torch.from_numpy(case["jitter"]).cuda()
    tgt = torch.from_numpy(case["target"]).cuda()
    opt = torch.optim.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
    rgb, _ = model(rays, is_train=True, N_samples=167, jitter=jit)
    loss = torch.mean((rgb - tgt) ** 2)
    loss.backward()
    assert abs(float(loss) - ref["loss"]) <= 1e-5
    _compare(model, ref["grads"])
    opt.step()
    with torch.no_grad():
        rgb2, _ = model(rays, is_train=True, N_samples=167, jitter=jit)
    loss2 = torch.mean((rgb2 - tgt) ** 2)
    assert float(loss2) < float(loss), "one Adam step on the batch must reduce its loss"


@pytest.mark.parametrize("regime,white_bg", [("R2", True), ("R1", False)])
def test_backward_tensor_core_bf16(env, regime, white_bg):
    """k_app_bwd_tc: appearance backward on tcgen05 (bf16 operands, fp32 accumulation).  north_star states 1e-2 for the
    colours of the bf16 MLP mode and nothing for its gradients.  Measured on B200: density grids
    1.6e-4, last layer 2e-3, hidden layers / basis / appearance grids 1.7-3.6e-2 relative L2 -- the latter is dominated by
    ReLU units whose bf16 pre-activation changes sign against fp64 (the usual mixed-precision effect), not by rounding
    of the products (sparse regime R1: up to 5.1e-2).  Bounds: 0.15 of the largest entry element-wise, 8e-2 relative L2; the fp32 kernels stay the
    parity-grade path (1e-4, test_backward_matches_oracle).  (Element-wise bound 0.15: single texels in sparse regimes.)"""
    pkg, torch, fx, orc = env
    from util import gpu_model
    n, S = 640, 167
    case = fx.make_case(48, n, regime, mask_res=48, train=True)
    d_rgb = (fx.target_rgb(n, seed=7) - 0.5).astype(np.float32)
    ref = orc.backward_case(case, d_rgb_map=d_rgb.astype(np.float64), N_samples=S, white_bg=white_bg)
    model = gpu_model(pkg, case, mlp_mode="bf16")
    rays = torch.from_numpy(case["rays"]).cuda()
    jit = torch.from_numpy(case["jitter"]).cuda()
    rgb, _ = model(rays, is_train=True, white_bg=white_bg, N_samples=S, jitter=jit)
    (rgb * torch.from_numpy(d_rgb).cuda()).sum().backward()
    torch.cuda.synchronize()
    assert np.abs(rgb.detach().cpu().numpy() - ref["rgb_map"]).max() <= 1e-2
    worst = _compare(model, ref["grads"], rtol=0.15)
    l2 = {}
    for name, p in _names(model):
        g, r = p.grad.detach().cpu().numpy().astype(np.float64), ref["grads"][name]
        l2[name] = float(np.linalg.norm(g - r) / max(np.linalg.norm(r), 1e-30))
        assert l2[name] <= 8e-2, f"{name}: relative L2 error {l2[name]:.3e}"
        if name.startswith("density"):
            assert l2[name] <= 5e-3, name
    print(f"bf16 backward {regime}: worst max-rel", {k: f"{v:.1e}" for k, v in sorted(worst.items(), key=lambda kv: -kv[1])[:4]},
          "worst L2", {k: f"{v:.1e}" for k, v in sorted(l2.items(), key=lambda kv: -kv[1])[:3]})
    # the fp32 path on the same inputs: the two backward kernels agree to the bf16 tolerance as well
    m32 = gpu_model(pkg, case, mlp_mode="fp32")
    rgb32, _ = m32(rays, is_train=True, white_bg=white_bg, N_samples=S, jitter=jit)
    (rgb32 * torch.from_numpy(d_rgb).cuda()).sum().backward()
    for (name, p), (_, p32) in zip(_names(model), _names(m32)):
        a, b = p.grad, p32.grad
        assert float((a - b).norm() / b.norm()) <= 8e-2, name


def test_bf16_training_tracks_fp32(env):
    """40 Adam steps on one batch in both modes from the same initial parameters: the bf16 tensor-core step (k_app_tc +
    k_app_bwd_tc) must reduce the loss like the fp32 kernels do (final losses within 2 % of each other)."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    case = fx.make_case(48, 1024, "R2", mask_res=48, train=True)
    rays = torch.from_numpy(case["rays"]).cuda()
    jit = torch.from_numpy(case["jitter"]).cuda()
    tgt = torch.from_numpy(case["target"]).cuda()
    final, first = {}, {}
    for mode in ("fp32", "bf16"):
        model = gpu_model(pkg, case, mlp_mode=mode)
        opt = pkg.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
        for it in range(40):
            opt.zero_grad()
            rgb, _ = model(rays, is_train=True, N_samples=167, jitter=jit)
            loss = torch.mean((rgb - tgt) ** 2)
            loss.backward()
            opt.step()
            if it == 0:
                first[mode] = float(loss.detach())
        final[mode] = float(loss.detach())
    print("losses", first, "->", final)
    assert final["fp32"] < 0.8 * first["fp32"] and final["bf16"] < 0.8 * first["bf16"]
    assert abs(final["bf16"] - final["fp32"]) <= 0.02 * final["fp32"]


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_graph_captured_step_matches_eager(env, mode):
    """TrainStepGraph (the whole step of train.py:218-261 replayed as one CUDA graph) against the same step run eagerly:
    6 steps with TV regularisation, decaying learning rates and weights, identical jitter."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    n, S = 1024, 167
    case = fx.make_case(48, n, "R2", mask_res=48, train=True)
    rays = torch.from_numpy(case["rays"]).cuda()
    tgt = torch.from_numpy(case["target"]).cuda()
    jits = [torch.from_numpy(fx.jitter(n, seed=100 + i)).cuda() for i in range(6)]
    tv = pkg.TVLoss()
    losses = {}
    params = {}
    for kind in ("eager", "graph"):
        model = gpu_model(pkg, case, mlp_mode=mode)
        opt = pkg.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
        w_d, w_a = 0.5, 0.25
        g = pkg.TrainStepGraph(model, opt, n, S, white_bg=True, TV_weight_density=w_d, TV_weight_app=w_a) if kind == "graph" else None
        out = []
        for it in range(6):
            if kind == "graph":
                g.set_weights(TV_weight_density=w_d, TV_weight_app=w_a)
                out.append(float(g.step(rays, tgt, jitter=jits[it])))
            else:
                opt.zero_grad()
                rgb, _ = model(rays, is_train=True, white_bg=True, N_samples=S, jitter=jits[it])
                loss = torch.mean((rgb - tgt) ** 2)
                (loss + model.TV_loss_density(tv) * w_d + model.TV_loss_app(tv) * w_a).backward()
                opt.step()
                out.append(float(loss.detach()))
            for grp in opt.param_groups:                       # train.py:263-264
                grp["lr"] = grp["lr"] * 0.9
            w_d, w_a = w_d * 0.9, w_a * 0.9
        losses[kind] = out
        params[kind] = [p.detach().clone() for p in model.parameters()]
    print(mode, losses)
    assert losses["eager"][-1] < losses["eager"][0]
    tol = 1e-5 if mode == "fp32" else 2e-3           # float atomics reorder sums; bf16 ReLU masks amplify it
    assert np.allclose(losses["eager"], losses["graph"], rtol=tol, atol=tol * 1e-2)
    # Adam normalises the update, so an element whose tiny gradient changes sign under a different atomic order moves by
    # up to lr per step: bound the worst element loosely and the mean tightly
    for a, b in zip(params["eager"], params["graph"]):
        assert float((a - b).abs().max()) <= (1e-3 if mode == "fp32" else 2e-2)
        assert float((a - b).abs().mean()) <= (1e-5 if mode == "fp32" else 1e-3)


def test_graph_steps_without_host_sync(env):
    """The host may run any number of replays ahead of the device: per-step scalars (Adam bias correction, decayed learning
    rates, regulariser weights) must reach the step they belong to.  40 steps with a steep lr / weight decay are enqueued
    WITHOUT reading anything back and compared with the same 40 steps synchronised after every replay (identical jitter;
    float atomics are the only source of difference).  A step that picked up a later step's scalars changes the
    trajectory by far more than the tolerance."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    n, S, steps = 1024, 167, 40
    case = fx.make_case(48, n, "R2", mask_res=48, train=True)
    rays = torch.from_numpy(case["rays"]).cuda()
    tgt = torch.from_numpy(case["target"]).cuda()
    jits = [torch.from_numpy(fx.jitter(n, seed=300 + i)).cuda() for i in range(steps)]
    out = {}
    for kind in ("sync", "nosync"):
        model = gpu_model(pkg, case)
        opt = pkg.Adam(model.get_optparam_groups(0.05, 0.002), betas=(0.9, 0.99))
        g = pkg.TrainStepGraph(model, opt, n, S, white_bg=True, TV_weight_density=1.0, TV_weight_app=1.0)
        g.step(rays, tgt, jitter=jits[0])          # capture + step 0
        torch.cuda.synchronize()
        w = 1.0
        losses = torch.zeros(steps, device="cuda")
        for it in range(1, steps):
            for grp in opt.param_groups:
                grp["lr"] = grp["lr"] * 0.8      # steep decay: a scalar from a later step is off by up to 0.8^k
            w *= 0.8
            g.set_weights(TV_weight_density=w, TV_weight_app=w)
            loss = g.step(rays, tgt, jitter=jits[it])
            losses[it] = loss[0]                   # device-side copy, no host read
            if kind == "sync":
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        out[kind] = (losses.cpu().numpy(), [p.detach().clone() for p in model.parameters()], opt.n_step)
    assert out["sync"][2] == out["nosync"][2] == steps
    assert np.allclose(out["sync"][0], out["nosync"][0], rtol=1e-4, atol=1e-7), (out["sync"][0], out["nosync"][0])
    for a, b in zip(out["sync"][1], out["nosync"][1]):
        assert float((a - b).abs().mean()) <= 1e-5


def test_backward_after_another_forward_raises(env):
    """tvm_backward must follow its tvm_forward on the model's workspace: a second forward (here an evaluation render) in
    between, or an optimizer step, makes backward() raise instead of producing silently wrong gradients."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    case = fx.make_case(48, 256, "R2", mask_res=48, train=True)
    model = gpu_model(pkg, case)
    rays = torch.from_numpy(case["rays"]).cuda()
    jit = torch.from_numpy(case["jitter"]).cuda()
    rgb, _ = model(rays, is_train=True, N_samples=167, jitter=jit)
    with torch.no_grad():
        model(rays[:16], N_samples=167)
    with pytest.raises(RuntimeError, match="another forward"):
        rgb.sum().backward()
    rgb, _ = model(rays, is_train=True, N_samples=167, jitter=jit)
    with torch.no_grad():
        model.density_line[0].mul_(1.0001)
    with pytest.raises(RuntimeError, match="parameters"):
        rgb.sum().backward()
    # the normal order still works, twice in a row
    for _ in range(2):
        rgb, _ = model(rays, is_train=True, N_samples=167, jitter=jit)
        rgb.sum().backward()
    assert torch.isfinite(model.density_plane[0].grad).all()


def test_graph_recaptures_after_maintenance(env):
    """updateAlphaMask / a changed mlp_mode replace buffers a captured TrainStepGraph points at: step() notices and records
    the step again (Adam moments and the step counter survive); replaced parameters (upsample) need a new optimizer and
    raise."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    n, S = 512, 167
    case = fx.make_case(48, n, "R2", mask_res=48, train=True)
    model = gpu_model(pkg, case)
    rays = torch.from_numpy(case["rays"]).cuda()
    tgt = torch.from_numpy(case["target"]).cuda()
    opt = pkg.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
    g = pkg.TrainStepGraph(model, opt, n, S, white_bg=True)
    for _ in range(3):
        g.step(rays, tgt)
    torch.cuda.synchronize()
    graph0 = g.graph
    # an explicit re-capture leaves parameters, Adam moments and the step counter exactly as they were (the warm-up steps
    # of the capture protocol are real optimisation steps on scratch state)
    snap_p = [p.detach().clone() for p in model.parameters()]
    snap_m = [(mv[0].clone(), mv[1].clone()) for mv in opt.state.values()]
    g.capture()
    torch.cuda.synchronize()
    assert opt.n_step == 3 and g.graph is not graph0
    assert all(torch.equal(a, p.detach()) for a, p in zip(snap_p, model.parameters()))
    assert all(torch.equal(a[0], mv[0]) and torch.equal(a[1], mv[1]) for a, mv in zip(snap_m, opt.state.values()))
    graph0 = g.graph
    model.updateAlphaMask((48, 48, 48))               # new AlphaGridMask: new bits / bricks / dilated buffers
    l1 = float(g.step(rays, tgt))
    assert g.graph is not graph0 and opt.n_step == 4 and np.isfinite(l1)
    graph1 = g.graph
    g.step(rays, tgt)
    assert g.graph is graph1                           # nothing changed: no re-capture
    model.upsample_volume_grid((64, 64, 64))           # replaces every grid parameter
    with pytest.raises(RuntimeError, match="new optimizer"):
        g.step(rays, tgt)


@pytest.mark.parametrize("regime,pw", [("R2", 0.5), ("R1", 0.0)])
def test_reftensorf_backward(env, regime, pw):
    """REFTensoRF (configs/Scar.txt: model_name = REFTensoRF, normal_vector_penalty_weight = 0.5): gradients of
    sum(rgb_map * d) + pw * penalty w.r.t. every parameter incl. the four heads, against the fp64 oracle."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    n, S = 384, 139
    case = fx.make_case(40, n, regime, mask_res=40, train=True, variant="ref")
    d_rgb = (fx.target_rgb(n, seed=11) - 0.5).astype(np.float32)
    ref = orc.backward_case(case, d_rgb_map=d_rgb.astype(np.float64), N_samples=S, white_bg=True, penalty_weight=pw)
    model = gpu_model(pkg, case)
    assert isinstance(model, pkg.REFTensoRF)
    rays = torch.from_numpy(case["rays"]).cuda()
    jit = torch.from_numpy(case["jitter"]).cuda()
    rgb, _ = model(rays, is_train=True, white_bg=True, N_samples=S, jitter=jit)
    loss = (rgb * torch.from_numpy(d_rgb).cuda()).sum()
    if pw:
        assert abs(float(model.penalty.detach().sum()) - ref["penalty"]) <= 1e-5 * max(1.0, abs(ref["penalty"]))
        loss = loss + pw * model.penalty.sum()
    loss.backward()
    torch.cuda.synchronize()
    names = _names(model) + [(f"{h}_linear.{k}", getattr(getattr(model, h + "_linear"), k))
                             for h in ("normal", "diffuse", "specular") for k in ("weight", "bias")]
    worst = {}
    for name, p in names:
        g, r = p.grad.detach().cpu().numpy().astype(np.float64), ref["grads"][name]
        scale = np.abs(r).max()
        assert scale > 0, name
        worst[name] = np.abs(g - r).max() / scale
        assert worst[name] <= GRAD_RTOL, f"{name}: {worst[name]:.3e}"
    # rho only feeds the unused 1/rho argument of the MLP: zero gradient on both sides
    assert float(model.rho_linear.weight.grad.abs().max()) == 0.0 and np.abs(ref["grads"]["rho_linear.weight"]).max() == 0.0
    print(f"REF backward {regime} pw={pw}: worst", {k: f"{v:.1e}" for k, v in sorted(worst.items(), key=lambda kv: -kv[1])[:4]})
    # the same step on the tensor cores (k_app_tc<REF> + k_app_bwd_tc<REF>): bf16 bounds of test_backward_tensor_core_bf16
    m16 = gpu_model(pkg, case, mlp_mode="bf16")
    rgb16, _ = m16(rays, is_train=True, white_bg=True, N_samples=S, jitter=jit)
    loss16 = (rgb16 * torch.from_numpy(d_rgb).cuda()).sum()
    if pw:
        loss16 = loss16 + pw * m16.penalty.sum()
    loss16.backward()
    torch.cuda.synchronize()
    names16 = _names(m16) + [(f"{h}_linear.{k}", getattr(getattr(m16, h + "_linear"), k))
                             for h in ("normal", "diffuse", "specular") for k in ("weight", "bias")]
    l2 = {}
    for name, p in names16:
        g, r = p.grad.detach().cpu().numpy().astype(np.float64), ref["grads"][name]
        l2[name] = float(np.linalg.norm(g - r) / np.linalg.norm(r))
        # the normal head sees its gradient through the projection (I - n n^T) / |v|: cancellation amplifies the bf16 noise
        assert l2[name] <= (0.15 if name.startswith("normal") else 8e-2), f"bf16 {name}: relative L2 error {l2[name]:.3e}"
    print(f"REF bf16 backward {regime}: worst L2", {k: f"{v:.1e}" for k, v in sorted(l2.items(), key=lambda kv: -kv[1])[:4]})


@pytest.mark.gpu
@pytest.mark.parametrize("regime,G,S", [("R2", 48, 131), ("R0", 32, 100)])
def test_nerfplusplus_backward(env, regime, G, S):
    """NerfPlusPlus under autograd (configs/Scarf.txt trains this variant): gradients of the foreground grids / head
    (through fg and through bg_lambda = prod(1 - alpha + 1e-6)) and of every MLPNet parameter (through the 512-sample
    background compositing and the folded colour layer) against the fp64 oracle."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    n = 192
    case = fx.make_case(G, n, regime, mask_res=G, variant="npp")
    case["fg_rand"], case["bg_rand"] = fx.npp_rand(n, S)
    d_rgb = (fx.target_rgb(n, seed=11) - 0.5).astype(np.float32)
    ref = orc.backward_case(case, d_rgb_map=d_rgb.astype(np.float64), N_samples=S, white_bg=False)
    model = gpu_model(pkg, case)
    rays = torch.from_numpy(case["rays"]).cuda()
    fg, bg = torch.from_numpy(case["fg_rand"]).cuda(), torch.from_numpy(case["bg_rand"]).cuda()
    rgb, depth = model(rays, N_samples=S, fg_rand=fg, bg_rand=bg)
    assert rgb.requires_grad and not depth.requires_grad
    assert np.abs(rgb.detach().cpu().numpy() - ref["rgb_map"]).max() <= 1e-4
    (rgb * torch.from_numpy(d_rgb).cuda()).sum().backward()
    torch.cuda.synchronize()
    names = _names(model) + [(k, v) for k, v in model.named_parameters() if k.startswith("bg_net.")]
    assert len(names) == len(_names(model)) + 14
    worst = {}
    for name, p in names:
        assert p.grad is not None, name
        g, r = p.grad.detach().cpu().numpy().astype(np.float64), ref["grads"][name]
        scale = np.abs(r).max()
        if regime == "R0" and not name.startswith("bg_net."):
            continue      # reference init without a mask: weights < 1e-4, the appearance head never runs (zero gradients)
        assert scale > 0, name
        worst[name] = np.abs(g - r).max() / scale
    print(f"NeRF++ backward {regime}: worst", {k: f"{v:.1e}" for k, v in sorted(worst.items(), key=lambda kv: -kv[1])[:5]})
    for name, w in worst.items():
        # d sigma_j = T_j g'.c_j - suffix_j / (1 - alpha_j) is a difference of O(1) terms and the sigma head sums it over all
        # samples with mixed signs: fp32 (kernel) vs fp64 (oracle) leaves ~5e-4 of the largest entry there, <= 1.5e-4 elsewhere
        tol = 1.5e-3 if "sigma_layers" in name else (3e-4 if name.startswith("bg_net.") else GRAD_RTOL)
        assert w <= tol, f"{name}: {w:.3e}"
    # ... and the same step on the tensor cores (bf16 operands): k_bg_tc forward, k_bg_bwd_tc backward (tvm_bg_bwd_tc.cu)
    m16 = gpu_model(pkg, case, mlp_mode="bf16")
    rgb16, _ = m16(rays, N_samples=S, fg_rand=fg, bg_rand=bg)
    (rgb16 * torch.from_numpy(d_rgb).cuda()).sum().backward()
    torch.cuda.synchronize()
    l2 = {}
    for name, p in [(k, v) for k, v in m16.named_parameters() if k.startswith("bg_net.")]:
        g, r = p.grad.detach().cpu().numpy().astype(np.float64), ref["grads"][name]
        assert np.isfinite(g).all(), name
        l2[name] = float(np.linalg.norm(g - r) / np.linalg.norm(r))
    print(f"NeRF++ backward {regime} bf16 (tensor cores), relative L2:", {k[7:]: f"{v:.1e}" for k, v in sorted(l2.items(), key=lambda kv: -kv[1])})
    for name, v in l2.items():
        assert v <= 8e-2, f"{name}: relative L2 {v:.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("cd,ca,app_dim,view_pe,fea_pe", [(8, 24, 27, 2, 2), (4, 12, 9, 3, 1), (32, 16, 16, 0, 2)])
def test_generic_shapes_forward_and_backward(env, cd, ca, app_dim, view_pe, fea_pe):
    """Channel counts / head sizes other than the shipped 16 / 48 / 27 / pe 2 (the class defaults of TensorVMSplit are 8 / 24):
    the non-specialised instantiations (k_march<.., CD=0>, generic channel loops of the fp32 head and of both backward
    kernels) against the oracle; the tensor-core head refuses them loudly."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    case = fx.make_case(40, 256, "R2", mask_res=40, train=True, cd=cd, ca=ca, app_dim=app_dim, view_pe=view_pe, fea_pe=fea_pe)
    S = 139
    d_rgb = (fx.target_rgb(256, seed=5) - 0.5).astype(np.float32)
    ref_f = orc.run_case(case, N_samples=S)
    ref = orc.backward_case(case, d_rgb_map=d_rgb.astype(np.float64), N_samples=S, white_bg=True)
    model = gpu_model(pkg, case)
    rays = torch.from_numpy(case["rays"]).cuda()
    jit = torch.from_numpy(case["jitter"]).cuda()
    out = model.forward_with_aux(rays, white_bg=True, N_samples=S, jitter=jit)
    assert np.array_equal(pkg.unpack_bits(out["valid_bits"], S), ref_f["ray_valid"])
    assert np.abs(out["rgb_map"].cpu().numpy() - ref_f["rgb_map"]).max() <= 1e-4
    rgb, _ = model(rays, is_train=True, white_bg=True, N_samples=S, jitter=jit)
    (rgb * torch.from_numpy(d_rgb).cuda()).sum().backward()
    torch.cuda.synchronize()
    worst = _compare(model, ref["grads"])
    print(f"generic shapes cd={cd} ca={ca} app_dim={app_dim} pe=({view_pe},{fea_pe}): worst grad", max(worst.values()))
    model.mlp_mode = "bf16"
    with pytest.raises(pkg.TvmError):
        with torch.no_grad():
            model(rays, white_bg=True, N_samples=S)


def test_graph_captured_step_nerfplusplus(env):
    """TrainStepGraph for NerfPlusPlus (configs/Scarf.txt): tvm_forward_npp / tvm_backward_npp with the tensor-core background
    backward, the background network in the optimiser, its packed buffers re-built inside the graph.  The stratified draws come
    from the device generator, so the captured step is compared with eager steps statistically: both reduce the loss of one
    batch over 40 steps to within 15 % of each other, and every parameter group moves."""
    pkg, torch, fx, orc = env
    from util import gpu_model
    n, S = 512, 150
    case = fx.make_case(48, n, "R2", mask_res=48, variant="npp")
    rays = torch.from_numpy(case["rays"]).cuda()
    tgt = torch.from_numpy(fx.target_rgb(n, seed=3)).cuda() * 0.5
    final = {}
    for kind in ("eager", "graph"):
        torch.manual_seed(7)
        model = gpu_model(pkg, case, mlp_mode="bf16")
        p0 = [p.detach().clone() for p in model.parameters()]
        opt = pkg.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
        g = pkg.TrainStepGraph(model, opt, n, S, white_bg=False, TV_weight_density=0.5) if kind == "graph" else None
        tv = pkg.TVLoss()
        losses = []
        for it in range(40):
            if g is not None:
                losses.append(float(g.step(rays, tgt)))
            else:
                opt.zero_grad()
                rgb, _ = model(rays, N_samples=S)
                loss = torch.mean((rgb - tgt) ** 2)
                (loss + model.TV_loss_density(tv) * 0.5).backward()
                opt.step()
                losses.append(float(loss.detach()))
        final[kind] = (np.mean(losses[:3]), np.mean(losses[-3:]))
        assert np.isfinite(losses).all() and final[kind][1] < 0.7 * final[kind][0], (kind, losses[:3], losses[-3:])
        moved = [float((p.detach() - q).abs().max()) for p, q in zip(model.parameters(), p0)]
        assert min(moved) > 0.0, "a parameter tensor did not move"
    print("NeRF++ 40 steps, loss first/last:", final)
    assert abs(final["graph"][1] - final["eager"][1]) <= 0.15 * final["eager"][1]
