"""Parity of the PRODUCTION instantiation of the march (k_march<AUX=false>: empty-space skipping, window cut at the slab exit,
early ray termination) -- the kernel bench.py times -- against the oracle's masks, at BASELINE's full grid sizes.

(bench.py and every evaluation render add TVM_EVAL_ONLY: the same march without the per-block tables, compositing inside the
appearance head; it is held to the stash launch below: identical counters, entries and weights, pixels within 1e-6.)
The production kernel writes no per-sample parity arrays; what it leaves in the caller's workspace
(tvm_workspace_layout: app_mask bits per 32-sample block, the compacted (ray, sample) entries, weights, acc_map) and its
work counters are compared with the reference's masks (tensorBase.py:491-518):
  * TVM_NO_ERT, skipping on:  CNT_M_V == popcount(ray_valid) exactly (a conservative-skip bug that drops a valid sample
    changes the count), app bits == app_mask up to the ulp flips of the float threshold, entries in the reference's
    compaction order (ray-major, sample-minor), weights of the entries within 2e-6.
  * ERT on: identical app bits (a dropped sample has weight <= T < 1e-7 < 1e-4); the samples dropped per ray are exactly
    those behind the first 32-sample block at whose end the transmittance is below kErtEps = 1e-7 (rays whose T is within
    0.1 % of the threshold may go either way), and their summed reference weight is below 1e-7."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ERT_EPS = 1e-7      # csrc/tvm_common.cuh: kErtEps


@pytest.fixture(scope="module")
def env(built_lib):
    import torch
    from oracle import fixtures as fx, tensorf_oracle as orc
    built_lib._lib.require_cuda()
    return built_lib, torch, fx, orc


def _edge_rays():
    return np.array([[0.3, 0.2, 12.0, 0, 0, -1],            # straight down the z axis: two zero direction components
                     [12.0, 0.1, -0.2, -1, 0, 0],
                     [0.0, 0.0, 0.0, 0.6, 0.8, 0.0],         # starts inside the box (t_min clamps to near)
                     [20.0, 20.0, 20.0, 0.0, 0.0, 1.0],      # misses the box
                     [5.0, 0.0, 12.0, 0, 0, -1],             # grazes the x = +5 face exactly (strict > keeps it inside)
                     [-12.0, 5.0, 5.0, 1, 0, 0],             # runs along an edge
                     [0.0, 0.0, -11.0, 0.0, 0.0, 1.0]], dtype=np.float32)


def _app_bits(pkg, view, S):
    return pkg.unpack_bits(view["blk_mask"], S)


def _production(pkg, torch, model, rays, S, ert, eval_only=False):
    model.early_termination = ert
    model.collect_counters = True
    model.counters.zero_()
    with torch.no_grad():
        rgb, depth = model._forward_raw(rays, None, model._flags(True) | (pkg._lib.EVAL_ONLY if eval_only else 0), S)
    torch.cuda.synchronize()
    cnt = model.counters.cpu().numpy().copy()
    view = model.workspace_view(rays.shape[0], S)
    return rgb.cpu().numpy(), depth.cpu().numpy(), cnt, {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in view.items()}


@pytest.mark.parametrize("G,regime,mode", [(300, "R1", "fp32"), (300, "R2", "fp32"), (128, "R2", "fp16")])
def test_production_march_masks(env, G, regime, mode):
    pkg, torch, fx, orc = env
    from util import gpu_model
    n = 4096
    case = fx.make_case(G, n, regime)
    rays_np = np.ascontiguousarray(np.concatenate([case["rays"][:n - 7], _edge_rays()]))
    case["rays"] = rays_np
    ref = orc.run_case(case)
    S = ref["nSamples"]
    assert S == (1036 if G == 300 else 440)
    model = gpu_model(pkg, case, mlp_mode=mode)
    assert model.empty_space_skipping and model._model().alpha_bricks
    rays = torch.from_numpy(rays_np).cuda()
    tol = 1e-4                                                   # fp32 head, and the fp16 head is held to the same bound

    # ---- TVM_NO_ERT, empty-space skipping on ------------------------------------------------------------------------
    rgb, depth, cnt, v = _production(pkg, torch, model, rays, S, ert=False)
    L = pkg._lib
    assert cnt[L.CNT_M_V] == int(ref["ray_valid"].sum()), (cnt[L.CNT_M_V], int(ref["ray_valid"].sum()))
    app = _app_bits(pkg, v, S)
    flips = int((app != ref["app_mask"]).sum())
    allowed = max(2, int(2e-4 * max(1, ref["app_mask"].sum())))
    assert flips <= allowed, f"{flips} app_mask mismatches"
    assert cnt[L.CNT_M_A] == v["n_entries"] == int(app.sum())
    # a flipped bit sits on the threshold: its reference weight is within an ulp-scale distance of 1e-4
    assert np.abs(ref["weight"][app != ref["app_mask"]] - 1e-4).max(initial=0) <= 2e-6
    # entries: every (ray, k) has its bit set, appears once, and within a ray the entries are in sample order (the
    # compaction order of the reference's boolean indexing); weights match the reference's
    ent = v["ent"].astype(np.int64)
    assert ent.shape[0] == app.sum() and app[ent[:, 0], ent[:, 1]].all()
    flat = ent[:, 0] * S + ent[:, 1]
    assert np.unique(flat).size == flat.size
    order = np.lexsort((np.arange(flat.size), ent[:, 0]))         # stable: by ray, keeping list order within a ray
    assert (np.diff(flat[order]) > 0).all(), "entries of a ray are not in sample order"
    assert np.abs(v["ent_w"] - ref["weight"][ent[:, 0], ent[:, 1]]).max(initial=0) <= 2e-6
    assert np.abs(v["acc"] - ref["acc_map"]).max() <= 2e-5
    assert np.abs(rgb - ref["rgb_map"]).max() <= tol
    assert np.abs(depth - ref["depth_map"]).max() <= 1e-4
    # per-sample colours of the production head at the entries
    both = ref["app_mask"][ent[:, 0], ent[:, 1]]
    assert np.abs(v["ent_rgb"][both] - ref["rgb"][ent[both, 0], ent[both, 1]]).max(initial=0) <= (2e-5 if mode == "fp32" else tol)

    # ---- the same launch with TVM_EVAL_ONLY -- what evaluation renders and bench.py pass: the appearance head composites as it
    #      goes (fixed-point sums per ray), no per-block tables, no per-entry colours.  Same march: identical counters, identical
    #      entries and weights (the list order is the order of the atomics, compare as sets), pixels equal to the stash path's.
    rgb_f, depth_f, cnt_f, vf = _production(pkg, torch, model, rays, S, ert=False, eval_only=True)
    assert np.array_equal(cnt_f, cnt)
    ent_f = vf["ent"].astype(np.int64)
    flat_f = ent_f[:, 0] * S + ent_f[:, 1]
    of, o0 = np.argsort(flat_f), np.argsort(flat)
    assert np.array_equal(flat_f[of], flat[o0]) and np.array_equal(vf["ent_w"][of], v["ent_w"][o0])
    assert np.array_equal(depth_f, depth) and np.array_equal(vf["acc"], v["acc"])
    assert np.abs(rgb_f - rgb).max() <= 1e-6, np.abs(rgb_f - rgb).max()
    assert np.abs(rgb_f - ref["rgb_map"]).max() <= tol

    # ---- early ray termination on ------------------------------------------------------------------------------------
    rgb_e, depth_e, cnt_e, ve = _production(pkg, torch, model, rays, S, ert=True)
    app_e = _app_bits(pkg, ve, S)
    assert np.array_equal(app_e, app), "early termination dropped or added a weighted sample"
    # reference transmittance at the end of every 32-sample block, in float64 from the reference's sigma
    dist = np.full(S, ref["stepSize"] * float(case["model"].distance_scale))
    dist[-1] = 0.0                                                # tensorBase.py:488: the last sample has dist 0
    T = np.cumprod(np.exp(-ref["sigma"].astype(np.float64) * dist[None, :]) + 1e-10, axis=1)
    NB = (S + 31) // 32
    ends = np.minimum(np.arange(NB) * 32 + 31, S - 1)
    T_end = T[:, ends]                                            # [n, NB]
    valid_blk = np.add.reduceat(ref["ray_valid"].astype(np.int64), np.arange(NB) * 32, axis=1)   # valid samples per block

    def kept(thr):
        # blocks up to and including the first VISITED block (one with valid samples: others leave T unchanged and are
        # skipped) whose end transmittance is below thr
        below = (T_end < thr) & (valid_blk > 0)
        first = np.where(below.any(1), below.argmax(1), NB)
        return (np.arange(NB)[None, :] <= first[:, None])
    lo = int((valid_blk * kept(ERT_EPS * 1.001)).sum())           # terminates earliest
    hi = int((valid_blk * kept(ERT_EPS * 0.999)).sum())
    assert lo <= cnt_e[L.CNT_M_V] <= hi, (lo, int(cnt_e[L.CNT_M_V]), hi)
    if regime == "R1":
        assert cnt_e[L.CNT_M_V] < cnt[L.CNT_M_V], "early termination never triggered in the surface regime"
    dropped = ~np.repeat(kept(ERT_EPS * 1.001), 32, axis=1)[:, :S]
    lost = (ref["weight"].astype(np.float64) * dropped).sum(1)
    assert lost.max() < ERT_EPS * 1.01, lost.max()
    assert np.abs(rgb_e - rgb).max() <= 2e-6 and np.abs(depth_e - depth).max() <= 2e-5
    rgb_ef, depth_ef, cnt_ef, _ = _production(pkg, torch, model, rays, S, ert=True, eval_only=True)      # the bench's own launch
    assert np.array_equal(cnt_ef, cnt_e) and np.array_equal(depth_ef, depth_e) and np.abs(rgb_ef - rgb_e).max() <= 1e-6
    print(f"production march G={G} {regime} {mode}: M_v={int(cnt[L.CNT_M_V])} (ERT {int(cnt_e[L.CNT_M_V])}, bounds {lo}..{hi}), "
          f"M_a={int(cnt[L.CNT_M_A])}, app flips {flips}/{allowed}, max lost weight {lost.max():.2e}")
