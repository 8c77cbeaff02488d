"""Pins the oracle against golden vectors recorded by running the reference's UNMODIFIED python
(tensorf-myc/models/tensorBase.py, tensoRF.py) over oracle/jt_shim -- see tests/golden/make_golden.py.
Reads only the committed .npz files (never /root/reference)."""
import ast
import glob
import os

import numpy as np
import pytest

from oracle import fixtures as fx, tensorf_oracle as orc

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz"))
              if not os.path.basename(p).startswith(("maint_", "grad_")))


def load_case(path):
    g = np.load(path)
    a = [str(x) for x in g["args"]]
    G, n, regime, train, mask_res, white_bg, S = a[:7]
    variant = a[7] if len(a) > 7 else "vm"
    case = fx.make_case(ast.literal_eval(G), int(n), regime, mask_res=ast.literal_eval(mask_res),
                        train=(train == "True"), variant=variant)
    return g, case, white_bg == "True", int(S)


def test_golden_files_present():
    assert len(GOLD) >= 8


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference_python(path):
    g, case, white_bg, S = load_case(path)
    if case["model"].extra.get("variant") == "npp":
        case["fg_rand"], case["bg_rand"] = fx.npp_rand(case["rays"].shape[0], S)
        r = orc.run_case(case, N_samples=S, white_bg=white_bg)
        assert np.array_equal(r["bbox_valid"], g["bbox_valid"])
        assert np.array_equal(r["z_vals"], g["z_vals"])
        assert np.abs(r["rgb_map"] - g["rgb_map"]).max() <= 1e-5
        assert np.abs(r["depth_map"] - g["depth_map"]).max() <= 1e-4
        return
    r = orc.run_case(case, N_samples=S, white_bg=white_bg)
    assert r["nSamples"] == int(g["nSamples"])
    assert np.float32(r["stepSize"]) == g["stepSize"]
    assert np.array_equal(r["bbox_valid"], g["bbox_valid"])
    assert np.array_equal(r["ray_valid"], g["ray_valid"])          # alpha-mask decisions: bit-exact
    assert np.array_equal(r["z_vals"], g["z_vals"])
    assert (r["app_mask"] != g["app_mask"]).sum() <= 1
    assert np.allclose(r["sigma"], g["sigma"], rtol=1e-5, atol=1e-9)
    assert np.abs(r["weight"] - g["weight"]).max() <= 1e-6
    both = r["app_mask"] & g["app_mask"]
    assert np.abs(r["rgb"][both] - g["rgb"][both]).max(initial=0) <= 1e-5
    assert np.abs(r["rgb_map"] - g["rgb_map"]).max() <= 1e-5
    assert np.abs(r["depth_map"] - g["depth_map"]).max() <= 1e-4
    assert np.abs(r["bg_weight"] - g["bg_weight"]).max() <= 1e-6
    if "penalty" in g.files and float(g["penalty"]) != 0:
        assert abs(r["penalty"] - float(g["penalty"])) <= 1e-5 * max(1.0, abs(float(g["penalty"])))


GRAD_GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grad_*.npz")))


@pytest.mark.parametrize("path", GRAD_GOLD, ids=[os.path.basename(p)[:-4] for p in GRAD_GOLD])
def test_oracle_gradients_match_reference_python(path):
    """tests/golden/make_golden_grads.py: the reference's unmodified python differentiated by torch autograd over the shim.
    The oracle's backward_case (what every GPU gradient test is checked against) must give the same gradients for every
    parameter the reference attaches -- including the density gradient NerfPlusPlus receives through bg_lambda."""
    assert len(GRAD_GOLD) >= 3
    g = np.load(path)
    G, n, regime, mask_res, white_bg, S, variant, pw = [str(x) for x in g["args"]]
    case = fx.make_case(ast.literal_eval(G), int(n), regime, mask_res=ast.literal_eval(mask_res), train=True, variant=variant)
    if variant == "npp":
        case["fg_rand"], case["bg_rand"] = fx.npp_rand(int(n), int(S))
    d_rgb = (fx.target_rgb(int(n), seed=13) - 0.5).astype(np.float32)
    import torch
    r = orc.backward_case(case, d_rgb_map=d_rgb.astype(np.float64), dtype=torch.float64, N_samples=int(S),
                          white_bg=(white_bg == "True"), penalty_weight=float(pw))
    assert np.abs(r["rgb_map"] - g["rgb_map"]).max() <= 1e-5
    names = [k[5:] for k in g.files if k.startswith("grad:")]
    assert len(names) >= 19
    worst = {}
    for k in names:
        ref, got = g["grad:" + k].astype(np.float64), r["grads"][k]
        scale = np.abs(ref).max()
        if scale == 0:
            assert np.abs(got).max() == 0, k
            continue
        worst[k] = np.abs(got - ref).max() / scale
    # the golden gradients are fp32 (reference python over torch), the oracle runs in fp64
    bad = {k: v for k, v in worst.items() if v > 2e-4}
    assert not bad, bad
    # every parameter the oracle differentiates is attached in the reference as well (nothing extra, nothing missing)
    extra = [k for k, v in r["grads"].items() if k not in names and np.abs(v).max() > 0]
    assert not extra, extra
