"""The kernels' per-sample arithmetic (csrc/tvm_math.cuh compiled for the host) against the oracle:
mask bits exact, colours within tolerance.  Validates index/layout conventions without a GPU."""
import numpy as np
import pytest


@pytest.mark.parametrize("regime,train,G", [("R0", False, 128), ("R1", False, 128), ("R2", True, 64)])
def test_emulated_device_math_matches_oracle(built_lib, regime, train, G):
    from oracle import fixtures as fx, tensorf_oracle as orc
    import emul_util as eu
    case = fx.make_case(G, 256, regime, train=train, mask_res=G)
    e = eu.emul_forward(built_lib, case)
    r = orc.run_case(case)
    assert e["S"] == r["nSamples"]
    assert np.array_equal(e["bbox"].astype(bool), r["bbox_valid"])
    assert np.array_equal(e["valid"].astype(bool), r["ray_valid"])
    assert (e["app"].astype(bool) != r["app_mask"]).sum() <= 1
    assert np.allclose(e["sigma"], r["sigma"], rtol=2e-5, atol=1e-7)
    assert np.abs(e["weight"] - r["weight"]).max() <= 1e-6
    assert np.abs(e["rgb_map"] - r["rgb_map"]).max() <= 1e-5
    assert np.abs(e["depth_map"] - r["depth_map"]).max() <= 1e-4


def test_emulated_edge_rays(built_lib):
    from oracle import fixtures as fx, tensorf_oracle as orc
    import emul_util as eu
    case = fx.make_case((24, 40, 32), 16, "R2", mask_res=(20, 30, 25))
    rays = case["rays"].copy()
    rays[0] = [0.3, 0.2, 12.0, 0, 0, -1]
    rays[1] = [12.0, 0.1, -0.2, -1, 0, 0]
    rays[2] = [0.0, 0.0, 0.0, 0.6, 0.8, 0.0]
    rays[3] = [20.0, 20.0, 20.0, 0.0, 0.0, 1.0]
    rays[4] = [5.0, 0.0, 12.0, 0, 0, -1]
    rays[5] = [-12.0, 5.0, 5.0, 1, 0, 0]
    case["rays"] = rays
    for S in (-1, 33, 1):
        e = eu.emul_forward(built_lib, case, S=S)
        r = orc.run_case(case, N_samples=S)
        assert np.array_equal(e["bbox"].astype(bool), r["bbox_valid"])
        assert np.array_equal(e["valid"].astype(bool), r["ray_valid"])
        assert np.abs(e["rgb_map"] - r["rgb_map"]).max() <= 1e-5


@pytest.mark.parametrize("regime,train,G,mres", [("R1", False, 64, 64), ("R2", True, (24, 40, 32), (20, 30, 25)),
                                                 ("R0", False, 32, None)])
def test_coarse_block_skip_is_conservative(built_lib, regime, train, G, mres):
    """The brick / bbox block test may only drop 32-sample blocks that hold no valid sample."""
    from oracle import fixtures as fx, tensorf_oracle as orc
    import emul_util as eu
    case = fx.make_case(G, 400, regime, train=train, mask_res=mres)
    rays = case["rays"]
    rays[0] = [0.3, 0.2, 12.0, 0, 0, -1]
    rays[1] = [5.0, 0.0, 12.0, 0, 0, -1]
    rays[2] = [0.0, 0.0, 0.0, 0.6, 0.8, 0.0]
    visit, S = eu.emul_block_maybe(built_lib, case)
    # the 3x3x3-dilated brick index is a pure accelerator: the same blocks are visited with and without it
    visit_exact, _ = eu.emul_block_maybe(built_lib, case, bricks3=False)
    assert np.array_equal(visit, visit_exact)
    r = orc.run_case(case)
    valid = r["ray_valid"]
    NB = visit.shape[1]
    pad = np.zeros((valid.shape[0], NB * 32), bool)
    pad[:, :S] = valid
    has_valid = pad.reshape(valid.shape[0], NB, 32).any(-1)
    assert not (has_valid & ~visit).any(), "a block holding a valid sample was skipped"
    if mres == 64:
        assert visit.sum() < 0.6 * visit.size       # and the test does skip most empty blocks


def test_emulated_nerfpp_sampling(built_lib):
    """NerfPlusPlus.sample_ray (sphere-bounded, stratified) in the kernels' arithmetic: masks bit-exact."""
    from oracle import fixtures as fx, tensorf_oracle as orc
    import emul_util as eu
    S = 75
    case = fx.make_case(32, 200, "R2", mask_res=32, variant="npp")
    case["fg_rand"], case["bg_rand"] = fx.npp_rand(200, S)
    e = eu.emul_forward(built_lib, case, S=S, white_bg=False)
    r = orc.run_case(case, N_samples=S, white_bg=False)
    assert np.array_equal(e["bbox"].astype(bool), r["bbox_valid"])
    assert np.array_equal(e["valid"].astype(bool), r["ray_valid"])
    assert np.abs(e["weight"] - r["weight"]).max() <= 1e-6
    assert np.abs(e["rgb_map"] - r["fg_rgb_map"]).max() <= 1e-5


def test_axis_pair_equals_axis_taps(built_lib):
    """k_march gathers through axis_pair (adjacent texel pair, one address per row): identical to axis_taps on every u,
    including the faces (u = 0, u = size-1), lattice nodes and out-of-range coordinates."""
    import ctypes as C
    import emul_util as eu
    lib = eu.build_emul()
    rng = np.random.default_rng(5)
    for size in (2, 3, 17, 300):
        T = rng.standard_normal(size).astype(np.float32)
        u = np.concatenate([rng.uniform(-2.5, size + 1.5, 4000), np.arange(-2, size + 2),
                            np.nextafter(np.float32(size - 1), np.float32(0))[None], np.nextafter(np.float32(size - 1), np.float32(2 * size))[None],
                            np.float32(-1e-7)[None], np.float32(1e-7)[None]]).astype(np.float32)
        a, b = np.zeros_like(u), np.zeros_like(u)
        lib.emul_axis_pair_check.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        assert lib.emul_axis_pair_check(u.ctypes.data, u.size, size, T.ctypes.data, a.ctypes.data, b.ctypes.data) == 0
        assert np.array_equal(a, b), (size, u[a != b][:5])
