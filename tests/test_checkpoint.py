"""Checkpoint container and key layout of TensorBase.save / load (tensorBase.py:229-271)."""
import os
import pickle
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_container_roundtrip_and_trailer(tmp_path):
    import torch
    import jittor_myc_nerfs_b200 as pkg
    obj = {"kwargs": {"aabb": torch.tensor([[-5., -5, -5], [5, 5, 5]]), "gridSize": [8, 9, 10]},
           "state_dict": {"density_plane.0": torch.arange(24, dtype=torch.float32).reshape(1, 2, 3, 4)},
           "global_step": 7, "optimizer": [{"lr": 0.02}]}
    f = str(tmp_path / "a.th")
    pkg.save_checkpoint(obj, f)
    raw = open(f, "rb").read()
    assert raw.endswith(b"HCAJSLHD")
    back = pkg.load_checkpoint(f)
    assert isinstance(back["kwargs"]["aabb"], np.ndarray) and back["global_step"] == 7
    assert np.array_equal(back["state_dict"]["density_plane.0"], obj["state_dict"]["density_plane.0"].numpy())
    # a bare pickle (no trailer) opens as well; a flipped payload byte is detected
    g = str(tmp_path / "b.th")
    open(g, "wb").write(pickle.dumps({"x": 1}))
    assert pkg.load_checkpoint(g) == {"x": 1}
    bad = bytearray(raw)
    bad[10] ^= 0xFF
    open(g, "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        pkg.load_checkpoint(g)


def test_alpha_mask_bits_follow_the_reference_formula():
    """tensorBase.py:258-267: np.packbits(volume.reshape(-1)) on save, np.unpackbits(...)[:length].reshape(shape) on load."""
    from jittor_myc_nerfs_b200 import checkpoint as ck
    rng = np.random.default_rng(3)
    vol = (rng.random((1, 1, 5, 7, 9)) > 0.5).astype(np.float32)
    d = ck.pack_alpha_volume(vol)
    assert d["alphaMask.shape"] == vol.shape
    assert np.array_equal(d["alphaMask.mask"], np.packbits(vol.astype(bool).reshape(-1)))
    assert np.array_equal(ck.unpack_alpha_volume(d), vol.astype(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["vm", "ref", "npp"])
def test_save_load_renders_identically(tmp_path, variant):
    """save -> the driver's reload sequence (train.py:148-163: kwargs -> constructor -> set_nerfplusplus -> load)
    reproduces the render bit for bit, alpha mask included."""
    import torch
    import jittor_myc_nerfs_b200 as pkg
    from oracle import fixtures as fx
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import gpu_model
    case = fx.make_case(32, 256, "R2", mask_res=24, variant=variant)
    model = gpu_model(pkg, case)
    rays = torch.from_numpy(case["rays"]).cuda()
    S = 97
    kw_render = dict(N_samples=S)
    if variant == "npp":
        g = torch.Generator(device="cuda").manual_seed(1)
        kw_render.update(fg_rand=torch.rand((256, S), device="cuda", generator=g),
                         bg_rand=torch.rand((256, 512), device="cuda", generator=g))
    with torch.no_grad():
        rgb0, dep0 = model(rays, white_bg=variant != "npp", **kw_render)
    f = str(tmp_path / "m.th")
    model.save(f, {"global_step": 41})
    ckpt = pkg.load_checkpoint(f)
    assert ckpt["global_step"] == 41 and "alphaMask.mask" in ckpt
    kwargs = ckpt["kwargs"]
    bg = None
    if "bg_freq" in kwargs:
        bg = [kwargs.pop(k) for k in ("bg_freq", "bg_view_freq", "bg_D", "radii")]
    assert (variant == "npp") == (bg is not None)
    kwargs.update({"device": torch.device("cuda:0")})
    m2 = type(model)(**kwargs)
    if bg is not None:
        m2.set_nerfplusplus(*bg)
    m2.load(ckpt)
    m2.mlp_mode = model.mlp_mode
    with torch.no_grad():
        rgb1, dep1 = m2(rays, white_bg=variant != "npp", **kw_render)
    assert torch.equal(rgb0, rgb1) and torch.equal(dep0, dep1)
    assert set(ckpt["state_dict"]) == set(model.state_dict())


def test_driver_helpers_match_survey_values():
    """utils.N_to_reso / cal_n_samples (tensorf-myc/utils.py:56-62) at the shipped configs' sizes (SURVEY.md §8 shape table)
    and SimpleSampler's permutation refresh (train.py:25-37)."""
    import torch
    from jittor_myc_nerfs_b200.utils import N_to_reso, cal_n_samples, SimpleSampler
    bbox = torch.tensor([[-5.0, -5.0, -5.0], [5.0, 5.0, 5.0]])
    assert N_to_reso(2097156, bbox) == [128, 128, 128]            # configs/Scar.txt: N_voxel_init
    assert N_to_reso(27000000, bbox) in ([300, 300, 300], [299, 299, 299])   # fp32: the schedule's last step evaluates to 299^3
    assert cal_n_samples([128, 128, 128], 0.5) == 443 and cal_n_samples([300, 300, 300], 0.5) == 1039
    coffee = torch.tensor([[-0.2350, -1.7393, -1.5537], [0.2350, 2.0214, 1.4530]])     # configs/Coffee.txt: non-cubic grid
    r = N_to_reso(2097156, coffee)
    assert r[0] < r[2] < r[1] and abs(r[0] * r[1] * r[2] - 2097156) / 2097156 < 0.05
    s = SimpleSampler(10, 4, seed=0)
    ids = [s.nextids().tolist() for _ in range(5)]
    assert all(len(i) == 4 for i in ids)
    assert len(set(ids[0]) & set(ids[1])) == 0                              # no repeats inside one permutation
