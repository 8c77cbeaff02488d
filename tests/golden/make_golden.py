"""Generates tests/golden/*.npz by executing the reference's UNMODIFIED python modules
(/root/reference/tensorf-myc/models/{tensorBase,tensoRF}.py) over oracle/jt_shim (Jittor is absent
from the image; see the shim's docstring for what it supplies).  Run in the build container:

    python tests/golden/make_golden.py

The vectors pin the reference's python logic (op order, indexing, masks, quirks); Jittor's own op
numerics stay assumed (A1-A3, oracle/tensorf_oracle.py).  /root/reference is NOT needed by any test:
the tests read only the committed .npz files.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "jt_shim"))
sys.path.insert(0, "/root/reference/tensorf-myc")

import torch
import jittor as jt
from oracle import fixtures as fx

with contextlib.redirect_stdout(io.StringIO()):
    from models.tensoRF import TensorVMSplit, AlphaGridMask      # the reference, unmodified
    from models.REFTensoRF import REFTensoRF
    from models.nerfplusplus import NerfPlusPlus

CASES = {
    # name: (G, n_rays, regime, train, mask_res, white_bg, N_samples)
    "g48_R0_eval": (48, 96, "R0", False, None, True, -1),
    "g48_R1_eval": (48, 96, "R1", False, 48, True, -1),
    "g48_R2_train_blackbg": (48, 96, "R2", True, 40, False, 167),
    "g32x40x48_R2_eval": ((32, 40, 48), 64, "R2", False, (30, 36, 44), True, -1),
    # REFTensoRF variant (models/REFTensoRF.py): 8th field = variant
    "ref_g40_R2_eval": (40, 96, "R2", False, 40, True, -1, "ref"),
    "ref_g40_R1_train": (40, 96, "R1", True, 32, True, 139, "ref"),
    # NerfPlusPlus variant (models/nerfplusplus.py): fg always jittered, 512-sample background MLP
    "npp_g40_R2": (40, 48, "R2", False, 40, False, 120, "npp"),
    "npp_g40_R1": (40, 48, "R1", False, 32, False, 97, "npp"),
}


def build_reference_model(case):
    p = case["model"]
    with contextlib.redirect_stdout(io.StringIO()):
        cls = {"ref": REFTensoRF, "npp": NerfPlusPlus}.get(p.extra.get("variant"), TensorVMSplit)
        m = cls(jt.Var(p.aabb), list(p.gridSize), "cpu", density_n_comp=list(p.density_n_comp),
                          appearance_n_comp=list(p.app_n_comp), app_dim=p.app_dim, near_far=list(p.near_far),
                          shadingMode="MLP_Fea", alphaMask_thres=0.001, density_shift=p.density_shift,
                          distance_scale=p.distance_scale, pos_pe=6, view_pe=p.view_pe, fea_pe=p.fea_pe,
                          featureC=p.featureC, step_ratio=p.step_ratio, fea2denseAct=p.fea2denseAct)
    with torch.no_grad():
        for k in range(3):
            m.density_plane[k].copy_(torch.from_numpy(p.density_plane[k]))
            m.density_line[k].copy_(torch.from_numpy(p.density_line[k]))
            m.app_plane[k].copy_(torch.from_numpy(p.app_plane[k]))
            m.app_line[k].copy_(torch.from_numpy(p.app_line[k]))
        m.basis_mat.weight.copy_(torch.from_numpy(p.basis_mat))
        for i, li in enumerate((0, 2, 4)):
            m.renderModule.mlp[li].weight.copy_(torch.from_numpy(p.mlp_w[i]))
            m.renderModule.mlp[li].bias.copy_(torch.from_numpy(p.mlp_b[i]))
        if p.extra.get("variant") == "npp":
            e = p.extra
            m.set_nerfplusplus(bg_freq=e["bg_freq"], bg_view_freq=e["bg_view_freq"], bg_D=e["bg_D"], radii=e["radii"])
            cp = lambda lin, wb: (lin.weight.copy_(torch.from_numpy(wb[0])), lin.bias.copy_(torch.from_numpy(wb[1])))
            for i, wb in enumerate(e["bg_base"]):
                cp(m.bg_net.base_layers[i][0], wb)
            cp(m.bg_net.sigma_layers[0], e["bg_sigma"])
            cp(m.bg_net.base_remap_layers[0], e["bg_remap"])
            cp(m.bg_net.rgb_layers[0], e["bg_rgb0"])
            cp(m.bg_net.rgb_layers[2], e["bg_rgb1"])
        if p.extra.get("variant") == "ref":
            for n in ("normal", "diffuse", "specular", "rho"):
                getattr(m, n + "_linear").weight.copy_(torch.from_numpy(p.extra[n + "_w"]))
                getattr(m, n + "_linear").bias.copy_(torch.from_numpy(p.extra[n + "_b"]))
    if case["alpha_volume"] is not None:
        m.alphaMask = AlphaGridMask("cpu", jt.Var(case["alpha_aabb"]), jt.Var(case["alpha_volume"]))
    return m


def main():
    for name, spec in CASES.items():
        G, n, regime, train, mask_res, white_bg, S = spec[:7]
        variant = spec[7] if len(spec) > 7 else "vm"
        case = fx.make_case(G, n, regime, mask_res=mask_res, train=train, variant=variant)
        m = build_reference_model(case)
        rays = jt.Var(case["rays"])
        if variant == "npp":
            fg_rand, bg_rand = fx.npp_rand(n, S)
            jt._rand_queue += [fg_rand, bg_rand]
            with torch.no_grad():
                rgb_map, depth_map = m(rays, white_bg=white_bg, is_train=train, ndc_ray=False, N_samples=S)
                jt._rand_queue += [fg_rand]
                xyz, z_vals, bbox_valid = m.sample_ray(rays[:, :3], rays[:, 3:6], is_train=train, N_samples=S)
            np.savez_compressed(os.path.join(HERE, name + ".npz"), rgb_map=rgb_map.numpy(), depth_map=depth_map.numpy(),
                                bbox_valid=bbox_valid.numpy(), z_vals=z_vals.numpy(), nSamples=np.int64(S),
                                stepSize=np.float32(m.stepSize.item()),
                                args=np.array([str(G), str(n), regime, str(train), str(mask_res), str(white_bg), str(S), variant]))
            print(name, "-> rgb mean", rgb_map.numpy().mean(0))
            continue
        if train:
            jt._rand_queue.append(case["jitter"].reshape(-1, 1))
        with torch.no_grad():
            rgb_map, depth_map, rgb, sigma, alpha, weight, bg_weight = m(
                rays, white_bg=white_bg, is_train=train, ndc_ray=False, N_samples=S, additional_output=True)
            if train:
                jt._rand_queue.append(case["jitter"].reshape(-1, 1))
            xyz, z_vals, bbox_valid = m.sample_ray(rays[:, :3], rays[:, 3:6], is_train=train, N_samples=S)
        assert not jt._rand_queue
        app_mask = weight > m.rayMarch_weight_thres
        out = dict(rgb_map=rgb_map.numpy(), depth_map=depth_map.numpy(), rgb=rgb.numpy(), sigma=sigma.numpy(),
                   alpha=alpha.numpy(), weight=weight.numpy(), bg_weight=bg_weight.numpy(),
                   bbox_valid=bbox_valid.numpy(), ray_valid=(sigma.numpy() > 0), app_mask=app_mask.numpy(),
                   z_vals=np.broadcast_to(z_vals.numpy(), weight.shape).copy(),
                   nSamples=np.int64(m.nSamples), stepSize=np.float32(m.stepSize.item()),
                   penalty=np.float32(float(m.penalty.sum())) if variant == "ref" else np.float32(0),
                   args=np.array([str(G), str(n), regime, str(train), str(mask_res), str(white_bg), str(S), variant]))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "->", os.path.getsize(path) // 1024, "KiB; M_v", int(out["ray_valid"].sum()), "M_a",
              int(out["app_mask"].sum()))


if __name__ == "__main__":
    main()
