"""Gradient golden vectors: the reference's UNMODIFIED python (tensorf-myc/models/*.py over oracle/jt_shim, as
make_golden.py) differentiated by torch autograd -- the role Jittor's autograd plays behind optimizer.backward
(train.py:260).  loss = sum(rgb_map * d_rgb) (+ normal_vector_penalty_weight * tensorf.penalty for REFTensoRF,
train.py:253-255).  Pins the oracle's backward_case, in particular which tensors the reference leaves attached
(NerfPlusPlus takes bg_lambda from the live alpha, nerfplusplus.py:276-278).

    python tests/golden/make_golden_grads.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg          # noqa: E402  (sets up sys.path, imports the reference over the shim)
import torch                      # noqa: E402
import jittor as jt               # noqa: E402
from oracle import fixtures as fx # noqa: E402

CASES = {
    # name: (G, n_rays, regime, mask_res, white_bg, N_samples, variant, penalty_weight)
    "grad_vm_g32_R2": (32, 64, "R2", 32, True, 111, "vm", 0.0),
    "grad_ref_g32_R2": (32, 64, "R2", 32, True, 111, "ref", 0.5),
    "grad_npp_g32_R2": (32, 48, "R2", 32, False, 97, "npp", 0.0),
}


def main():
    for name, (G, n, regime, mask_res, white_bg, S, variant, pw) in CASES.items():
        case = fx.make_case(G, n, regime, mask_res=mask_res, train=True, variant=variant)
        m = mg.build_reference_model(case)
        rays = jt.Var(case["rays"])
        d_rgb = torch.from_numpy((fx.target_rgb(n, seed=13) - 0.5).astype(np.float32))
        if variant == "npp":
            fg_rand, bg_rand = fx.npp_rand(n, S)
            jt._rand_queue += [fg_rand, bg_rand]
            rgb_map, _ = m(rays, white_bg=white_bg, is_train=True, ndc_ray=False, N_samples=S)
        else:
            jt._rand_queue.append(case["jitter"].reshape(-1, 1))
            rgb_map, _ = m(rays, white_bg=white_bg, is_train=True, ndc_ray=False, N_samples=S)
        assert not jt._rand_queue
        loss = (rgb_map * d_rgb).sum()
        if pw:
            loss = loss + pw * m.penalty.sum()
        loss.backward()
        out = {"rgb_map": rgb_map.detach().numpy(), "loss": np.float64(float(loss)),
               "args": np.array([str(G), str(n), regime, str(mask_res), str(white_bg), str(S), variant, str(pw)])}
        n_grad = 0
        for k, p in m.named_parameters():
            if p.grad is not None:
                out["grad:" + k] = p.grad.detach().numpy()
                n_grad += 1
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "->", n_grad, "gradient tensors, loss", float(loss))


if __name__ == "__main__":
    main()
