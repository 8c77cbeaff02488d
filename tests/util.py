"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np


def gpu_model(pkg, case, device="cuda:0", mlp_mode="fp32"):
    return pkg.model_from_params(case["model"], device, case["alpha_volume"], case["alpha_aabb"], mlp_mode)


def psnr(a, b):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 99.0 if mse == 0 else -10.0 * np.log10(mse)
