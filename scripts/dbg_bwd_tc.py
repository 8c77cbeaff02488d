"""Per-tensor error table of the tensor-core backward (k_app_bwd_tc) against the fp64 oracle and the fp32 kernels."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import jittor_myc_nerfs_b200 as pkg
from oracle import fixtures as fx, tensorf_oracle as orc
from util import gpu_model
from test_gpu_backward import _names
n, S = int(sys.argv[1]) if len(sys.argv) > 1 else 640, 167
regime = sys.argv[2] if len(sys.argv) > 2 else "R2"
case = fx.make_case(48, n, regime, mask_res=48, train=True)
d_rgb = (fx.target_rgb(n, seed=7) - 0.5).astype(np.float32)
ref = orc.backward_case(case, d_rgb_map=d_rgb.astype(np.float64), N_samples=S, white_bg=True)
rays = torch.from_numpy(case["rays"]).cuda(); jit = torch.from_numpy(case["jitter"]).cuda()
res = {}
for mode in ("bf16", "fp32"):
    model = gpu_model(pkg, case, mlp_mode=mode)
    rgb, _ = model(rays, is_train=True, white_bg=True, N_samples=S, jitter=jit)
    (rgb * torch.from_numpy(d_rgb).cuda()).sum().backward()
    torch.cuda.synchronize()
    res[mode] = {k: p.grad.detach().cpu().numpy().astype(np.float64) for k, p in _names(model)}
    print(mode, "rgb err", float(np.abs(rgb.detach().cpu().numpy() - ref["rgb_map"]).max()), flush=True)
print(f"{'tensor':32s} {'bf16 maxrel':>12s} {'bf16 L2':>10s} {'fp32 maxrel':>12s} {'bf16-vs-fp32 L2':>16s}")
for k in res["bf16"]:
    r = ref["grads"][k]; a = res["bf16"][k]; b = res["fp32"][k]
    sc = np.abs(r).max()
    print(f"{k:32s} {np.abs(a-r).max()/sc:12.3e} {np.linalg.norm(a-r)/np.linalg.norm(r):10.3e} {np.abs(b-r).max()/sc:12.3e} {np.linalg.norm(a-b)/np.linalg.norm(b):16.3e}")
