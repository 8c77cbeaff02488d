import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import jittor_myc_nerfs_b200 as pkg
from oracle import fixtures as fx
from util import gpu_model
n, S = 1024, 139
case = fx.make_case(40, n, "R2", mask_res=40, train=True, variant="ref")
model = gpu_model(pkg, case, mlp_mode=sys.argv[1] if len(sys.argv) > 1 else "fp32")
rays = torch.from_numpy(case["rays"]).cuda(); tgt = torch.from_numpy(case["target"]).cuda()
opt = pkg.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
g = pkg.TrainStepGraph(model, opt, n, S, white_bg=True, TV_weight_density=0.5, normal_vector_penalty_weight=0.5)
orig = pkg._lib.check
def chk(rc, what):
    err = torch.cuda.is_current_stream_capturing()
    if rc != 0:
        print("FAILED", what, flush=True)
    return orig(rc, what)
pkg._lib.check = chk
import jittor_myc_nerfs_b200.tensorf as T, jittor_myc_nerfs_b200.train_ops as O
T.L.check = chk; O.L.check = chk
for i in range(0 if os.environ.get("SKIP1") else 3):
    print("step", i, float(g.step(rays, tgt)), flush=True)
print("---- eager steps first, then a new graph (bench order) ----", flush=True)
model2 = gpu_model(pkg, case, mlp_mode=sys.argv[1] if len(sys.argv) > 1 else "fp32")
opt2 = pkg.Adam(model2.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
tv = pkg.TVLoss()
jit = torch.from_numpy(case["jitter"]).cuda()
for i in range(0 if os.environ.get("NO_EAGER") else 3):
    for p in model2.parameters():
        p.grad = None
    rgb, _ = model2(rays, is_train=True, white_bg=True, N_samples=S, jitter=jit)
    loss = torch.mean((rgb - tgt) ** 2) + 0.5 * model2.penalty.sum() + model2.TV_loss_density(tv) * 2.0
    loss.backward(); opt2.step()
print("eager ok", flush=True)
if not os.environ.get("NO_PROF"):
    model2.collect_counters = True; model2.counters.zero_(); pkg._lib.profile_enable(True); pkg._lib.profile_collect()
    rgb, _ = model2(rays, is_train=True, white_bg=True, N_samples=S, jitter=jit); rgb.sum().backward()
    print(pkg._lib.profile_collect()[0]); pkg._lib.profile_enable(False); model2.collect_counters = False
g2 = pkg.TrainStepGraph(model2, opt2, n, S, white_bg=True, TV_weight_density=2.0, TV_weight_app=float(os.environ.get("TVA", "2.0")),
                        normal_vector_penalty_weight=float(os.environ.get("PEN", "0.5")))
for i in range(2):
    print("graph step", i, float(g2.step(rays, tgt)), flush=True)
