#!/usr/bin/env python
"""k_march diagnostics: stage times of the frame workload without early ray termination (so that experiment builds whose
density differs visit the same samples).  TVM_LIB selects the library.  scripts/march_diag.py [steps]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import jittor_myc_nerfs_b200 as pkg
import synthetic as fx
L = pkg._lib
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = "cuda:0"
reg = fx.REGIMES["R1"]
mp = fx.make_model(300, density_shift=reg["density_shift"])
rays = torch.from_numpy(fx.frame_rays(azimuth=0.7)).to(dev)
model = pkg.model_from_params(mp, dev, fx.ball_alpha_volume(200), mp.aabb.copy(), "fp16")
model.app_planes_bf16 = True
model.early_termination = os.environ.get("ERT", "0") == "1"
model.collect_counters = True
n, S = rays.shape[0], model.nSamples
model.ws_budget_bytes = max(model.ws_budget_bytes, 2 * model.workspace_bytes(n, S) + (1 << 20))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
with torch.no_grad():
    for _ in range(3):
        model(rays, is_train=False, white_bg=True)
    torch.cuda.synchronize()
    model.counters.zero_()
    L.profile_enable(True); L.profile_collect()
    for _ in range(steps):
        flush.zero_()
        model(rays, is_train=False, white_bg=True)
    torch.cuda.synchronize()
    ms, cnt = L.profile_collect()
    L.profile_enable(False)
c = model.counters.cpu().numpy().astype(np.float64) / steps
print(json.dumps({"lib": os.environ.get("TVM_LIB", "default"), "ert": model.early_termination,
                  "stage_ms": {k: round(v / steps, 4) for k, v in ms.items()} if isinstance(ms, dict) else [round(float(v) / steps, 4) for v in ms],
                  "M_in": c[L.CNT_M_IN], "M_v": c[L.CNT_M_V], "M_a": c[L.CNT_M_A]}))
