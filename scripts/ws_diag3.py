#!/usr/bin/env python
"""diagnostic: regime R2 (fog) through bounded workspaces: per-frame event times, overflows, plan, in both check modes"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import jittor_myc_nerfs_b200 as pkg
import bench, synthetic as fx
sys.argv = [sys.argv[0]]
args = bench.parse()
case = bench.make_case(args, 0)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rays = torch.from_numpy(case["rays"]).to(dev)
model = pkg.model_from_params(case["model"], "cuda:0", case["alpha_volume"], case["alpha_aabb"], "fp16")
model.app_planes_bf16 = True
model.density_shift = fx.REGIMES["R2"]["density_shift"]
model._model_struct = None
n, S = rays.shape[0], model.nSamples
def step():
    with torch.no_grad():
        return pkg.OctreeRender_trilinear_fast(rays, model, white_bg=True, is_train=False, device=dev)
model.ws_budget_bytes = 80 << 30
ref = step()[0].clone()
torch.cuda.synchronize()
for gib, defer in ((80, False), (2, False), (2, True), (8, True)):
    model.ws_budget_bytes = gib << 30
    model._ws = None; model._ws2 = None; torch.cuda.empty_cache()
    model.defer_overflow_check = defer
    model._epr_hint = 21.7
    for _ in range(3):
        step()
    rep0 = model.verify_renders()
    torch.cuda.synchronize()
    ov0 = model.ws_overflows
    evs = []
    for _ in range(6):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = step(); b.record(); evs.append((a, b))
    rep = model.verify_renders()
    torch.cuda.synchronize()
    same = torch.equal(out[0], ref)
    print(f"budget {gib} GiB defer={defer}: warm repairs {rep0}; frames {[round(a.elapsed_time(b), 2) for a, b in evs]} ms; repaired after {rep}, overflows in loop "
          f"{model.ws_overflows - ov0}; plan {model._plan_launch(n, S)} hint {model._epr_hint:.1f}; pixels equal {same}", flush=True)
