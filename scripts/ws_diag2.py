#!/usr/bin/env python
"""diagnostic: per-frame event / host times of the deferred bounded render at 2 GiB for forced hints"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import jittor_myc_nerfs_b200 as pkg
import bench
sys.argv = [sys.argv[0]]
args = bench.parse()
case = bench.make_case(args, 0)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rays = torch.from_numpy(case["rays"]).to(dev)
model = pkg.model_from_params(case["model"], "cuda:0", case["alpha_volume"], case["alpha_aabb"], "fp16")
model.app_planes_bf16 = True
model.ws_budget_bytes = 2 << 30
model.defer_overflow_check = True
def step():
    with torch.no_grad():
        return pkg.OctreeRender_trilinear_fast(rays, model, white_bg=True, is_train=False, device=dev)
for hint in (21.7, 46.0, 21.7, 130.0):
    model._epr_hint = hint
    step(); step()
    torch.cuda.synchronize()
    model._pending_checks = []
    evs, hosts = [], []
    for _ in range(6):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); t0 = time.perf_counter(); step(); hosts.append((time.perf_counter() - t0) * 1e3); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    model._pending_checks = []
    print(f"hint {hint}: plan {model._plan_launch(rays.shape[0], model.nSamples)} events {[round(a.elapsed_time(b), 3) for a, b in evs]} host {[round(h, 3) for h in hosts]}", flush=True)
