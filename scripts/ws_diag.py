#!/usr/bin/env python
"""What does the workspace policy cost per frame?  Frame workload of bench.py, budgets x {check at once, deferred check,
worst-case chunks}: CUDA-event ms per frame (L2 flushed before each, as bench.py times it) and host ms per call (diagnostic)."""
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
for gib, defer in ((2, False), (2, True), (8, False), (8, True), (64, False)):
    model = pkg.model_from_params(case["model"], "cuda:0", case["alpha_volume"], case["alpha_aabb"], "fp16")
    model.app_planes_bf16 = True
    model.ws_budget_bytes = gib << 30
    model.defer_overflow_check = defer
    def step():
        with torch.no_grad():
            return pkg.OctreeRender_trilinear_fast(rays, model, white_bg=True, is_train=False, device=dev)
    for _ in range(4):
        step()
    model.verify_renders()
    torch.cuda.synchronize()
    evs, host = [], 0.0
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); t0 = time.perf_counter(); step(); host += time.perf_counter() - t0; b.record(); evs.append((a, b))
    rep = model.verify_renders()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / 10
    print(f"budget {gib} GiB defer={defer}: {ms:.3f} ms/frame (events), host {host * 100:.3f} ms/call, plan {model._plan_launch(rays.shape[0], model.nSamples)}, "
          f"hint {model._epr_hint}, overflows {model.ws_overflows}, repaired {rep}", flush=True)
    del model
    torch.cuda.empty_cache()
