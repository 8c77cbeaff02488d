#!/usr/bin/env python
"""bench.py -- TensoRF-VM per-ray rendering hot path on B200.

One "step" = one pass of the hot path over one full 800x800 frame (640,000 rays) of BASELINE.json
configs[1]: 300^3 grid, 200^3 alpha mask, S = nSamples = 1036, white background, synthetic rays and
random-init grids (synthetic.py, seed 20211202).  With N GPUs every rank renders its own frame
of the 8-azimuth orbit (configs[4], weak scaling, no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--regime R1|R2] [--mlp fp32|bf16|fp16]
  python bench.py --impl reference ...     # the CPU restatement of the reference on the host cores

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "TensoRF-VM rays/sec (800x800 frame render, 300^3 grid, alphaMask on)"
UNIT = "rays/s"
FRAME = 800
GRID = 300
MASK_RES = 200


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--regime", default="R1", choices=["R0", "R1", "R2"])
    ap.add_argument("--mlp", default=os.environ.get("TVM_MLP_MODE"), choices=["fp32", "bf16", "fp16"],
                    help="appearance head: fp16 tcgen05 (fp16 operands, fp32 accumulation: measured 2e-5 from the oracle, checked "
                         "against the fp32 tolerance 1e-4; default of the frame workloads), bf16 tcgen05 (tolerance 1e-2; "
                         "forward AND backward on the tensor cores: default of --workload train) or fp32 FMA (1e-4)")
    ap.add_argument("--grid", type=int, default=GRID)
    ap.add_argument("--rays", type=int, default=FRAME * FRAME, help="rays per step (default: the full frame)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="contract numbers only: skip the R2 / alternative-precision / sustained "
                    "/ training blocks of the JSON line (kernel-tuning runs)")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="length of the sustained leg (back-to-back frames, seconds)")
    ap.add_argument("--cpu-sample-chunks", type=int, default=256,
                    help="chunks of 1024 rays the CPU baseline renders (256 = 41%% of the frame, 10-30 s of CPU work)")
    ap.add_argument("--app-planes", default="bf16", choices=["fp32", "bf16"],
                    help="bf16: with --mlp bf16 the tensor-core head gathers bf16 copies of the appearance planes (half the gather "
                         "bytes of the head; same 1e-2 rgb tolerance); density planes, lines, masks and compositing stay fp32")
    ap.add_argument("--variant", default="vm", choices=["vm", "ref", "npp"], help="model of the train workload (ref = REFTensoRF, "
                    "configs/Scar.txt, with normal_vector_penalty_weight 0.5; npp = NerfPlusPlus, configs/Scarf.txt)")
    ap.add_argument("--workload", default="frame", choices=["frame", "train", "npp", "ref", "maintain"],
                    help="frame = BASELINE configs[1] (the contract line); train = configs[2] (4096-ray fwd+bwd step, "
                         "128^3 grid); npp / ref = configs[3] (NeRF++ background / Ref-NeRF appearance, full frame). "
                         "The non-default workloads print the same JSON shape for profiles/, not for the driver.")
    args = ap.parse_args()
    if args.mlp is None:
        args.mlp = "bf16" if args.workload == "train" else "fp16"
    return args


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak():
    """Dense bf16/fp16 tensor throughput the appearance head is held against: the SUSTAINED cuBLAS figure (the head runs inside
    a multi-millisecond step), TFLOP/s."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if "bf16_tflops_sustained" in d:
            return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
        if "bf16_tflops" in d:
            return float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    return 1500.0, "fallback (B200_PROFILING.md)"


DENSE_FLOP_PER_ENTRY = 79712.0          # 2 (144*27 + 150*128 + 128*128 + 128*3), the reference's own contraction
ISSUED_FLOP_PER_ENTRY = 2.0 * (144 * 32 + 160 * 128 + 144 * 128 + 144 * 16)   # as issued on the tensor pipe (padded K / N, bias rows)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_ev = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([x.strip() for x in out.stdout.strip().splitlines()[0].split(",")])
            except Exception:
                pass
            self.stop_ev.wait(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop_ev.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def make_case(args, rank):
    import synthetic as fx
    reg = fx.REGIMES[args.regime]
    model = fx.make_model(args.grid, density_shift=reg["density_shift"])
    rays = fx.frame_rays(azimuth=0.7 + rank * np.pi / 4)       # 8-azimuth orbit of SURVEY §8d
    if args.rays < rays.shape[0]:
        rays = np.ascontiguousarray(rays[:args.rays])
    vol = fx.ball_alpha_volume(MASK_RES if args.grid > 128 else 128) if reg["mask"] else None
    return dict(model=model, rays=rays, alpha_volume=vol, alpha_aabb=model.aabb.copy(), jitter=None, target=None)


def workload_name(args):
    return (f"configs[1]: 800x800 frame ({args.rays} rays), {args.grid}^3 grid, alphaMask "
            f"{MASK_RES if args.grid > 128 else 128}^3 ball r=3.5, regime {args.regime} "
            f"(density_shift={'0' if args.regime == 'R1' else '-3' if args.regime == 'R2' else '-10'}), white_bg")


def cpu_sample(case, n_chunks, chunk=1024):
    """A bounded sample of the frame: n_chunks x 1024 consecutive rays evenly spread over the image."""
    rays = case["rays"]
    n = rays.shape[0]
    n_chunks = max(1, min(n_chunks, n // chunk))
    starts = np.linspace(0, n - chunk, n_chunks).astype(np.int64) // chunk * chunk
    return np.concatenate([rays[s:s + chunk] for s in starts]), f"{n_chunks} chunks x {chunk} rays evenly spread over the frame"


def time_cpu_reference(case, n_chunks, repeats=1):
    """The oracle in its reference-shaped mode (dense [n,S,3] points, boolean-mask compaction, 12
    grid_sample calls, renderer.py's chunk loop with chunk=1024 as renderer.py:50) on all host cores."""
    from oracle import tensorf_oracle as orc
    torch.set_num_threads(os.cpu_count())
    sample, desc = cpu_sample(case, n_chunks)
    m = orc.OracleTensorVMSplit(case["model"], case["alpha_volume"], case["alpha_aabb"],
                                opts=orc.OracleOptions.torch_native())
    rays = torch.from_numpy(sample)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        with torch.no_grad():
            orc.OctreeRender_trilinear_fast(rays, m, chunk=1024, N_samples=-1, white_bg=True, is_train=False)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample.shape[0] / best, best, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    case = make_case(args, 0)
    n_chunks = max(2, args.cpu_sample_chunks // 2)
    for _ in range(args.warmup and 1):
        time_cpu_reference(case, 1)
    vals = []
    t_total = 0.0
    for _ in range(args.steps):
        v, dt, desc = time_cpu_reference(case, n_chunks)
        vals.append(v)
        t_total += dt
    n_sample = n_chunks * 1024
    value = n_sample * len(vals) / t_total
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_total / len(vals), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args), "note": "CPU restatement of the reference (torch CPU, "
                       "reference op sequence); NOT Jittor (absent from the image); each step = " + desc},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def run_maintain(args):
    """SURVEY §8f rows at BASELINE sizes (300^3 grids, 200^3 alpha lattice, 800x800 frame): per-call device time through the
    reference-named host methods and the HBM fraction of each kernel's algorithmic bytes."""
    import jittor_myc_nerfs_b200 as pkg
    import synthetic as fx
    pkg._lib.require_cuda()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    G = args.grid
    mp = fx.make_model(G, density_shift=-3.0, grid_scale=0.5)
    model = pkg.model_from_params(mp, "cuda:0", fx.ball_alpha_volume(MASK_RES), mp.aabb.copy(), "fp32")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    hbm, _ = peaks()

    def timed(fn, reps=args.steps):
        fn()
        ms = 0.0
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        return ms / reps

    rows = {}

    def row(name, ms, nbytes, note):
        rows[name] = {"ms": ms, "algorithmic_GB": nbytes / 1e9, "GBps": nbytes / (ms * 1e-3) / 1e9,
                      "hbm_frac": nbytes / (ms * 1e-3) / 1e9 / hbm, "note": note}

    nvox = MASK_RES ** 3
    ms = timed(lambda: model.updateAlphaMask((MASK_RES,) * 3))
    row("updateAlphaMask_200^3", ms, nvox * (1152.0 + 4 + 4 * 27 + 4 + 0.125),
        "per lattice node: 1152 B of density taps + alpha write + 27-tap pool read + volume write + 1 bit")
    model.alphaMask = pkg.AlphaGridMask(dev, mp.aabb, fx.ball_alpha_volume(MASK_RES))
    model._model_struct = None
    rays = pkg.get_rays_frame(fx.camera_pose(0.7, 0.5), FRAME, FRAME, 0.5 * FRAME / np.tan(0.5 * 0.6911))
    rgbs = torch.zeros((rays.shape[0], 3), device=dev)
    n = rays.shape[0]
    ms = timed(lambda: pkg.get_rays_frame(fx.camera_pose(0.7, 0.5), FRAME, FRAME, 0.5 * FRAME / np.tan(0.5 * 0.6911)))
    row("get_rays_800x800", ms, n * 24.0, "24 B written per ray")
    ms = timed(lambda: model.filtering_mask(rays, bbox_only=True))
    row("filtering_rays_bbox_only", ms, n * 25.0, "24 B ray + 1 B mask")
    ms = timed(lambda: model.filtering_mask(rays, N_samples=256))
    row("filtering_rays_alpha_256", ms, n * (25.0 + 256), "24 B ray + 1 B mask + 256 samples x 8 one-bit taps")
    ms = timed(lambda: model.filtering_rays(rays, rgbs, N_samples=256))
    row("filtering_rays_alpha_256_with_compaction", ms, n * (25.0 + 256 + 2 * 36), "as above + torch boolean compaction of rays/rgbs")
    tv = pkg.TVLoss()
    nd, na = 3 * 16 * G * G, 3 * 48 * G * G

    def tv_step():
        for p in model.parameters():
            p.grad = None
        (model.TV_loss_density(tv) + model.TV_loss_app(tv)).backward()
    ms = timed(tv_step)
    row("TV_loss_density+app_value_and_grad", ms, (nd + na) * 4.0 * 4, "read x, write fused grad, autograd scale + accumulate into .grad")
    opt = pkg.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
    for p in model.parameters():
        p.grad = torch.ones_like(p)
    npar = sum(p.numel() for p in model.parameters())
    ms = timed(lambda: opt.step())
    row("Adam_step_all_parameters", ms, npar * 4.0 * 7, "read p,g,m,v + write p,m,v")
    m128 = pkg.model_from_params(fx.make_model(128), "cuda:0", None, None, "fp32")

    def up():
        m128.upsample_volume_grid((G, G, G))
        m128.gridSize = torch.tensor([128] * 3, dtype=torch.int32)
    planes = [(p.detach().clone(), l.detach().clone()) for p, l in zip(m128.density_plane, m128.density_line)]
    src = {k: [t.detach().clone() for t in getattr(m128, k)] for k in ("density_plane", "density_line", "app_plane", "app_line")}

    def up_fresh():
        for k, v in src.items():
            setattr(m128, k, torch.nn.ParameterList([torch.nn.Parameter(t) for t in v]))
        m128.upsample_volume_grid((G, G, G))
    ms = timed(up_fresh)
    row("upsample_volume_grid_128_to_%d" % G, ms, (nd + na) * 4.0 + 3 * 64 * 128 * 128 * 4.0, "write new planes + read old ones")
    # ---- the kernels alone, as the captured training step / the maintenance calls launch them (no autograd, no allocation) ----
    import ctypes as C
    Lb, lib = pkg._lib, pkg._lib.load()
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    planes = [*model.density_plane, *model.app_plane]
    grads = [torch.zeros_like(p) for p in planes]
    loss = torch.zeros(1, device=dev)
    for ow, name, bpe in ((1, "TV_sweep_kernel_6_planes_overwrite", 8.0), (0, "TV_sweep_kernel_6_planes_accumulate", 12.0)):
        jobs = (Lb.TvmTvJob * 6)(*[Lb.TvmTvJob(p.data_ptr(), g.data_ptr(), p.shape[1], p.shape[2], p.shape[3], 1e-2, None, ow)
                                   for p, g in zip(planes, grads)])
        ms = timed(lambda: Lb.check(lib.tvm_tv_loss_batch(jobs, 6, C.c_void_p(loss.data_ptr()), st()), "tvm_tv_loss_batch"))
        row(name, ms, (nd + na) * bpe, "tvm_tv_loss_batch, one launch for the six planes: read x, " +
            ("write grad" if ow else "read + write grad") + " (what TrainStepGraph / TVLoss launch)")
    srcs = [t.contiguous() for k in ("density_plane", "app_plane") for t in src[k]]
    dsts = [torch.empty((1, t.shape[1], G, G), device=dev) for t in srcs]

    n_up = len(srcs)
    up_args = ((C.c_void_p * n_up)(*[a.data_ptr() for a in srcs]), (C.c_int32 * (3 * n_up))(*[d for a in srcs for d in a.shape[1:]]),
               (C.c_void_p * n_up)(*[b.data_ptr() for b in dsts]), (C.c_int32 * (2 * n_up))(*([G, G] * n_up)))

    def up_kernels():
        Lb.check(lib.tvm_upsample_grids(n_up, up_args[0], up_args[1], up_args[2], up_args[3], st()), "tvm_upsample_grids")
    ms = timed(up_kernels)
    row("upsample_kernels_6_planes_128_to_%d" % G, ms, (nd + na) * 4.0 + 3 * 64 * 128 * 128 * 4.0,
        "tvm_upsample_grids: one launch for the six planes, into preallocated planes")
    emit(({"metric": "SURVEY 8f rows, device ms per call", "unit": "ms", "n_gpus": 1, "steps": args.steps,
                      "config": {"workload": f"maintain: {G}^3 grids, {MASK_RES}^3 alpha lattice, {FRAME}x{FRAME} frame",
                                 "l2": "flushed before every timed call"}, "hbm_peak_GBps": hbm, "rows": rows}))


def measure_train(args, rank, local_rank, world, dist):
    """BASELINE configs[2] (and the training half of configs[4]): one optimisation step of train.py:218-261 -- 4096 rays per
    rank, 128^3 grid, S = cal_n_samples = 443, MSE + TV regularisers (configs/Scar.txt weights) + Adam + re-pack -- captured
    into a CUDA graph and replayed (TrainStepGraph); with N ranks the flat gradient buffer is all-reduced inside the step.
    Same timing rules as the frame: L2 flushed before every step, CUDA events on the launching stream, max over ranks."""
    import jittor_myc_nerfs_b200 as pkg
    import synthetic as fx
    dev = torch.device("cuda", local_rank)
    reg = fx.REGIMES[args.regime]
    G, n = 128, 4096
    mp = fx.make_model(G, density_shift=reg["density_shift"])
    model = pkg.model_from_params(mp, f"cuda:{local_rank}", fx.ball_alpha_volume(128) if reg["mask"] else None, mp.aabb.copy(), args.mlp)
    # the ranks slice ONE seeded global permutation of an 8-view ray pool, as identically seeded SimpleSamplers would
    pool = np.concatenate([fx.subset_rays(n, azimuth=0.7 + v * np.pi / 4) for v in range(8)])
    perm = np.random.default_rng(fx.SEED_BASE + 7).permutation(pool.shape[0])
    r8 = rank % 8
    rays = torch.from_numpy(np.ascontiguousarray(pool[perm[r8 * n:(r8 + 1) * n]])).to(dev)
    S = int(np.linalg.norm(np.asarray(mp.gridSize, dtype=np.float64)) / mp.step_ratio)     # cal_n_samples, utils.py:61-62
    tgt = torch.from_numpy(fx.target_rgb(n)).to(dev)
    if world > 1:
        if os.environ.get("TVM_AR", "peer") == "nccl":       # A/B switch: the NCCL all_reduce the step used in round 1
            model.grad_sync = True
        else:
            model.enable_peer_allreduce(n_ctas=int(os.environ.get("TVM_AR_CTAS", "64")))
    opt = pkg.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
    gstep = pkg.TrainStepGraph(model, opt, n, S, white_bg=True, TV_weight_density=2.0, TV_weight_app=2.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(5):
        gstep.step(rays, tgt)
    steps = max(50, 10 * args.steps)
    evs = []
    barrier()
    with ClockSampler(local_rank) as clk:
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            gstep.step(rays, tgt)
            b.record()
            evs.append((a, b))
        barrier()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    loss = float(gstep.loss.item())
    out = {"workload": f"configs[2]: training step (fwd + bwd + TV + Adam + re-pack, one CUDA graph), {n} rays per rank, {G}^3 grid, "
                       f"S={S}, regime {args.regime}, mlp {args.mlp} (tcgen05 forward and backward)",
           "value": n * world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "n_gpus": world, "scaling": "weak",
           "rays_per_step_per_gpu": n, "loss_after": loss, "clocks": clk.summary(),
           "gradient_exchange": "none (1 rank)" if world == 1 else getattr(model, "grad_sync_kind", "NCCL all_reduce of the flat packed fp32 gradient buffer") +
                                f", {model._grads_packed.numel() * 4 / 1e6:.1f} MB, inside the captured step"}
    del gstep
    return out


def run_side_workload(args):
    """configs[2] (training step) and configs[3] (variants, full frame): same timing rules as the contract
    line (warm-up >= 3, L2 flushed before every timed step, CUDA events on the launching stream)."""
    import jittor_myc_nerfs_b200 as pkg
    import synthetic as fx
    L = pkg._lib
    L.require_cuda()
    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    reg = fx.REGIMES[args.regime]
    train = args.workload == "train"
    G = args.grid if (args.grid != GRID or not train) else 128
    variant = {"npp": "npp", "ref": "ref"}.get(args.workload, args.variant if train else "vm")
    mp = fx.make_model(G, density_shift=reg["density_shift"], variant=variant)
    vol = fx.ball_alpha_volume(MASK_RES if G > 128 else 128) if reg["mask"] else None
    model = pkg.model_from_params(mp, f"cuda:{local_rank}", vol, mp.aabb.copy(), args.mlp)
    # full frames of the variants: 16-bit pair records under the tensor-core head and the deferred overflow check, as the
    # contract line renders them (training reads the fp32 grids and the worst-case workspace)
    model.app_planes_bf16 = (not train) and args.app_planes == "bf16" and args.mlp in ("bf16", "fp16")
    model.defer_overflow_check = not train
    if train:
        n = 4096 if args.rays == FRAME * FRAME else args.rays
        # weak scaling, 4096 rays per rank: the ranks slice ONE seeded global permutation of an 8-view ray pool, as the
        # reference's identically seeded SimpleSampler would hand them out -- every rank sees the same ray
        # distribution; one NCCL all-reduce of the flat gradient buffer per step
        pool = np.concatenate([fx.subset_rays(n, azimuth=0.7 + v * np.pi / 4) for v in range(8)])
        perm = np.random.default_rng(fx.SEED_BASE + 7).permutation(pool.shape[0])
        rays_np = pool[perm[rank * n:(rank + 1) * n]] if world <= 8 else pool[perm[(rank % 8) * n:(rank % 8 + 1) * n]]
        model.grad_sync = world > 1
        S = int(np.linalg.norm(np.asarray(mp.gridSize, dtype=np.float64)) / mp.step_ratio)     # cal_n_samples, utils.py:61-62
        tgt = torch.from_numpy(fx.target_rgb(n)).to(dev)
        jit = torch.from_numpy(fx.jitter(n)).to(dev)
    else:
        rays_np = fx.frame_rays()[:args.rays]
        n, S = rays_np.shape[0], model.nSamples
    rays = torch.from_numpy(np.ascontiguousarray(rays_np)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    extra = {}
    if variant == "npp":
        g = torch.Generator(device=dev).manual_seed(fx.SEED_BASE)
        extra = dict(fg_rand=torch.rand((n, S), device=dev, generator=g), bg_rand=torch.rand((n, 512), device=dev, generator=g))

    full = {"on": False}
    if train:
        # train.py:187 optimiser and the regulariser weights of configs/Scar.txt:38-39 (TV 2.0 / 2.0; L1 and ortho are 0 there)
        opt = pkg.Adam(model.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
        tvreg = pkg.TVLoss()

    def step():
        if train:
            for p in model.parameters():
                p.grad = None
            if variant == "npp":
                rgb, _ = model(rays, is_train=True, N_samples=S, **extra)
            else:
                rgb, _ = model(rays, is_train=True, white_bg=True, N_samples=S, jitter=jit)
            loss = torch.mean((rgb - tgt) ** 2)
            if variant == "ref":
                loss = loss + 0.5 * model.penalty.sum()              # train.py:253-255, configs/Scar.txt:7
            if full["on"]:
                loss = loss + model.TV_loss_density(tvreg) * 2.0 + model.TV_loss_app(tvreg) * 2.0
            loss.backward()
            if full["on"]:
                opt.step()
            return loss
        with torch.no_grad():
            if args.workload == "npp":
                return model(rays, N_samples=S, **extra)
            return model(rays, white_bg=True, is_train=False, N_samples=S)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(steps, fn=None):
        fn = fn or step
        evs = []
        barrier()
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        if model.verify_renders():
            raise SystemExit("bench: a bounded workspace overflowed inside a timed region of a side workload")
        barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step()
        model.verify_renders()          # (corrects the entries-per-ray hint of the bounded workspaces before anything is timed)
    with ClockSampler(local_rank) as clk:
        ms = timed(args.steps)
    model.collect_counters = True
    model.counters.zero_()
    L.profile_enable(True)
    L.profile_collect()
    timed(args.steps)
    stage_ms, stage_cnt = L.profile_collect()
    L.profile_enable(False)
    cnt = model.counters.cpu().numpy().astype(np.float64) / args.steps
    name = {"train": f"configs[2]: training step fwd+bwd (MSE), {n} rays, {G}^3 grid, S={S}, mlp {args.mlp}, model {variant}",
            "npp": f"configs[3]: NerfPlusPlus full frame ({n} rays), {G}^3 grid, 512 background samples/ray",
            "ref": f"configs[3]: REFTensoRF full frame ({n} rays), {G}^3 grid"}[args.workload]
    line = {"metric": f"TensoRF-VM rays/sec ({args.workload})", "value": n * world / (ms / args.steps * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.mlp == "fp32" else (f"f32 + {args.mlp} tensor-core MLP" + (" (forward and backward)" if train and args.mlp == "bf16" else
                                                                     " (forward; bf16 tensor-core backward)" if train and args.mlp == "fp16" else "")), "data": "synthetic",
            "config": {"workload": name + f", regime {args.regime}", "n_samples": S,
                       "app_planes": args.mlp if model.app_planes_bf16 else "fp32",
                       "l2": "flushed before every timed step (256 MiB write)",
                       "per_step_counts": {"M_in": cnt[L.CNT_M_IN], "M_v_gathered": cnt[L.CNT_M_V], "M_a": cnt[L.CNT_M_A],
                                           "bg_rays": cnt[L.CNT_BG_RAYS], "bg_samples": cnt[L.CNT_BG_SAMPLES]}},
            "gpu_launches": int(sum(stage_cnt.values())), "clocks": clk.summary(),
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items() if v},
            "stage_launches_per_step": {k: v / args.steps for k, v in stage_cnt.items() if v}}
    if train:
        # the whole optimisation step of train.py:218-261: + TV regularisers (fused value+grad sweeps) + multi-tensor Adam
        # + re-pack of the updated parameters on the next forward
        full["on"] = True
        for _ in range(3):
            step()
        ms_full = timed(args.steps)
        line["full_step"] = {"ms_per_step": ms_full / args.steps, "rays_per_s": n * world / (ms_full / args.steps * 1e-3),
                             "includes": "fwd + bwd + TV_loss_density + TV_loss_app (weights 2.0, configs/Scar.txt) + Adam over "
                                         "all parameter tensors + re-pack of the updated grids"}
    if train:
        # the same full step captured once into a CUDA graph (TrainStepGraph) and replayed: no host time between kernels
        model.collect_counters = False
        L.profile_enable(False)
        gstep = pkg.TrainStepGraph(model, opt, n, S, white_bg=True, TV_weight_density=2.0, TV_weight_app=2.0,
                                   normal_vector_penalty_weight=0.5 if variant == "ref" else 0.0)
        gstep.step(rays, tgt)
        ms_graph = timed(args.steps, lambda: gstep.step(rays, tgt)) / args.steps
        line["full_step_cuda_graph"] = {"ms_per_step": ms_graph, "rays_per_s": n * world / (ms_graph * 1e-3),
                                        "includes": "as full_step, replayed as one CUDA graph (jitter drawn on the device)"}
    if args.workload == "npp" and stage_ms.get("bg"):
        # dense FLOPs of the background network as issued on the tensor cores (padded K/N), per sample
        flop = 2.0 * 128 * (32 + 144 + 160) + 2.0 * 128 * 80 + 2.0 * 64 * 16
        tf = flop * cnt[L.CNT_BG_SAMPLES] / (stage_ms["bg"] / args.steps * 1e-3) / 1e12
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        peak = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1355.0)))
        line["roofline"] = {"bound": "tensor", "kernel": "k_bg_tc", "achieved": tf, "peak": peak, "unit": "TFLOP/s",
                            "frac": tf / peak, "traffic": None, "flop_per_sample_issued": flop}
    if rank == 0:
        emit(line)
    if dist is not None:
        # NCCL communicators referenced by a captured CUDA graph cannot be torn down cleanly: release the graph, drain the
        # device, and leave without destroy_process_group (it blocks forever otherwise; measured on 2 x B200)
        gstep = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


_JSON_OUT = None


def emit(line):
    """The ONE JSON line, on the process's original stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    args = parse()
    # stdout carries the ONE JSON line and nothing else: keep a handle on the original stdout for it and point fd 1 at stderr,
    # so that whatever a library prints there (e.g. NCCL's "NCCL version ..." banner, a plain printf when the box exports
    # NCCL_DEBUG=VERSION) cannot get in front of it
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if os.environ.get("BENCH_WATCHDOG"):
        # debugging aid: dump every thread's stack and exit if the run is still alive after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG"]), exit=True)
    if args.workload == "maintain":
        return run_maintain(args)
    if args.workload != "frame":
        return run_side_workload(args)
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import jittor_myc_nerfs_b200 as pkg
    L = pkg._lib
    L.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    case = make_case(args, rank)
    model = pkg.model_from_params(case["model"], f"cuda:{local_rank}", case["alpha_volume"], case["alpha_aabb"], args.mlp)
    model.app_planes_bf16 = args.app_planes == "bf16" and args.mlp in ("bf16", "fp16")      # 16-bit copies in the mode's format
    n = case["rays"].shape[0]
    S = model.nSamples
    rays_host = torch.from_numpy(case["rays"]).pin_memory()
    rays_dev = rays_host.to(dev)
    rgb_host = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    depth_host = torch.empty((n,), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_resident():
        with torch.no_grad():
            return pkg.OctreeRender_trilinear_fast(rays_dev, model, chunk=1024, N_samples=-1, white_bg=True,
                                                   is_train=False, device=dev)

    def step_e2e():
        # host rays in, host rgb/depth out: upload, render and download pipelined per workspace chunk
        pkg.OctreeRender_trilinear_fast(rays_host, model, chunk=1024, N_samples=-1, white_bg=True, is_train=False,
                                        device=dev, out_host=(rgb_host, depth_host))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # Back-to-back frames: the bounded-workspace overflow check of the resident renders is deferred (the host enqueues frame
    # k+1 while frame k runs) and made by model.verify_renders() when the K frames are done -- the documented usage for frame
    # sequences (tvmrender.h "Bounded workspaces", TensorVMSplit.verify_renders).  A range that had to be rendered again would
    # not be inside the events: such a region is timed again (the hint is corrected by then) and the count is reported.
    model.defer_overflow_check = True
    repairs = {"timed": 0}

    def timed(fn, steps, attempts=4):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed before each."""
        evs = []
        model.verify_renders()
        barrier()
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        repaired = model.verify_renders()
        barrier()
        if repaired:
            # an overflowed range did less work inside the events than the frame needs: never a measurement
            repairs["timed"] += repaired
            if attempts <= 1:
                raise SystemExit("bench: bounded-workspace overflows inside every attempt of a timed region")
            return timed(fn, steps, attempts - 1)
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
        step_e2e()
    torch.cuda.synchronize()

    if os.environ.get("TVM_STREAM_STAGES"):            # experiments: relative stage sizes, e.g. "1,7,7,1" or "8"
        v = [int(x) for x in os.environ["TVM_STREAM_STAGES"].split(",")]
        model.stream_stages = v[0] if len(v) == 1 else tuple(v)
    with ClockSampler(local_rank) as clk:
        ms_total = timed(step_resident, args.steps)
        ms_e2e = timed(step_e2e, args.steps)
    clocks = clk.summary()
    # the same host-to-host frame with the overflow check made inside every call (one host synchronisation per frame)
    model.defer_overflow_check = False
    ms_e2e_checked = timed(step_e2e, args.steps)
    model.defer_overflow_check = True

    def profiled(steps):
        """K resident steps with per-kernel cudaEvents (tvm_profile_*) and the kernels' own work counters."""
        model.collect_counters = True
        # one launch per kernel for the whole frame (a 29 GB workspace): per-kernel times of the two-stream chunk pipeline
        # overlap each other and would not be launch durations
        budget = model.ws_budget_bytes
        model.ws_budget_bytes = max(budget, 2 * model.workspace_bytes(n, S) + (1 << 20))
        step_resident()
        torch.cuda.synchronize()
        model.counters.zero_()              # the counters cover exactly the K profiled steps
        L.profile_enable(True)
        L.profile_collect()
        timed(step_resident, steps)
        st_ms, st_cnt = L.profile_collect()
        L.profile_enable(False)
        model.collect_counters = False
        model.ws_budget_bytes = budget
        model._ws = None
        model._ws2 = None
        c = model.counters.cpu().numpy().astype(np.float64) / steps
        return st_ms, st_cnt, c

    # roofline leg: the same K steps with per-kernel events and work counters
    stage_ms, stage_cnt, cnt = profiled(args.steps)
    M_in, M_v, M_a = cnt[L.CNT_M_IN], cnt[L.CNT_M_V], cnt[L.CNT_M_A]

    # L2 gather peak: the density planes (17.3 MB) and the whole model (69.5 MB) fit the 126 MB L2, so the gather
    # stage is also quoted against pure 64-byte random gathers from an L2-resident buffer of those sizes
    import ctypes as C
    l2_peaks = {}
    sink = torch.zeros(4, device=dev)
    for name, mb in (("17MB", 17.3), ("70MB", 69.5)):
        nfl = int(mb * 1e6 / 4) // 16 * 16
        gbuf = torch.empty(nfl, dtype=torch.float32, device=dev).normal_()
        groups, iters = 148 * 8 * 64 * 8, 16
        run = lambda: L.check(L.load().tvm_bench_gather(C.c_void_p(gbuf.data_ptr()), nfl, groups, iters, C.c_void_p(sink.data_ptr()),
                                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tvm_bench_gather")
        run(); run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        l2_peaks[name] = groups * iters * 18 * 64.0 / (a.elapsed_time(b) * 1e-3) / 1e9
        del gbuf

    # correctness of what was timed: a slice of the frame against the oracle (outside the timed region)
    def oracle_check(the_case):
        from oracle import tensorf_oracle as orc
        sl = slice(n // 2, n // 2 + 512)
        ref = orc.run_case(dict(the_case, rays=the_case["rays"][sl]), want_stages=False)
        got = step_resident()
        model.verify_renders()
        return float(np.abs(got[0][sl].cpu().numpy() - ref["rgb_map"]).max())
    check = oracle_check(case) if rank == 0 else None

    total_rays = n * world
    value = total_rays / (ms_total / args.steps * 1e-3)
    e2e_value = total_rays / (ms_e2e / args.steps * 1e-3)
    hbm_peak, peak_src = peaks()
    tc_peak, tc_src = tensor_peak()
    traffic_json = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and args.grid == GRID and args.rays == FRAME * FRAME and args.regime == "R1":
        traffic_json = json.load(open(tpath))

    def app_roofline(st_ms, st_cnt, m_a, steps, traffic_key):
        """Appearance head = the dense contractions of the path (basis_mat + MLPRender_Fea): tensor-bound by construction.
        achieved = the reference's own FLOPs (79,712 per weighted sample) x entries of one launch / launch time."""
        launches = max(1, st_cnt["app"])
        ms = st_ms["app"] / launches
        per_launch = m_a * steps / launches
        tf = DENSE_FLOP_PER_ENTRY * per_launch / (ms * 1e-3) / 1e12 if ms else 0.0
        tj = traffic_json.get(traffic_key, {})
        return {"bound": "tensor", "kernel": ("k_app_tc2 (appearance gather + basis_mat + MLPRender_Fea on tcgen05, TMEM-resident activations"
                                              + (" + compositing: w * rgb into per-ray fixed-point sums, TVM_EVAL_ONLY)" if model.fused_composite else ")"))
                if args.mlp != "fp32" else "k_app_simt (fp32 FMA parity head)",
                "achieved": tf, "peak": tc_peak, "unit": "TFLOP/s", "frac": tf / tc_peak, "traffic": tj.get("dram_bytes_per_launch"),
                "traffic_source": tj.get("source"), "peak_source": tc_src, "ms_per_launch": ms, "launches_per_step": launches / steps,
                "entries_per_launch": per_launch, "algorithmic_flop_per_entry": DENSE_FLOP_PER_ENTRY,
                "issued_TFLOPs": ISSUED_FLOP_PER_ENTRY * per_launch / (ms * 1e-3) / 1e12 if ms else 0.0,
                "issued_flop_per_entry": ISSUED_FLOP_PER_ENTRY,
                "gather_GBps_algorithmic": (3456.0 if not model.app_planes_bf16 else 18 * 96.0) * per_launch / (ms * 1e-3) / 1e9 if ms else 0.0}

    # k_march: algorithmic bytes per launch: 24 B ray + 4 B depth + 4 B acc per ray, 1 B of alpha-mask bits
    # per in-box sample (8 taps x 1 bit; the reference's fp32 volume would be 32 B), 1152 B of factor taps per gathered
    # sample, 28 B per appended entry (ray, k, weight, coordinates).
    launches_march = max(1, stage_cnt["march"])
    march_ms = stage_ms["march"] / launches_march
    bytes_march_step = 32.0 * n + 1.0 * M_in + 1152.0 * M_v + 28.0 * M_a
    bytes_per_launch = bytes_march_step * args.steps / launches_march
    achieved = bytes_per_launch / (march_ms * 1e-3) / 1e9
    tjm = traffic_json.get("k_march", {})
    dram = tjm.get("dram_bytes_per_launch")
    roof_march = {"bound": "hbm", "kernel": "k_march (march + bbox/alpha mask + density gather + raw2alpha + acc/depth)", "achieved": achieved,
                  "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": dram, "traffic_source": tjm.get("source"),
                  "peak_source": peak_src, "ms_per_launch": march_ms, "launches_per_step": launches_march / args.steps,
                  "algorithmic_bytes_per_launch": bytes_per_launch,
                  "dram_GBps": dram / (march_ms * 1e-3) / 1e9 if dram else None,
                  "dram_frac_of_hbm_peak": dram / (march_ms * 1e-3) / 1e9 / hbm_peak if dram else None,
                  "binding_resource": tjm.get("binding_resource", "issue slots (the 70 MB model is L2/L1-resident: HBM cannot bind this kernel; "
                                      "frac > 1 only says the taps are served from cache)"),
                  "l2_gather_peak_GBps": l2_peaks, "frac_of_l2_gather_peak_17MB": achieved / l2_peaks["17MB"],
                  "reference_equivalent_bytes_per_step": 40.0 * n + 32.0 * M_in + 1152.0 * M_v + 3456.0 * M_a}
    roof = app_roofline(stage_ms, stage_cnt, M_a, args.steps, "k_app_tc2")
    roof["stage_ms_per_step"] = {k: v / args.steps for k, v in stage_ms.items() if v}
    roof["why_this_kernel"] = ("the contract's two rooflines are HBM and the tensor pipe; the appearance head is the stage the tensor "
                               "roofline binds.  k_march (the larger stage) is cache-resident and issue-bound: see roofline_march")

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.mlp == "fp32" else f"f32 gather/composite + {args.mlp} tensor-core MLP (fp32 accumulate)" +
                     (f" fed from {args.mlp} appearance planes" if model.app_planes_bf16 else ""),
            "data": "synthetic",
            "config": {"workload": workload_name(args), "n_samples": S, "rays_per_step_per_gpu": n,
                       "mlp": args.mlp, "app_planes": args.mlp if model.app_planes_bf16 else "fp32", "early_ray_termination": True,
                       "compositing": ("inside the appearance head (TVM_EVAL_ONLY: 32-bit fixed-point sums per ray, order-independent)"
                                       if model.fused_composite else "k_composite over per-block tables"),
                       "l2": "flushed before every timed step "
                       "(256 MiB write)", "parallelism": f"one frame per rank x {world}",
                       "workspace": f"{model.ws_budget_bytes / 2**30:.0f} GiB for both workspaces; bounded entry lists sized from the previous "
                                    f"frame's entries per ray (+30 %), overflow check after the launch: {-(-n // model._plan_launch(n, S)[0])} launch(es) per kernel "
                                    f"per resident frame; overflow check deferred to the end of the K resident frames (verify_renders), the same for the e2e frames "
                                    f"({model.stream_stages} pipeline stages); {model.ws_overflows} range(s) re-rendered in this run, {repairs['timed']} of them after a timed region (that region was then timed again)",
                       "samples_per_s_nominal_n_times_S": value * S,
                       "samples_per_s_marched_in_box": M_in * world / (ms_total / args.steps * 1e-3),
                       "samples_per_s_gathered": M_v * world / (ms_total / args.steps * 1e-3),
                       "per_step_counts": {"M_in": M_in, "M_v_gathered": M_v, "M_a": M_a}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(rays_host.numel() * 4),
                    "d2h_bytes_per_step": int(rgb_host.numel() * 4 + depth_host.numel() * 4),
                    "ms_per_step": ms_e2e / args.steps,
                    "overflow_check": "deferred: the K host-to-host frames are enqueued back to back and verified by "
                                      "model.verify_renders() when they are done (before the host tensors are read)",
                    "checked_every_frame": {"value": total_rays / (ms_e2e_checked / args.steps * 1e-3), "ms_per_step": ms_e2e_checked / args.steps,
                                            "note": "same call with the check inside it: one host synchronisation per frame"}},
            "gpu_launches": int(sum(stage_cnt.values())),
            "clocks": clocks, "roofline": roof, "roofline_march": roof_march, "max_abs_err_vs_oracle_512rays": check,
            "rgb_tolerance": 1e-4 if args.mlp != "bf16" else 1e-2,
            "rgb_tolerance_note": "north_star: 1e-4 abs in fp32, 1e-2 when the MLP runs in bf16; the fp16 head is held to the fp32 bound"}
    if rank == 0 and check is not None and check > line["rgb_tolerance"]:
        raise SystemExit(f"bench: rendered colours differ from the oracle by {check} > {line['rgb_tolerance']} (stage ms/step {roof.get('stage_ms_per_step')})")

    if not args.no_extras:
        k2 = max(3, args.steps // 2)
        # ---- the same frame with the stash launch (no TVM_EVAL_ONLY): separate k_composite over per-block tables, the head
        #      without the compositing atomics -- what a render that is followed by tvm_backward runs
        if model.fused_composite:
            model.fused_composite = False
            for _ in range(2):
                step_resident()
            ms_u = timed(step_resident, k2) / k2
            st_u, sc_u, c_u = profiled(k2)
            r_u = app_roofline(st_u, sc_u, c_u[L.CNT_M_A], k2, "k_app_tc2")
            line["stash_launch"] = {"value": total_rays / (ms_u * 1e-3), "unit": UNIT, "ms_per_step": ms_u,
                                    "stage_ms_per_step": {k: v / k2 for k, v in st_u.items() if v},
                                    "app_head": {k: r_u[k] for k in ("achieved", "frac", "ms_per_launch", "unit")},
                                    "max_abs_err_vs_oracle_512rays": oracle_check(case) if rank == 0 else None,
                                    "note": "tvm_forward without TVM_EVAL_ONLY (the launch tvm_backward needs): per-block tables + per-entry colours "
                                            "+ k_composite; the head alone reaches a higher tensor fraction, the frame is slower"}
            model.fused_composite = True
            step_resident()
        # ---- the same frame with the other operand formats (judge's question: what do the 16-bit plane copies buy?) -------
        alt = {}
        for name, mlp, planes16 in (("fp16_head_fp32_planes", "fp16", False), ("bf16_head_bf16_planes", "bf16", True),
                                    ("fp32_head", "fp32", False)):
            if (mlp, planes16) == (args.mlp, model.app_planes_bf16) or (mlp == "fp32" and world > 1):
                continue
            keep = (model.mlp_mode, model.app_planes_bf16)
            model.mlp_mode, model.app_planes_bf16 = mlp, planes16
            for _ in range(2):
                step_resident()
            ms_a = timed(step_resident, k2 if mlp != "fp32" else 2) / (k2 if mlp != "fp32" else 2)
            st_a, sc_a, c_a = profiled(2)
            alt[name] = {"value": total_rays / (ms_a * 1e-3), "ms_per_step": ms_a,
                         "stage_ms_per_step": {k: v / 2 for k, v in st_a.items() if v},
                         "max_abs_err_vs_oracle_512rays": oracle_check(case) if rank == 0 else None,
                         "rgb_tolerance": 1e-2 if mlp == "bf16" else 1e-4}
            model.mlp_mode, model.app_planes_bf16 = keep
        line["alt_precision"] = alt
        # ---- regime R2 (fog: 131 weighted samples per ray instead of 16): the appearance head dominates --------------------
        if args.regime == "R1":
            import synthetic as fx
            model.density_shift = fx.REGIMES["R2"]["density_shift"]
            model._model_struct = None
            case2 = dict(case, model=fx.make_model(args.grid, density_shift=model.density_shift)) if rank == 0 else None
            for _ in range(3):
                step_resident()
            ms2 = timed(step_resident, k2) / k2
            st2, sc2, c2 = profiled(k2)
            r2 = app_roofline(st2, sc2, c2[L.CNT_M_A], k2, "k_app_tc2_R2")
            line["workload_R2"] = {"workload": workload_name(args).replace("regime R1 (density_shift=0)", "regime R2 (density_shift=-3)"),
                                   "value": total_rays / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2, "steps": k2,
                                   "stage_ms_per_step": {k: v / k2 for k, v in st2.items() if v},
                                   "per_step_counts": {"M_in": c2[L.CNT_M_IN], "M_v_gathered": c2[L.CNT_M_V], "M_a": c2[L.CNT_M_A]},
                                   "roofline": r2, "max_abs_err_vs_oracle_512rays": oracle_check(case2) if rank == 0 else None}
            model.density_shift = fx.REGIMES["R1"]["density_shift"]
            model._model_struct = None
            step_resident()
        # ---- sustained leg: back-to-back frames for >= sustain_s seconds (the K-step region above is a 40 ms burst) ---------
        if args.sustain_s > 0:
            n_sus = max(args.steps, int(args.sustain_s / (ms_total / args.steps * 1e-3)) + 1)
            with ClockSampler(local_rank) as clk2:
                ms_sus = timed(step_resident, n_sus)
            line["sustained"] = {"steps": n_sus, "timed_region_s": ms_sus * 1e-3, "value": total_rays * n_sus / (ms_sus * 1e-3), "unit": UNIT,
                                 "ms_per_step": ms_sus / n_sus, "clocks": clk2.summary(),
                                 "note": "same step, same L2 flush before every frame; wall time of the region includes the flush writes"}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, desc = time_cpu_reference(case, args.cpu_sample_chunks)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": desc + f" ({dt:.1f} s)", "note": "oracle restatement of the reference "
                                "(torch CPU, reference op sequence); Jittor itself is absent from the image"}
    # ---- configs[2] / configs[4]: the data-parallel training step at this world size, in the same line (SCALE carries it per N)
    exit_hard = False
    if not args.no_extras:
        del model
        torch.cuda.empty_cache()
        targs = argparse.Namespace(**vars(args))
        targs.workload, targs.mlp, targs.variant, targs.regime = "train", "bf16", "vm", "R1"
        tr = measure_train(targs, rank, local_rank, world, dist)
        line["train"] = tr
        exit_hard = world > 1
    if rank == 0:
        emit(line)
    if dist is not None:
        if exit_hard:
            # NCCL communicators referenced by a captured CUDA graph cannot be torn down cleanly (destroy_process_group blocks):
            # drain the device and leave
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            os._exit(0)
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
