#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU; tests/test_gpu_dist.py launches it when >= 2 devices are visible):
(a) ray-sharded render == single-GPU render, bit for bit;
(b) data-parallel training step, gradients exchanged by NCCL all_reduce of the flat buffer == oracle full-batch gradient;
(c) the same step with libtvmrender's own all-reduce over NVLink peer memory (tvm_allreduce_sum: NVLS multicast when the fabric
    has it, and the peer load/store path forced with TVM_AR_NO_MULTICAST) == oracle, == the NCCL result to fp32 rounding,
    bit-identical across ranks; known-answer test of the collective on ragged sizes and sub-ranges;
(d) the captured training step (TrainStepGraph) with the peer all-reduce inside the graph: the ranks stay in lockstep
    (identical parameters on every rank after 20 replays) and the loss decreases."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import jittor_myc_nerfs_b200 as pkg
import synthetic as fx
from oracle import tensorf_oracle as orc

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
rank, world = pkg.dist.init_from_env("nccl", device=dev)
n_all = 1024
case = fx.make_case(48, n_all, "R2", mask_res=48, train=True)
model = pkg.model_from_params(case["model"], f"cuda:{local}", case["alpha_volume"], case["alpha_aabb"])
rays = torch.from_numpy(case["rays"]).to(dev)
with torch.no_grad():
    full_rgb, _, full_depth, _, _ = pkg.OctreeRender_trilinear_fast(rays, model, N_samples=167, is_train=False)
    rgb, depth = pkg.dist.render_sharded(rays, model, pkg.OctreeRender_trilinear_fast, gather=True, N_samples=167,
                                         is_train=False, white_bg=True)
assert torch.equal(rgb, full_rgb) and torch.equal(depth, full_depth), "sharded render differs"

# DP step: each rank differentiates the MSE of its slice; grads are averaged by ONE flat all-reduce
s, e = pkg.dist.shard_bounds(n_all, rank, world)     # equal shard sizes => mean of per-rank means == global mean
jit = torch.from_numpy(case["jitter"]).to(dev)
tgt = torch.from_numpy(case["target"]).to(dev)
ref = orc.backward_case(case, N_samples=167)
names = lambda m: (("density_plane.0", m.density_plane[0]), ("app_line.2", m.app_line[2]), ("basis_mat.weight", m.basis_mat.weight),
                   ("renderModule.mlp.0.weight", m.renderModule.mlp[0].weight))


def dp_grads(m):
    for p in m.parameters():
        p.grad = None
    out, _ = m(rays[s:e], is_train=True, N_samples=167, jitter=jit[s:e])
    torch.mean((out - tgt[s:e]) ** 2).backward()
    torch.cuda.synchronize()
    worst = 0.0
    for name, p in names(m):
        g, r = p.grad.cpu().numpy().astype(np.float64), ref["grads"][name]
        worst = max(worst, np.abs(g - r).max() / np.abs(r).max())
    return worst, [p.grad.detach().clone() for p in m.parameters()]


model.grad_sync = world > 1
worst_nccl, g_nccl = dp_grads(model)
assert worst_nccl <= 1e-4, worst_nccl
msg = f"NCCL {worst_nccl:.2e}"
if world > 1:
    # (c) known-answer test of the collective itself, both paths
    for no_mc in ("", "1"):
        if no_mc:
            os.environ["TVM_AR_NO_MULTICAST"] = "1"
        else:
            os.environ.pop("TVM_AR_NO_MULTICAST", None)
        comm = pkg.dist.PeerComm(1 << 20, dev, n_ctas=32)
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        for off, cnt in ((0, 1 << 20), (4096, 777 * 4), (12, 4), (0, 8 * world)):
            x = torch.randint(-64, 65, (1 << 20,), generator=g, device=dev).float()       # exact in fp32: any sum order agrees
            comm.buf.copy_(x)
            want = x.clone()
            dist.all_reduce(want[off:off + cnt])
            torch.cuda.synchronize(); dist.barrier()
            comm.allreduce_(off, cnt)
            torch.cuda.synchronize()
            assert torch.equal(comm.buf, want), (no_mc, off, cnt, float((comm.buf - want).abs().max()))
            dist.barrier()
        kind = "p2p" if not comm.multicast else "nvls"
        # the training step through it
        m2 = pkg.model_from_params(case["model"], f"cuda:{local}", case["alpha_volume"], case["alpha_aabb"])
        m2.enable_peer_allreduce(n_ctas=32)
        worst, g2 = dp_grads(m2)
        assert worst <= 1e-4, (kind, worst)
        for a, b in zip(g2, g_nccl):
            assert float((a - b).abs().max()) <= 1e-6 * max(1.0, float(b.abs().max())) + 1e-9, kind
        # every rank holds the same bits
        flat = torch.cat([t.reshape(-1) for t in g2])
        ref0 = flat.clone()
        dist.broadcast(ref0, 0)
        assert torch.equal(flat, ref0), f"{kind}: gradients differ between ranks"
        msg += f", peer[{kind}] {worst:.2e}"
    os.environ.pop("TVM_AR_NO_MULTICAST", None)
    # (d) captured step with the peer all-reduce inside the graph
    m3 = pkg.model_from_params(case["model"], f"cuda:{local}", case["alpha_volume"], case["alpha_aabb"], "bf16")
    m3.enable_peer_allreduce()
    opt = pkg.Adam(m3.get_optparam_groups(0.02, 0.001), betas=(0.9, 0.99))
    gs = pkg.TrainStepGraph(m3, opt, e - s, 167, white_bg=True, TV_weight_density=1.0, TV_weight_app=1.0)
    jits = torch.from_numpy(fx.jitter(e - s, seed=5)).to(dev)
    l0 = float(gs.step(rays[s:e], tgt[s:e], jitter=jits))
    for _ in range(19):
        l1 = gs.step(rays[s:e], tgt[s:e], jitter=jits)
    l1 = float(l1)
    flat = torch.cat([p.detach().reshape(-1) for p in m3.parameters()])
    ref0 = flat.clone()
    dist.broadcast(ref0, 0)
    assert torch.equal(flat, ref0), "captured DP step: parameters diverged between ranks"
    lt = torch.tensor([l0, l1], device=dev)
    dist.all_reduce(lt)
    assert float(lt[1]) < float(lt[0]), (float(lt[0]), float(lt[1]))
    msg += f", graph loss {float(lt[0]) / world:.4f} -> {float(lt[1]) / world:.4f}"
if rank == 0:
    print(f"dp_check OK: world={world}, sharded render bit-identical, DP grad rel err vs oracle: {msg}")
if world > 1:
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)      # communicators referenced by captured graphs / symmetric memory: leave without a teardown
