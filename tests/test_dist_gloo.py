"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: ray sharding + result gather, and the
flat-buffer gradient all-reduce.  The compute inside each rank is the ORACLE (no GPU here); what is
tested is that (a) sharded render == unsharded render bit for bit, (b) averaging the ranks' gradients of
their batch slices reproduces the full-batch gradient, (c) the sampler slices one global permutation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    torch.set_num_threads(2)
    import importlib.util
    spec = importlib.util.spec_from_file_location("tvm_dist", os.path.join(root, "jittor-myc-nerfs_b200", "dist.py"))
    D = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(D)
    from oracle import fixtures as fx, tensorf_oracle as orc
    r, w = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    case = fx.make_case(24, 101, "R2", mask_res=24, train=True, cd=4, ca=8, app_dim=6)
    model = orc.OracleTensorVMSplit(case["model"], case["alpha_volume"], case["alpha_aabb"])
    rays = torch.from_numpy(case["rays"])

    def renderer(rs, m, **kw):
        with torch.no_grad():
            return orc.OctreeRender_trilinear_fast(rs, m, chunk=32, **kw)

    rgb, depth = D.render_sharded(rays, model, renderer, gather=True, N_samples=40, white_bg=True, is_train=False)
    # gradient exchange: each rank differentiates the MSE of ITS slice of the global batch
    s, e = D.shard_bounds(rays.shape[0], rank, world)
    sub = dict(case, rays=case["rays"][s:e], jitter=case["jitter"][s:e], target=case["target"][s:e])
    g = orc.backward_case(sub, N_samples=40, dtype=torch.float64)["grads"]
    flat = torch.cat([torch.from_numpy(np.ascontiguousarray(v)).reshape(-1) * (e - s) for v in g.values()])
    D.allreduce_flat_(flat, average=False)          # sum of (n_r * mean-gradient_r) ...
    flat /= rays.shape[0]                           # ... / N == full-batch mean gradient
    samp = D.ShardedSampler(1000, 64, rank, world)
    ids = [samp.nextids() for _ in range(3)]
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), rgb=rgb.numpy(), depth=depth.numpy(), flat=flat.numpy(),
             ids=np.concatenate(ids))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_render_and_grad_allreduce(tmp_path):
    world = 2
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    from oracle import fixtures as fx, tensorf_oracle as orc
    case = fx.make_case(24, 101, "R2", mask_res=24, train=True, cd=4, ca=8, app_dim=6)
    model = orc.OracleTensorVMSplit(case["model"], case["alpha_volume"], case["alpha_aabb"])
    with torch.no_grad():
        rgb, _, depth, _, _ = orc.OctreeRender_trilinear_fast(torch.from_numpy(case["rays"]), model, chunk=101,
                                                              N_samples=40, white_bg=True, is_train=False)
    full = orc.backward_case(case, N_samples=40, dtype=torch.float64)["grads"]
    flat_ref = np.concatenate([np.ascontiguousarray(v).reshape(-1) for v in full.values()])
    outs = [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]
    for o in outs:
        assert np.array_equal(o["rgb"], rgb.numpy()) and np.array_equal(o["depth"], depth.numpy())
        assert np.allclose(o["flat"], flat_ref, rtol=1e-9, atol=1e-12)
    # the two ranks' id slices tile each global batch of ONE shared permutation
    ids = np.concatenate([outs[0]["ids"].reshape(3, -1), outs[1]["ids"].reshape(3, -1)], 1)
    assert len(np.unique(ids)) == ids.size


def test_shard_bounds_cover():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("tvm_dist", os.path.join(root, "jittor-myc-nerfs_b200", "dist.py"))
    D = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(D)
    for n in (0, 1, 7, 640000, 640001):
        for w in (1, 2, 3, 8):
            b = [D.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1
