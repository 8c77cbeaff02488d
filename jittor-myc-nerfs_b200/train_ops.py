"""Training-step remainder: the regularisers of train.py:233-251 and the optimiser of train.py:187,
with the reference's names and call shapes; the arithmetic runs in libtvmrender.so (csrc/tvm_train.cu).

  TVLoss()                      utils.py:123-142  (callable on one NCHW plane; passed to TV_loss_density / TV_loss_app)
  RegularizerMixin              TensorVMSplit.TV_loss_density / TV_loss_app / density_L1 / vector_comp_diffs (tensoRF.py:177-207)
  Adam(grad_vars, lr, betas)    jt.optim.Adam as used at train.py:187: zero_grad() / backward(loss) / step(), param_groups[i]['lr']

Every regulariser is ONE fused value+gradient sweep per tensor: the autograd node keeps the gradient it already computed
and scales it by the incoming scalar in backward.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch

from . import _lib as L


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _launch(kind, x, weight, loss, grad, weight_dev=None):
    """One fused value+gradient sweep: *loss += w f(x), grad += w df/dx with w = weight (* *weight_dev)."""
    lib, st = L.load(), _stream_ptr()
    if kind == "tv":
        _, Cc, H, W = x.shape
        L.check(lib.tvm_tv_loss(_ptr(x), Cc, H, W, float(weight), _ptr(weight_dev), _ptr(loss), _ptr(grad), st), "tvm_tv_loss")
    elif kind == "l1":
        L.check(lib.tvm_l1_loss(_ptr(x), x.numel(), float(weight), _ptr(weight_dev), _ptr(loss), _ptr(grad), st), "tvm_l1_loss")
    else:
        _, Cc, Ln, _ = x.shape
        L.check(lib.tvm_vector_diffs(_ptr(x), Cc, Ln, float(weight), _ptr(weight_dev), _ptr(loss), _ptr(grad), st),
                "tvm_vector_diffs")


class _RegFn(torch.autograd.Function):
    """loss = sum_i weight_i * f_kind(x_i); gradients are produced by the same kernels that produce the value."""

    @staticmethod
    def forward(ctx, kind, weights, *tensors):
        L.require_cuda()
        loss = torch.zeros((), dtype=torch.float32, device=tensors[0].device)
        need = [t.requires_grad for t in tensors]
        grads = []
        for t, w, n in zip(tensors, weights, need):
            x = t.detach().contiguous()
            g = torch.zeros_like(x) if n else None
            _launch(kind, x, w, loss, g)
            grads.append(g)
        ctx.grads = grads
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        return (None, None, *[None if g is None else g * d_loss for g in ctx.grads])


class TVLoss(torch.nn.Module):
    """utils.TVLoss (utils.py:123-142)."""

    def __init__(self, TVLoss_weight=1):
        super().__init__()
        self.TVLoss_weight = TVLoss_weight

    def forward(self, x):
        if x.shape[0] != 1:
            raise NotImplementedError("TVLoss is applied to the [1,C,H,W] factor planes")
        return _RegFn.apply("tv", [self.TVLoss_weight], x)

    execute = forward


class RegularizerMixin:
    """Mixed into TensorVMSplit (tensorf.py)."""

    def vectorDiffs(self, vector_comps):
        return _RegFn.apply("ortho", [1.0] * len(vector_comps), *vector_comps)

    def vector_comp_diffs(self):
        return self.vectorDiffs(list(self.density_line)) + self.vectorDiffs(list(self.app_line))

    def density_L1(self):
        ts = [*self.density_plane, *self.density_line]
        return _RegFn.apply("l1", [1.0] * len(ts), *ts)

    def TV_loss_density(self, reg):
        return _RegFn.apply("tv", [reg.TVLoss_weight * 1e-2] * 3, *self.density_plane)

    def TV_loss_app(self, reg):
        return _RegFn.apply("tv", [reg.TVLoss_weight * 1e-2] * 3, *self.app_plane)


class Adam:
    """jt.optim.Adam(grad_vars, lr=..., betas=(0.9, 0.99)) (train.py:187) over libtvmrender's multi-tensor kernel:
    one launch per 32 parameter tensors instead of four elementwise passes per tensor."""

    def __init__(self, params, lr=0.001, eps=1e-8, betas=(0.9, 0.999), weight_decay=0):
        if weight_decay:
            raise NotImplementedError("weight_decay is not used by the reference's training loop")
        if isinstance(params, (list, tuple)) and params and isinstance(params[0], dict):
            self.param_groups = [dict(g, params=list(g["params"])) for g in params]
        else:
            self.param_groups = [{"params": list(params)}]
        for g in self.param_groups:
            g.setdefault("lr", lr)
        self.lr, self.eps, self.betas = lr, eps, betas
        self.n_step = 0
        self.state = {}

    def zero_grad(self):
        for g in self.param_groups:
            for p in g["params"]:
                p.grad = None

    def backward(self, loss):
        loss.backward()

    def _state(self, p):
        s = self.state.get(id(p))
        if s is None or s[0].shape != p.shape:
            s = (torch.zeros_like(p.data), torch.zeros_like(p.data))
            self.state[id(p)] = s
        return s

    # -- hyper-parameters live in a small device buffer refreshed from pinned host memory, so that a captured CUDA
    #    graph (TrainStepGraph) can be replayed with the next step's bias correction / decayed learning rates
    def _prepare_hyper(self):
        """Host side of one step: advance the step counter and stage {c_step, lr per group} in pinned memory."""
        self.n_step += 1
        n = float(self.n_step)
        if getattr(self, "_hyper_host", None) is None or self._hyper_host.numel() != 1 + len(self.param_groups):
            self._hyper_host = torch.zeros(1 + len(self.param_groups), dtype=torch.float32).pin_memory()
            self._hyper_dev = None
        b0, b1 = self.betas
        self._hyper_host[0] = math.sqrt(1.0 - b1 ** n) / (1.0 - b0 ** n)
        for i, g in enumerate(self.param_groups):
            self._hyper_host[1 + i] = float(g["lr"])

    @torch.no_grad()
    def _launch(self):
        """Device side of one step (capturable): refresh the hyper buffer, one multi-tensor kernel per 32 tensors."""
        entries, keep = [], []
        dev = None
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.data.is_contiguous()):
                    raise L.TvmError("Adam: parameters must be contiguous fp32 CUDA tensors (no CPU fallback)")
                dev = p.device
                m, v = self._state(p)
                gr = p.grad.contiguous()
                keep += [gr, p]
                e = L.TvmAdamTensor()
                e.p, e.g, e.m, e.v, e.n = p.data.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()
                e.lr, e.lr_index = float(g["lr"]), gi
                entries.append(e)
        if not entries:
            return
        if self._hyper_dev is None:
            self._hyper_dev = torch.zeros_like(self._hyper_host, device=dev)
        self._hyper_dev.copy_(self._hyper_host, non_blocking=True)
        arr = (L.TvmAdamTensor * len(entries))(*entries)
        L.check(L.load().tvm_adam_step(arr, len(entries), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                       int(self.n_step), _ptr(self._hyper_dev), _stream_ptr()), "tvm_adam_step")
        # the kernel wrote through raw pointers: bump the version counters so that the packed device image is rebuilt
        for t in keep:
            if isinstance(t, torch.nn.Parameter):
                torch.autograd.graph.increment_version(t)

    def step(self, loss=None):
        if loss is not None:
            self.zero_grad()
            loss.backward()
        self._prepare_hyper()
        self._launch()


class TrainStepGraph:
    """One optimisation step of train.py:218-261 (sample jitter, render, MSE, backward, regularisers, Adam, re-pack of the
    updated grids) captured ONCE into a CUDA graph and replayed per iteration: the step is launch-bound (~60 kernels of
    10-100 us), so the graph removes the host time between them.  The captured body calls the library directly
    (tvm_forward, tvm_mse_loss, tvm_backward, the value+gradient regulariser sweeps, tvm_adam_step) -- no autograd engine,
    hence no second thread touching CUDA during capture.  Per-step scalars (Adam bias correction, decayed learning
    rates, regulariser weights) reach the graph through a pinned-host -> device copy node.

        g = TrainStepGraph(model, opt, n_rays=4096, N_samples=S, white_bg=True, TV_weight_density=2.0, TV_weight_app=2.0)
        loss = g.step(rays, rgbs)        # device scalar; call g.set_weights(...) / change opt.param_groups[i]['lr'] freely
    """

    def __init__(self, model, optimizer, n_rays, N_samples, white_bg=True, TV_weight_density=0.0, TV_weight_app=0.0,
                 L1_reg_weight=0.0, Ortho_reg_weight=0.0, normal_vector_penalty_weight=0.0):
        self.model, self.opt = model, optimizer
        dev = model.device
        self.rays = torch.zeros((n_rays, 6), dtype=torch.float32, device=dev)
        self.target = torch.zeros((n_rays, 3), dtype=torch.float32, device=dev)
        self.S, self.white_bg = int(N_samples), bool(white_bg)
        self.use = dict(tv_d=TV_weight_density > 0, tv_a=TV_weight_app > 0, l1=L1_reg_weight > 0, ortho=Ortho_reg_weight > 0,
                        pen=normal_vector_penalty_weight > 0)
        self._w_host = torch.tensor([TV_weight_density, TV_weight_app, L1_reg_weight, Ortho_reg_weight,
                                     normal_vector_penalty_weight], dtype=torch.float32).pin_memory()
        self._w_dev = self._w_host.to(dev)
        self._tv = TVLoss()
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)          # MSE of the step (train.py:228)
        self.reg_loss = torch.zeros(1, dtype=torch.float32, device=dev)      # sum of the weighted regularisers
        self.jitter = None                  # optional static [n] buffer (step(..., jitter=...)); default: torch.rand per step
        self.graph = None

    def set_weights(self, TV_weight_density=None, TV_weight_app=None, L1_reg_weight=None, Ortho_reg_weight=None,
                    normal_vector_penalty_weight=None):
        for i, v in enumerate((TV_weight_density, TV_weight_app, L1_reg_weight, Ortho_reg_weight, normal_vector_penalty_weight)):
            if v is not None:
                self._w_host[i] = float(v)

    def _body(self):
        """The step without the autograd engine: every kernel is enqueued by this thread on the capturing stream."""
        m, lib = self.model, L.load()
        self._w_dev.copy_(self._w_host, non_blocking=True)
        n = self.rays.shape[0]
        jitter = self.jitter if self.jitter is not None else torch.rand(n, dtype=torch.float32, device=self.rays.device)
        flags = m._flags(self.white_bg)
        rgb, _ = m._forward_raw(self.rays, jitter, flags, self.S)
        d_rgb = torch.empty_like(rgb)
        L.check(lib.tvm_mse_loss(_ptr(rgb), _ptr(self.target), n, 1.0, _ptr(self.loss), _ptr(d_rgb), _stream_ptr()),
                "tvm_mse_loss")                                                       # train.py:228
        d_pen = self._w_dev[4:5] if (self.use["pen"] and m.VARIANT == L.VARIANT_REF) else None
        grads = m._backward_raw(self.rays, jitter, flags, self.S, rgb, d_rgb, d_pen)
        params = m._param_list()
        for p, g in zip(params, grads):
            p.grad = g
        # regularisers (train.py:233-251): value + gradient sweeps straight into .grad, weights read from the device
        reg = self.reg_loss
        reg.zero_()
        w = self._w_dev
        tv_jobs = []
        for on, planes, wd in ((self.use["tv_d"], m.density_plane, w[0:1]), (self.use["tv_a"], m.app_plane, w[1:2])):
            if on:
                for p in planes:
                    _, Cc, H, W = p.shape
                    tv_jobs.append(L.TvmTvJob(p.data_ptr(), p.grad.data_ptr(), Cc, H, W, 1e-2, wd.data_ptr()))
        if tv_jobs:        # all TV sweeps of the step in one launch
            L.check(lib.tvm_tv_loss_batch((L.TvmTvJob * len(tv_jobs))(*tv_jobs), len(tv_jobs), _ptr(reg), _stream_ptr()),
                    "tvm_tv_loss_batch")
        if self.use["l1"]:
            for p in [*m.density_plane, *m.density_line]:
                _launch("l1", p.detach(), 1.0, reg, p.grad, w[2:3])
        if self.use["ortho"]:
            for p in [*m.density_line, *m.app_line]:
                _launch("ortho", p.detach(), 1.0, reg, p.grad, w[3:4])
        self.opt._launch()
        m._pack(force=True)                       # the next forward (and any render in between) sees the updated grids

    def capture(self):
        """Warm up on a side stream (torch's capture protocol), then record the step.  The warm-up steps are real
        optimisation steps on whatever the static buffers hold: parameters and optimiser state are snapshotted and restored."""
        m, opt = self.model, self.opt
        params = [p for g in opt.param_groups for p in g["params"]]
        snap = [p.detach().clone() for p in params]
        n_step0 = opt.n_step
        s = torch.cuda.Stream(device=m.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                opt._prepare_hyper()
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        opt._prepare_hyper()
        with torch.cuda.graph(self.graph, capture_error_mode=os.environ.get("TVM_GRAPH_CAPTURE_MODE", "global")):
            self._body()
        with torch.no_grad():
            for p, q in zip(params, snap):
                p.copy_(q)
            for mv in opt.state.values():
                mv[0].zero_()
                mv[1].zero_()
        opt.n_step = n_step0
        m._pack(force=True)

    def step(self, rays, target, jitter=None):
        if jitter is not None and self.jitter is None:
            if self.graph is not None:
                raise RuntimeError("pass jitter from the first step on (the graph was captured with on-device random jitter)")
            self.jitter = torch.zeros(self.rays.shape[0], dtype=torch.float32, device=self.rays.device)
        if self.graph is None:
            self.capture()
        if jitter is not None:
            self.jitter.copy_(jitter, non_blocking=True)
        self.rays.copy_(rays, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        self.opt._prepare_hyper()
        self.graph.replay()
        return self.loss
