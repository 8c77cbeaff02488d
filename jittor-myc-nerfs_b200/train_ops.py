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


class _PinnedRing:
    """Host staging for small per-step scalars that a stream-ordered H2D copy reads LATER than the host writes them.
    Each push() takes the next of `slots` pinned buffers; a slot is reused only after the copy that read it has executed
    (an event recorded behind that copy), so the host may run any number of steps ahead of the device without a later
    step's values reaching an earlier step's kernels."""

    def __init__(self, n, slots=8):
        self.bufs = [torch.zeros(n, dtype=torch.float32).pin_memory() for _ in range(slots)]
        self.views = [b.numpy() for b in self.bufs]          # host writes go through numpy (a tensor __setitem__ costs ~3 us each)
        self.events = [None] * slots
        self.i = 0

    def upload(self, values, dst):
        """dst (device, [n]) <- values, enqueued on the current stream."""
        k = self.i
        self.i = (k + 1) % len(self.bufs)
        if self.events[k] is not None:
            self.events[k].synchronize()          # only ever waits when the host is `slots` steps ahead
        self.views[k][:len(values)] = values
        dst.copy_(self.bufs[k], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[k] = ev


class _RegFn(torch.autograd.Function):
    """loss = sum_i weight_i * f_kind(x_i); gradients are produced by the same kernels that produce the value."""

    @staticmethod
    def forward(ctx, kind, weights, *tensors):
        L.require_cuda()
        loss = torch.zeros((), dtype=torch.float32, device=tensors[0].device)
        need = [t.requires_grad for t in tensors]
        grads = []
        if kind == "tv" and len(tensors) <= 8:
            # all planes of the call in ONE launch, gradients written (not accumulated) into fresh buffers: per element one read
            # of x and one write of the gradient
            lib, jobs, keep = L.load(), [], []
            for t, w, n in zip(tensors, weights, need):
                x = t.detach().contiguous()
                g = torch.empty_like(x) if n else None
                keep.append(x)
                _, Cc, H, W = x.shape
                jobs.append(L.TvmTvJob(x.data_ptr(), g.data_ptr() if g is not None else None, Cc, H, W, float(w), None, 1))
                grads.append(g)
            L.check(lib.tvm_tv_loss_batch((L.TvmTvJob * len(jobs))(*jobs), len(jobs), _ptr(loss), _stream_ptr()), "tvm_tv_loss_batch")
        else:
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
            # a ParameterList is kept BY REFERENCE, as jt.optim does with the list objects of get_optparam_groups
            # (tensoRF.py:168-174): shrink() replaces entries of those lists in place (tensoRF.py:300-301) and the next step
            # must update the cropped grids (fresh moments: _state notices the new tensors); generators are materialised
            self.param_groups = [dict(g, params=g["params"] if isinstance(g["params"], (list, torch.nn.ParameterList))
                                      else list(g["params"])) for g in params]
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
        """Host side of one step: advance the step counter and send {c_step, lr per group} to the device buffer the kernel
        reads, through a ring of pinned slots (stream-ordered; safe however far the host runs ahead of the device).
        Runs OUTSIDE a captured graph: the graph only ever reads the device buffer."""
        self.n_step += 1
        n = float(self.n_step)
        k = 1 + len(self.param_groups)
        device = next(p.device for g in self.param_groups for p in g["params"])
        if getattr(self, "_hyper_ring", None) is None or self._hyper_dev.numel() != k or self._hyper_dev.device != device:
            self._hyper_ring = _PinnedRing(k)
            self._hyper_dev = torch.zeros(k, dtype=torch.float32, device=device)
        b0, b1 = self.betas
        self._hyper_ring.upload([math.sqrt(1.0 - b1 ** n) / (1.0 - b0 ** n)] + [float(g["lr"]) for g in self.param_groups],
                                self._hyper_dev)

    @torch.no_grad()
    def _launch(self, only=None):
        """Device side of one step (capturable): one multi-tensor kernel per 32 tensors, hyper-parameters from the device
        buffer _prepare_hyper filled.  `only`: restrict the launch to these parameters (the pipelined data-parallel step
        updates the appearance half while the density half of the gradient is still being exchanged)."""
        entries, keep = [], []
        dev = None
        only = None if only is None else {id(p) for p in only}
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                if p.grad is None or (only is not None and id(p) not in only):
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
        self._w = [float(TV_weight_density), float(TV_weight_app), float(L1_reg_weight), float(Ortho_reg_weight),
                   float(normal_vector_penalty_weight)]
        self._w_ring = _PinnedRing(len(self._w))
        self._w_dev = torch.tensor(self._w, dtype=torch.float32, device=dev)
        self._signature = None
        self._tv = TVLoss()
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)          # MSE of the step (train.py:228)
        self.reg_loss = torch.zeros(1, dtype=torch.float32, device=dev)      # sum of the weighted regularisers
        self.jitter = None                  # optional static [n] buffer (step(..., jitter=...)); default: torch.rand per step
        self.graph = None
        self.npp = hasattr(model, "bg_net")         # NerfPlusPlus: tvm_forward_npp / tvm_backward_npp, background net in the optimiser

    def set_weights(self, TV_weight_density=None, TV_weight_app=None, L1_reg_weight=None, Ortho_reg_weight=None,
                    normal_vector_penalty_weight=None):
        for i, v in enumerate((TV_weight_density, TV_weight_app, L1_reg_weight, Ortho_reg_weight, normal_vector_penalty_weight)):
            if v is not None:
                self._w[i] = float(v)

    def _body(self):
        """The step without the autograd engine: every kernel is enqueued by this thread on the capturing stream."""
        m, lib = self.model, L.load()
        n = self.rays.shape[0]
        if self.npp:
            return self._body_npp()
        jitter = self.jitter if self.jitter is not None else torch.rand(n, dtype=torch.float32, device=self.rays.device)
        flags = m._flags(self.white_bg)
        rgb, _ = m._forward_raw(self.rays, jitter, flags, self.S)
        d_rgb = torch.empty_like(rgb)
        L.check(lib.tvm_mse_loss(_ptr(rgb), _ptr(self.target), n, 1.0, _ptr(self.loss), _ptr(d_rgb), _stream_ptr()),
                "tvm_mse_loss")                                                       # train.py:228
        d_pen = self._w_dev[4:5] if (self.use["pen"] and m.VARIANT == L.VARIANT_REF) else None
        reg = self.reg_loss
        reg.zero_()
        w = self._w_dev
        if m.grad_sync and m._peer_comm is not None and os.environ.get("TVM_AR_OVERLAP", "2") == "2":
            return self._body_pipelined(jitter, flags, rgb, d_rgb, d_pen)
        grads = m._backward_raw(self.rays, jitter, flags, self.S, rgb, d_rgb, d_pen)
        self._regularisers_and_update(m._param_list(), grads)
        m._pack(force=True)                       # the next forward (and any render in between) sees the updated grids

    def _regularisers_and_update(self, params, grads):
        """Tail of the serial step: .grad <- gradients, regularisers (train.py:233-251) as value + gradient sweeps straight into
        .grad with device-side weights, multi-tensor Adam."""
        m, lib = self.model, L.load()
        for p, g in zip(params, grads):
            p.grad = g
        reg, w = self.reg_loss, self._w_dev
        tv_jobs = []
        for on, planes, wd in ((self.use["tv_d"], m.density_plane, w[0:1]), (self.use["tv_a"], m.app_plane, w[1:2])):
            if on:
                for p in planes:
                    _, Cc, H, W = p.shape
                    tv_jobs.append(L.TvmTvJob(p.data_ptr(), p.grad.data_ptr(), Cc, H, W, 1e-2, wd.data_ptr()))
        if tv_jobs:
            L.check(lib.tvm_tv_loss_batch((L.TvmTvJob * len(tv_jobs))(*tv_jobs), len(tv_jobs), _ptr(reg), _stream_ptr()),
                    "tvm_tv_loss_batch")
        if self.use["l1"]:
            for p in [*m.density_plane, *m.density_line]:
                _launch("l1", p.detach(), 1.0, reg, p.grad, w[2:3])
        if self.use["ortho"]:
            for p in [*m.density_line, *m.app_line]:
                _launch("ortho", p.detach(), 1.0, reg, p.grad, w[3:4])
        self.opt._launch()

    def _body_npp(self):
        """NerfPlusPlus (configs/Scarf.txt): foreground on black + 512-sample background network; the stratified draws of
        perturb_samples (nerfplusplus.py:196-205) come from the device generator inside the graph."""
        m, lib = self.model, L.load()
        n, dev = self.rays.shape[0], self.rays.device
        fg_rand = torch.rand((n, self.S), dtype=torch.float32, device=dev)
        bg_rand = torch.rand((n, 512), dtype=torch.float32, device=dev)
        flags = m._flags(False)
        rgb, _ = m._forward_npp_raw(self.rays, fg_rand, bg_rand, flags, self.S)
        d_rgb = torch.empty_like(rgb)
        L.check(lib.tvm_mse_loss(_ptr(rgb), _ptr(self.target), n, 1.0, _ptr(self.loss), _ptr(d_rgb), _stream_ptr()), "tvm_mse_loss")
        self.reg_loss.zero_()
        grads = m._backward_npp_raw(self.rays, fg_rand, bg_rand, flags, self.S, rgb, d_rgb)
        self._regularisers_and_update([*m._param_list(), *m._bg_param_list()], grads)
        m._pack(force=True)
        m._bg_struct(force=True)                  # the background network's packed fp32 buffer, fold and bf16 image follow the update

    def _body_pipelined(self, jitter, flags, rgb, d_rgb, d_pen):
        """Tail of the data-parallel step with the gradient exchange hidden behind it (peer all-reduce, tvm_backward_dp phases):
        side stream:  [all-reduce appearance half]            [all-reduce density half]
        this stream:  app backward | density scatter | unpack + TV + Adam of the APPEARANCE half | ... of the DENSITY half | pack
        Every tensor sees exactly the arithmetic of the serial body; only the launch order differs."""
        m, lib = self.model, L.load()
        main = torch.cuda.current_stream()
        gp, items, ev_app, ev_den = m._backward_raw(self.rays, jitter, flags, self.S, rgb, d_rgb, d_pen, pipelined=True)
        reg, w = self.reg_loss, self._w_dev
        for p in m._param_list():
            p.grad = None
        for part, ev, planes, tv_on, wd in (("app", ev_app, m.app_plane, self.use["tv_a"], w[1:2]),
                                            ("density", ev_den, m.density_plane, self.use["tv_d"], w[0:1])):
            main.wait_event(ev)
            filled = m._unpack_part(gp, items, part)
            if tv_on:
                jobs = [L.TvmTvJob(p.data_ptr(), p.grad.data_ptr(), p.shape[1], p.shape[2], p.shape[3], 1e-2, wd.data_ptr()) for p in planes]
                L.check(lib.tvm_tv_loss_batch((L.TvmTvJob * len(jobs))(*jobs), len(jobs), _ptr(reg), _stream_ptr()), "tvm_tv_loss_batch")
            if part == "density" and self.use["l1"]:
                for p in [*m.density_plane, *m.density_line]:
                    _launch("l1", p.detach(), 1.0, reg, p.grad, w[2:3])
            if self.use["ortho"]:
                for p in (m.app_line if part == "app" else m.density_line):
                    _launch("ortho", p.detach(), 1.0, reg, p.grad, w[3:4])
            self.opt._launch(only=filled)
        m._pack(force=True)

    def _upload_scalars(self):
        """Per-step scalars -> device, eagerly and stream-ordered in front of the replay (never from inside the graph: a
        captured pinned-host read would see whatever the host has written by the time the replay runs)."""
        self.opt._prepare_hyper()
        self._w_ring.upload(self._w, self._w_dev)

    def _capture_signature(self):
        """Everything a captured graph holds by raw pointer or by value.  Maintenance calls (updateAlphaMask, shrink,
        upsample_volume_grid, load) and changes of mlp_mode / app_planes_bf16 replace these; step() then re-captures."""
        m = self.model
        am = m.alphaMask
        return (tuple(p.data_ptr() for p in m._param_list()), tuple(int(g) for g in m.gridSize),
                None if am is None else (am.bits.data_ptr(), am.bricks.data_ptr(), am.dilated.data_ptr(), tuple(int(g) for g in am.gridSize)),
                tuple(float(a) for a in m.aabb.reshape(-1)), m.mlp_mode, m.app_planes_bf16, m.early_termination,
                m.empty_space_skipping, bool(m.grad_sync),
                None if m._packed is None else m._packed.data_ptr(),
                None if getattr(m, "_grads_packed", None) is None else m._grads_packed.data_ptr(),
                None if m._ws is None else (m._ws.data_ptr(), m._ws.numel()),
                None if m._tc is None else m._tc.data_ptr(),
                None if getattr(m, "_app16", None) is None else m._app16.data_ptr(),
                None if getattr(m, "_bg_packed", None) is None else m._bg_packed.data_ptr(),
                None if getattr(m, "_bg_tc", None) is None else m._bg_tc.data_ptr())

    def capture(self):
        """Warm up on a side stream (torch's capture protocol), then record the step.  The warm-up steps are real
        optimisation steps on whatever the static buffers hold: parameters, the step counter AND the Adam moments are
        snapshotted and restored (moments created by the warm-up itself are zeroed), so capturing -- or re-capturing after a
        maintenance call -- does not disturb a run in progress."""
        m, opt = self.model, self.opt
        params = [p for g in opt.param_groups for p in g["params"]]
        live = {id(p) for p in m._param_list()} | ({id(p) for p in m._bg_param_list()} if self.npp else set())
        if not live <= {id(p) for p in params}:
            raise RuntimeError("TrainStepGraph: the optimizer does not hold the model's current parameters (the grids were "
                               "replaced by upsample_volume_grid / shrink / load): build a new optimizer and a new TrainStepGraph")
        snap = [p.detach().clone() for p in params]
        state_snap = {k: (mv[0].clone(), mv[1].clone()) for k, mv in opt.state.items()}
        n_step0 = opt.n_step
        s = torch.cuda.Stream(device=m.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._upload_scalars()
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        self._upload_scalars()
        with torch.cuda.graph(self.graph, capture_error_mode=os.environ.get("TVM_GRAPH_CAPTURE_MODE", "global")):
            self._body()
        with torch.no_grad():
            for p, q in zip(params, snap):
                p.copy_(q)
            for k, mv in opt.state.items():
                if k in state_snap and state_snap[k][0].shape == mv[0].shape:
                    mv[0].copy_(state_snap[k][0])
                    mv[1].copy_(state_snap[k][1])
                else:
                    mv[0].zero_()
                    mv[1].zero_()
        opt.n_step = n_step0
        m._pack(force=True)
        self._signature = self._capture_signature()

    def step(self, rays, target, jitter=None):
        if jitter is not None and self.jitter is None:
            if self.graph is not None:
                raise RuntimeError("pass jitter from the first step on (the graph was captured with on-device random jitter)")
            self.jitter = torch.zeros(self.rays.shape[0], dtype=torch.float32, device=self.rays.device)
        if self.graph is None or self._signature != self._capture_signature():
            # first use, or a maintenance call replaced buffers the graph points at: record the step again
            self.graph = None
            self.capture()
        if jitter is not None:
            self.jitter.copy_(jitter, non_blocking=True)
        self.rays.copy_(rays, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        self._upload_scalars()
        self.graph.replay()
        return self.loss
