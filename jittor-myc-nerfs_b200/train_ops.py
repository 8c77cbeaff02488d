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

import torch

from . import _lib as L


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _launch(kind, x, weight, loss, grad):
    lib, st = L.load(), _stream_ptr()
    if kind == "tv":
        _, Cc, H, W = x.shape
        L.check(lib.tvm_tv_loss(_ptr(x), Cc, H, W, float(weight), _ptr(loss), _ptr(grad), st), "tvm_tv_loss")
    elif kind == "l1":
        L.check(lib.tvm_l1_loss(_ptr(x), x.numel(), float(weight), _ptr(loss), _ptr(grad), st), "tvm_l1_loss")
    else:
        _, Cc, Ln, _ = x.shape
        L.check(lib.tvm_vector_diffs(_ptr(x), Cc, Ln, float(weight), _ptr(loss), _ptr(grad), st), "tvm_vector_diffs")


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

    @torch.no_grad()
    def step(self, loss=None):
        if loss is not None:
            self.zero_grad()
            loss.backward()
        self.n_step += 1
        entries, keep = [], []
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.data.is_contiguous()):
                    raise L.TvmError("Adam: parameters must be contiguous fp32 CUDA tensors (no CPU fallback)")
                m, v = self._state(p)
                gr = p.grad.contiguous()
                keep.append(gr)
                e = L.TvmAdamTensor()
                e.p, e.g, e.m, e.v, e.n, e.lr = p.data.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(g["lr"])
                entries.append(e)
                keep.append(p)
        if not entries:
            return
        arr = (L.TvmAdamTensor * len(entries))(*entries)
        L.check(L.load().tvm_adam_step(arr, len(entries), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                       int(self.n_step), _stream_ptr()), "tvm_adam_step")
        # the kernel wrote through raw pointers: bump the version counters so that the packed device image is rebuilt
        for t in keep:
            if isinstance(t, torch.nn.Parameter):
                torch.autograd.graph.increment_version(t)
