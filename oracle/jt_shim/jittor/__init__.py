"""A torch-CPU stand-in for the ~40 Jittor calls made by tensorf-myc/models/{tensorBase,tensoRF,sh}.py.

TEST INFRASTRUCTURE.  Jittor is not installed in this image and cannot be (no network).  This shim
exists so that tests/golden/make_golden.py can execute the reference's python modules UNMODIFIED
(imported from /root/reference) and record golden vectors.  It pins the reference's own python
logic; Jittor's op numerics are supplied here from its public python definitions as recalled:
  cumprod(x, dim)      = exp(cumsum(log(x), dim))                             (jittor/misc.py)
  nn.softplus(x,1,20)  = log(1 + exp(min(x, 20))) + max(x - 20, 0)            (jittor/nn.py)
  nn.Linear            = x @ W.T + b, W [out,in]                               (jittor/nn.py)
  nn.grid_sample       = bilinear, zeros padding, align_corners honoured -- evaluated with
                         torch.nn.functional.grid_sample (an independent implementation of the same
                         definition, NOT the oracle's explicit restatement)
  Var.max(dim)         returns values only; Var.clamp(min_v, max_v); Var.transpose() reverses dims.
"""
import math
import sys
import types

import numpy as np
import torch
import torch.nn.functional as _F

_rand_queue = []          # arrays handed out by rand_like / rand, in call order (deterministic "jitter")


class Var(torch.Tensor):
    def __new__(cls, data=None, dtype=None):
        if isinstance(data, torch.Tensor):
            t = data.detach().clone()
        else:
            t = torch.as_tensor(np.asarray(data))
            if t.dtype == torch.float64:
                t = t.float()
        if dtype is not None:
            t = t.to(dtype)
        return t.as_subclass(cls)

    def __init__(self, *a, **k):
        pass

    def max(self, dim=None, keepdims=False, keepdim=False):
        if dim is None:
            return torch.Tensor.max(self)
        return torch.amax(self, dim, keepdim=keepdims or keepdim)

    def min(self, dim=None, keepdims=False, keepdim=False):
        if dim is None:
            return torch.Tensor.min(self)
        return torch.amin(self, dim, keepdim=keepdims or keepdim)

    def clamp(self, min_v=None, max_v=None):
        return torch.clamp(self, min=min_v, max=max_v)

    def float32(self):
        return self.to(torch.float32)

    def int32(self):
        return self.to(torch.int32)

    def transpose(self, *dims):
        if len(dims) == 0:
            return self.permute(*reversed(range(self.dim())))
        return torch.Tensor.transpose(self, *dims)

    def numpy(self):
        return self.detach().as_subclass(torch.Tensor).numpy()


def _v(t):
    return t.as_subclass(Var) if isinstance(t, torch.Tensor) else t


def _shape(args):
    if len(args) == 1 and isinstance(args[0], (tuple, list, torch.Size)):
        return tuple(int(a) for a in args[0])
    return tuple(int(a) for a in args)


def array(data, dtype=None):
    return Var(data, dtype)


def int32(data):
    return Var(data).to(torch.int32)


def zeros(*shape, dtype="float32"):
    return _v(torch.zeros(_shape(shape)))


def ones(*shape, dtype="float32"):
    return _v(torch.ones(_shape(shape)))


def empty(shape, dtype=None):
    return _v(torch.empty(_shape((shape,)), dtype=dtype if isinstance(dtype, torch.dtype) else None))


def randn(*shape):
    return _v(torch.randn(_shape(shape)))


def rand(*shape):
    if _rand_queue:
        return _v(torch.as_tensor(_rand_queue.pop(0)).reshape(_shape(shape)))
    return _v(torch.rand(_shape(shape)))


def rand_like(x):
    if _rand_queue:
        return _v(torch.as_tensor(_rand_queue.pop(0)).to(x.dtype).reshape(x.shape))
    return _v(torch.rand_like(x))


def concat(xs, dim=0):
    return _v(torch.cat(list(xs), dim))


def stack(xs, dim=0):
    return _v(torch.stack(list(xs), dim))


def cumprod(x, dim=0):
    return torch.exp(torch.cumsum(torch.log(x), dim))


def norm(x, p=2, dim=-1, keepdim=False, keepdims=False):
    return torch.norm(x, p=p, dim=dim, keepdim=keepdim or keepdims)


def meshgrid(*xs):
    if len(xs) == 1 and isinstance(xs[0], (list, tuple)):
        xs = xs[0]
    return [_v(t) for t in torch.meshgrid(*xs, indexing="ij")]


def arange(*a, **k):
    return _v(torch.arange(*a, **k))


def linspace(*a, **k):
    return _v(torch.linspace(*a, **k))


def split(x, size, dim=0):
    return [_v(t) for t in torch.split(x, size, dim)]


def normalize(input, p=2, dim=1, eps=1e-30):
    """jittor/misc.py (as recalled): input / input.norm(p, dim, keepdims=True, eps)."""
    return input / torch.sqrt(torch.clamp(torch.sum(input * input, dim, keepdim=True), min=eps))


def broadcast(x, shape, dims=()):
    """jt.broadcast(x, shape, dims): insert the listed dims, then expand to `shape`."""
    nd = len(shape)
    for d in sorted(d % nd for d in dims):
        x = x.unsqueeze(d)
    return x.expand(*shape)


def clamp(x, min_v=None, max_v=None):
    return torch.clamp(x, min=min_v, max=max_v)


def sqr(x):
    return x * x


def cross(a, b, dim=-1):
    return torch.cross(a, b, dim=dim)


def flip(x, dim=0):
    return torch.flip(x, [dim] if isinstance(dim, int) else list(dim))


float32 = torch.float32
int64 = torch.int64
asin = torch.asin
multiply = torch.mul
zeros_like = torch.zeros_like
ones_like = torch.ones_like
full_like = torch.full_like
where = torch.where
minimum = torch.minimum
maximum = torch.maximum
exp = torch.exp
log = torch.log
sin = torch.sin
cos = torch.cos
sigmoid = torch.sigmoid
sum = torch.sum
mean = torch.mean
sqrt = torch.sqrt
pow = torch.pow
abs = torch.abs
matmul = torch.matmul
round = torch.round
all = torch.all
relu = torch.relu
no_grad = torch.no_grad


class _Flags:
    no_grad = 0
    use_cuda = 0


flags = _Flags()


def sync_all(*a, **k):
    pass


def gc():
    pass


# ---- jittor.nn -------------------------------------------------------------------------------
nn = types.ModuleType("jittor.nn")


class Module(torch.nn.Module):
    def forward(self, *a, **k):
        return self.execute(*a, **k)


class Linear(Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        b = 1.0 / math.sqrt(in_features)
        self.weight = torch.nn.Parameter(Var(torch.empty(out_features, in_features).uniform_(-b, b)))
        self.bias = torch.nn.Parameter(Var(torch.empty(out_features).uniform_(-b, b))) if bias else None

    def execute(self, x):
        y = torch.matmul(x, self.weight.t())
        return y + self.bias if self.bias is not None else y


def grid_sample(input, grid, mode="bilinear", padding_mode="zeros", align_corners=False):
    return _v(_F.grid_sample(input.as_subclass(torch.Tensor), grid.as_subclass(torch.Tensor), mode=mode,
                             padding_mode=padding_mode, align_corners=align_corners))


def softplus(x, beta=1.0, threshold=20.0):
    return 1 / beta * torch.log(1 + torch.exp(torch.clamp(beta * x, max=threshold))) + \
        torch.clamp(x - threshold / beta, min=0.0)


def Parameter(x, requires_grad=True):
    return torch.nn.Parameter(x if isinstance(x, Var) else Var(x), requires_grad=requires_grad)


nn.Module = Module
nn.Linear = Linear
nn.ReLU = torch.nn.ReLU
nn.Sigmoid = torch.nn.Sigmoid
nn.ModuleList = torch.nn.ModuleList
nn.Sequential = torch.nn.Sequential
nn.Parameter = Parameter
nn.ParameterList = torch.nn.ParameterList
nn.init = torch.nn.init
nn.grid_sample = grid_sample
nn.softplus = softplus
nn.relu = torch.relu
nn.interpolate = _F.interpolate
nn.max_pool3d = _F.max_pool3d
sys.modules["jittor.nn"] = nn
