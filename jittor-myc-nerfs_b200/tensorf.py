"""Host-side mirror of tensorf-myc's TensorVMSplit / AlphaGridMask for the ray-rendering hot path.

Same constructor keywords, attribute names, parameter names/shapes and call signature as the
reference (tensorf-myc/models/tensorBase.py:140-176, 476-536; models/tensoRF.py:141-174), so the
reference's driver code (train.py:167-173, renderer.py:12-27) can hold this object instead.
All arithmetic happens in libtvmrender.so (hand-written sm_100a kernels, include/tvmrender.h);
torch is used for device memory, streams and autograd bookkeeping only.  No CPU path exists.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import weakref

import numpy as np
import torch

from . import _lib as L
from .checkpoint import CheckpointMixin
from .maintain import MaintainMixin
from .train_ops import RegularizerMixin

MAT_MODE = [[0, 1], [0, 2], [1, 2]]   # tensorBase.py:168
VEC_MODE = [2, 1, 0]                  # tensorBase.py:169

_MLP_FLAGS = {"fp32": L.MLP_FP32, "bf16": L.MLP_BF16, "fp16": L.MLP_FP16}


def derive_march_scalars(aabb, gridSize, step_ratio):
    """TensorBase.update_stepSize (tensorBase.py:197-209) in fp32, op for op."""
    f32 = np.float32
    aabb = np.asarray(aabb, dtype=f32).reshape(2, 3)
    aabbSize = aabb[1] - aabb[0]
    invaabbSize = f32(2.0) / aabbSize
    g = np.asarray([int(i) for i in gridSize], dtype=np.int32)
    units = aabbSize / (g - 1).astype(f32)
    stepSize = f32(np.mean(units, dtype=f32) * f32(step_ratio))
    aabbDiag = np.sqrt(np.sum(np.power(aabbSize, f32(2)), dtype=f32), dtype=f32)
    nSamples = int(f32(aabbDiag / stepSize)) + 1
    return dict(aabbSize=aabbSize, invaabbSize=invaabbSize.astype(f32), gridSize=g, units=units,
                stepSize=stepSize, aabbDiag=aabbDiag, nSamples=nSamples)


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class AlphaGridMask:
    """tensorBase.py:39-59.  Keeps the reference's float volume view and a bit-packed device copy."""

    def __init__(self, device, aabb, alpha_volume, packed_bits=None):
        self.device = device
        self.aabb = torch.as_tensor(np.asarray(aabb.detach().cpu() if torch.is_tensor(aabb) else aabb),
                                    dtype=torch.float32).reshape(2, 3)
        a = self.aabb.numpy().astype(np.float32)
        self.aabbSize = a[1] - a[0]
        self.invgridSize = (np.float32(1.0) / self.aabbSize * np.float32(2)).astype(np.float32)  # tensorBase.py:46
        vol = torch.as_tensor(alpha_volume, dtype=torch.float32)
        self.alpha_volume = vol.reshape(1, 1, *vol.shape[-3:]).to(device).contiguous()
        D, H, W = self.alpha_volume.shape[-3:]
        self.gridSize = torch.tensor([W, H, D], dtype=torch.int32)
        n_words = (D * H * W + 31) // 32
        lib = L.load()
        if packed_bits is not None:          # updateAlphaMask already produced the bit stream (tvm_alpha_mask_from_dense)
            assert packed_bits.numel() >= n_words and packed_bits.dtype == torch.int32
            self.bits = packed_bits
        else:
            self.bits = torch.empty(n_words + 8, dtype=torch.int32, device=device)
            L.check(lib.tvm_pack_alpha(_ptr(self.alpha_volume), D, H, W, _ptr(self.bits), _stream_ptr()),
                    "tvm_pack_alpha")
        n_bricks = ((D + 7) // 8) * ((H + 7) // 8) * ((W + 7) // 8)
        self.bricks = torch.empty((n_bricks + 31) // 32 + 8, dtype=torch.int32, device=device)
        L.check(lib.tvm_pack_alpha_bricks(_ptr(self.bits), D, H, W, _ptr(self.bricks), _stream_ptr()),
                "tvm_pack_alpha_bricks")
        self.bricks3 = torch.empty(n_bricks + 8, dtype=torch.int32, device=device)       # 27-bit neighbourhood word per brick
        L.check(lib.tvm_pack_alpha_bricks3(_ptr(self.bricks), D, H, W, _ptr(self.bricks3), _stream_ptr()),
                "tvm_pack_alpha_bricks3")
        self.dilated = torch.empty(n_words + 8, dtype=torch.int32, device=device)
        L.check(lib.tvm_pack_alpha_dilated(_ptr(self.bits), D, H, W, _ptr(self.dilated), _stream_ptr()),
                "tvm_pack_alpha_dilated")


class _Linear(torch.nn.Module):
    """Parameter holder with torch.nn.Linear's names (weight [out,in], bias [out])."""

    def __init__(self, in_c, out_c, bias=True):
        super().__init__()
        b = 1.0 / math.sqrt(in_c)
        self.weight = torch.nn.Parameter(torch.empty(out_c, in_c).uniform_(-b, b))
        self.bias = torch.nn.Parameter(torch.empty(out_c).uniform_(-b, b)) if bias else None


class MLPRender_Fea(torch.nn.Module):
    """Parameter container with the reference's names (tensorBase.py:62-74): mlp.0 / mlp.2 / mlp.4."""

    def __init__(self, inChanel, viewpe=6, feape=6, featureC=128, extra_in=0):
        super().__init__()
        self.in_mlpC = 2 * viewpe * 3 + 2 * feape * inChanel + 3 + inChanel + extra_in
        self.viewpe, self.feape = viewpe, feape
        self.mlp = torch.nn.ModuleList([_Linear(self.in_mlpC, featureC), torch.nn.Identity(),
                                        _Linear(featureC, featureC), torch.nn.Identity(),
                                        _Linear(featureC, 3)])
        torch.nn.init.constant_(self.mlp[-1].bias, 0)


class _RenderFn(torch.autograd.Function):
    """Autograd node: tvm_forward / tvm_backward; gradients come back in the parameters' NCHW shapes."""

    @staticmethod
    def forward(ctx, model, rays, jitter, flags, S, *params):
        rgb, depth = model._forward_raw(rays, jitter, flags, S)
        ctx.model, ctx.flags, ctx.S = model, flags, S
        ctx.stamp = model._forward_stamp()
        ctx.save_for_backward(rays, jitter if jitter is not None else torch.empty(0, device=rays.device), rgb)
        ctx.mark_non_differentiable(depth)
        # REFTensoRF: the normal penalty (REFTensoRF.py:236-238) is a third, differentiable output
        penalty = model._penalty_buf.clone() if model.VARIANT == L.VARIANT_REF else torch.zeros(1, device=rays.device)
        return rgb, depth, penalty

    @staticmethod
    def backward(ctx, d_rgb, _d_depth, d_penalty):
        rays, jitter, rgb = ctx.saved_tensors
        ctx.model._check_stamp(ctx.stamp)
        d_pen = None
        if ctx.model.VARIANT == L.VARIANT_REF and d_penalty is not None:
            d_pen = d_penalty.reshape(-1)[:1].to(torch.float32).contiguous()
        d_rgb = torch.zeros_like(rgb) if d_rgb is None else d_rgb.contiguous()
        grads = ctx.model._backward_raw(rays, jitter if jitter.numel() else None, ctx.flags, ctx.S, rgb, d_rgb, d_pen)
        return (None, None, None, None, None, *grads)


class _RenderNppFn(torch.autograd.Function):
    """Autograd node of NerfPlusPlus: tvm_forward_npp / tvm_backward_npp (foreground grids + head, background MLPNet)."""

    @staticmethod
    def forward(ctx, model, rays, fg_rand, bg_rand, flags, S, *params):
        rgb, depth = model._forward_npp_raw(rays, fg_rand, bg_rand, flags, S)
        ctx.model, ctx.flags, ctx.S = model, flags, S
        ctx.stamp = model._forward_stamp()
        ctx.save_for_backward(rays, fg_rand, bg_rand, rgb)
        ctx.mark_non_differentiable(depth)
        return rgb, depth

    @staticmethod
    def backward(ctx, d_rgb, _d_depth):
        rays, fg_rand, bg_rand, rgb = ctx.saved_tensors
        ctx.model._check_stamp(ctx.stamp)
        d_rgb = torch.zeros_like(rgb) if d_rgb is None else d_rgb.contiguous()
        grads = ctx.model._backward_npp_raw(rays, fg_rand, bg_rand, ctx.flags, ctx.S, rgb, d_rgb)
        return (None, None, None, None, None, None, *grads)


class TensorVMSplit(MaintainMixin, RegularizerMixin, CheckpointMixin, torch.nn.Module):
    VARIANT = L.VARIANT_VM

    def __init__(self, aabb, gridSize, device, density_n_comp=8, appearance_n_comp=24, app_dim=27,
                 shadingMode='MLP_PE', alphaMask=None, near_far=[2.0, 20.0],
                 density_shift=-10, alphaMask_thres=0.001, distance_scale=25, rayMarch_weight_thres=0.0001,
                 pos_pe=6, view_pe=6, fea_pe=6, featureC=128, step_ratio=2.0,
                 fea2denseAct='softplus'):
        super().__init__()
        L.require_cuda()
        if shadingMode != 'MLP_Fea':
            raise NotImplementedError("only shadingMode='MLP_Fea' (the mode of all shipped configs) is on the hot path")
        if isinstance(density_n_comp, int):
            density_n_comp = [density_n_comp] * 3
        if isinstance(appearance_n_comp, int):
            appearance_n_comp = [appearance_n_comp] * 3
        if len(set(density_n_comp)) != 1 or len(set(appearance_n_comp)) != 1:
            raise NotImplementedError("per-component channel counts must be equal")
        self.density_n_comp = list(density_n_comp)
        self.app_n_comp = list(appearance_n_comp)
        self.app_dim = app_dim
        self.device = device
        self.aabb = torch.as_tensor(np.asarray(aabb.detach().cpu() if torch.is_tensor(aabb) else aabb),
                                    dtype=torch.float32).reshape(2, 3)
        self.alphaMask = alphaMask
        self.density_shift = density_shift
        self.alphaMask_thres = alphaMask_thres
        self.distance_scale = distance_scale
        self.rayMarch_weight_thres = rayMarch_weight_thres
        self.fea2denseAct = fea2denseAct
        self.near_far = near_far
        self.step_ratio = step_ratio
        self.update_stepSize(gridSize)
        self.matMode = MAT_MODE
        self.vecMode = VEC_MODE
        self.comp_w = [1, 1, 1]
        self.init_svd_volume(gridSize[0], device)
        self.shadingMode, self.pos_pe, self.view_pe, self.fea_pe, self.featureC = \
            shadingMode, pos_pe, view_pe, fea_pe, featureC
        self.init_render_func(shadingMode, pos_pe, view_pe, fea_pe, featureC, device)
        # --- engine state -------------------------------------------------------------------
        self.mlp_mode = os.environ.get("TVM_MLP_MODE", "fp32")
        # mlp_mode "bf16" / "fp16": gather the appearance-plane texels from 16-bit copies in the mode's format (half the gather bytes of the head;
        # the plane x line products are rounded to bf16 as the GEMM operand in that mode anyway).  Backward kernels and
        # the fp32 mode always read the fp32 planes.
        self.app_planes_bf16 = os.environ.get("TVM_APP_PLANES", "fp32") == "bf16"
        self.early_termination = True
        self.fused_composite = os.environ.get("TVM_FUSED_COMPOSITE", "1") == "1"     # evaluation renders (no gradient) pass TVM_EVAL_ONLY: compositing inside the appearance head
        self.empty_space_skipping = True
        self.collect_counters = False
        self.counters = torch.zeros(L.CNT_WORDS, dtype=torch.int64, device=device)
        self.ws_budget_bytes = int(float(os.environ.get("TVM_WS_GIB", "2")) * (1 << 30))    # both workspaces of the evaluation pipeline; bounded entry lists (an 800x800 frame of a trained scene: one launch)
        self.grad_sync = False          # set True (after dist.init_from_env) for data-parallel training
        self.grad_sync_group = None
        self._peer_comm = None          # enable_peer_allreduce(): libtvmrender's own all-reduce over NVLink peer memory
        self.train_ws_budget_bytes = int(float(os.environ.get("TVM_TRAIN_WS_GIB", "16")) * (1 << 30))   # worst-case workspace of one training launch
        self.defer_overflow_check = False      # True: bounded evaluation renders are verified by verify_renders(), not at once
        self._pending_checks = []
        self.stream_stages = 8          # pipeline stages of a host-to-host frame: a count (equal pieces) or relative sizes (renderer._render_streamed)
        self.ws_overflows = 0           # ranges rendered a second time because their entry list overflowed
        self._epr_hint = None           # entries per ray (+30 %) the next bounded render sizes its entry lists from
        self._ws = None
        self._fwd_gen = 0               # bumped by every tvm_forward* on the shared workspace (see _forward_stamp)
        self._packed = None
        self._packed_versions = None
        self._packed_grid = None
        self._tc = None

    # ---- reference API: parameters --------------------------------------------------------
    def update_stepSize(self, gridSize):
        s = derive_march_scalars(self.aabb.numpy(), gridSize, self.step_ratio)
        self.aabbSize = torch.from_numpy(s["aabbSize"].copy())
        self.invaabbSize = torch.from_numpy(s["invaabbSize"].copy())
        self.gridSize = torch.from_numpy(s["gridSize"].copy())
        self.units = torch.from_numpy(s["units"].copy())
        self.stepSize = torch.tensor(s["stepSize"])
        self.aabbDiag = torch.tensor(s["aabbDiag"])
        self.nSamples = s["nSamples"]
        self._packed_grid = None

    def init_render_func(self, shadingMode, pos_pe, view_pe, fea_pe, featureC, device):
        self.renderModule = MLPRender_Fea(self.app_dim, view_pe, fea_pe, featureC).to(device)

    def head_dim(self):
        return 32

    def init_svd_volume(self, res, device):
        self.density_plane, self.density_line = self.init_one_svd(self.density_n_comp, self.gridSize, 0.1, device)
        self.app_plane, self.app_line = self.init_one_svd(self.app_n_comp, self.gridSize, 0.1, device)
        self.basis_mat = _Linear(sum(self.app_n_comp), self.app_dim, bias=False).to(device)

    def init_one_svd(self, n_component, gridSize, scale, device):
        plane_coef, line_coef = [], []
        for i in range(len(self.vecMode)):
            vec_id = self.vecMode[i]
            mat_id_0, mat_id_1 = self.matMode[i]
            plane_coef.append(torch.nn.Parameter(scale * torch.randn(
                (1, n_component[i], int(gridSize[mat_id_1]), int(gridSize[mat_id_0])), device=device)))
            line_coef.append(torch.nn.Parameter(scale * torch.randn(
                (1, n_component[i], int(gridSize[vec_id]), 1), device=device)))
        return torch.nn.ParameterList(plane_coef), torch.nn.ParameterList(line_coef)

    def get_optparam_groups(self, lr_init_spatialxyz=0.02, lr_init_network=0.001):
        """tensoRF.py:168-174."""
        return [{'params': self.density_line, 'lr': lr_init_spatialxyz},
                {'params': self.density_plane, 'lr': lr_init_spatialxyz},
                {'params': self.app_line, 'lr': lr_init_spatialxyz},
                {'params': self.app_plane, 'lr': lr_init_spatialxyz},
                {'params': self.basis_mat.parameters(), 'lr': lr_init_network},
                {'params': self.renderModule.parameters(), 'lr': lr_init_network}]

    def set_nerfplusplus(self, bg_freq=4, bg_view_freq=2, bg_D=4, radii=20):
        """tensorBase.py:538-539: a no-op for non-NeRF++ classes."""
        pass

    def load_numpy_params(self, p):
        """Copy an synthetic.ModelParams (numpy, reference shapes) into the parameters."""
        with torch.no_grad():
            for k in range(3):
                self.density_plane[k].copy_(torch.from_numpy(p.density_plane[k]))
                self.density_line[k].copy_(torch.from_numpy(p.density_line[k]))
                self.app_plane[k].copy_(torch.from_numpy(p.app_plane[k]))
                self.app_line[k].copy_(torch.from_numpy(p.app_line[k]))
            self.basis_mat.weight.copy_(torch.from_numpy(p.basis_mat))
            for i, li in enumerate((0, 2, 4)):
                self.renderModule.mlp[li].weight.copy_(torch.from_numpy(p.mlp_w[i]))
                self.renderModule.mlp[li].bias.copy_(torch.from_numpy(p.mlp_b[i]))

    def _param_list(self):
        m = self.renderModule.mlp
        return [*self.density_plane, *self.density_line, *self.app_plane, *self.app_line, self.basis_mat.weight,
                m[0].weight, m[0].bias, m[2].weight, m[2].bias, m[4].weight, m[4].bias]

    # ---- packed device image ----------------------------------------------------------------
    def _layout(self):
        """Offsets (in floats) of every packed tensor inside the flat parameter / gradient buffers."""
        G = [int(g) for g in self.gridSize]
        Cd, Ca = self.density_n_comp[0], self.app_n_comp[0]
        F, in_c = self.featureC, self.renderModule.in_mlpC
        items, off = {}, 0

        def add(name, n):
            nonlocal off
            items[name] = (off, n)
            off += (n + 63) // 64 * 64   # 256-byte alignment

        for k in range(3):
            m0, m1 = MAT_MODE[k]
            add(f"dp{k}", G[m1] * G[m0] * Cd)
        for k in range(3):
            add(f"dl{k}", G[VEC_MODE[k]] * Cd)
        for k in range(3):
            m0, m1 = MAT_MODE[k]
            add(f"ap{k}", G[m1] * G[m0] * Ca)
        for k in range(3):
            add(f"al{k}", G[VEC_MODE[k]] * Ca)
        add("basis_t", 3 * Ca * self.head_dim())
        add("head_bias", 64)
        add("w1_t", in_c * F)
        add("b1", F)
        add("w2_t", F * F)
        add("b2", F)
        add("w3", 3 * F)
        add("b3", 64)
        return items, off

    def _struct_for(self, buf, cls):
        items, _ = self._layout()
        base = buf.data_ptr()
        at = lambda name: base + 4 * items[name][0]
        s = cls()
        for k in range(3):
            s.density_plane[k] = at(f"dp{k}")
            s.density_line[k] = at(f"dl{k}")
            s.app_plane[k] = at(f"ap{k}")
            s.app_line[k] = at(f"al{k}")
        for name in ("basis_t", "w1_t", "b1", "w2_t", "b2", "w3", "b3"):
            setattr(s, name, at(name))
        if cls is L.TvmModel:
            s.variant = self.VARIANT
        s.head_bias = at("head_bias") if self.VARIANT == L.VARIANT_REF else None
        return s

    def _pack(self, force=False):
        """(Re)build the channels-last / transposed device image when a parameter changed."""
        params = self._param_list()
        versions = tuple((p.data_ptr(), p._version) for p in params)
        grid_key = tuple(int(g) for g in self.gridSize)
        if not force and self._packed is not None and versions == self._packed_versions \
                and self._packed_grid == grid_key:
            return
        lib = L.load()
        items, total = self._layout()
        if self._packed is None or self._packed.numel() != total:
            self._packed = torch.zeros(total, dtype=torch.float32, device=self.device)
            self._grads_packed = None
        st = _stream_ptr()
        base = self._packed.data_ptr()
        at = lambda name: C.c_void_p(base + 4 * items[name][0])
        Cd, Ca, F = self.density_n_comp[0], self.app_n_comp[0], self.featureC
        in_c = self.renderModule.in_mlpC
        jobs = []
        for k in range(3):
            for pref, plist, llist, c in (("d", self.density_plane, self.density_line, Cd),
                                          ("a", self.app_plane, self.app_line, Ca)):
                p, l = plist[k].detach(), llist[k].detach()
                assert p.is_contiguous() and l.is_contiguous() and p.dtype == torch.float32
                # NCHW [C][H*W] -> channels-last [H*W][C] (tvm_pack_grid), all 12 grids in one launch
                jobs.append(L.TvmTransposeJob(p.data_ptr(), at(f"{pref}p{k}").value, c, p.shape[2] * p.shape[3]))
                jobs.append(L.TvmTransposeJob(l.data_ptr(), at(f"{pref}l{k}").value, c, l.shape[2]))
        m = self.renderModule.mlp
        jobs += self._pack_heads(lib, at, items, st)
        J = L.TvmTransposeJob
        # Linear [out][in] -> [in][out_pad] (tvm_pack_linear) and the bias / last-layer copies (1-row jobs) ride in the same launch
        jobs += [J(m[0].weight.data_ptr(), at("w1_t").value, F, in_c, 0, F), J(m[2].weight.data_ptr(), at("w2_t").value, F, F, 0, F),
                 J(m[0].bias.data_ptr(), at("b1").value, 1, F), J(m[2].bias.data_ptr(), at("b2").value, 1, F),
                 J(m[4].weight.data_ptr(), at("w3").value, 1, 3 * F), J(m[4].bias.data_ptr(), at("b3").value, 1, 3)]
        L.check(lib.tvm_transpose_batch((J * len(jobs))(*jobs), len(jobs), st), "tvm_transpose_batch")
        self._packed_versions = versions
        self._packed_grid = grid_key
        self._model_struct = None
        self._tc_stale = True

    def _pack_heads(self, lib, at, items, st):
        """basis_mat [app_dim][3 Ca] -> [3 Ca][32]; returns transpose jobs for the caller's batch."""
        Ca = self.app_n_comp[0]
        return [L.TvmTransposeJob(self.basis_mat.weight.data_ptr(), at("basis_t").value, self.app_dim, 3 * Ca, 0, 32)]

    def _invalidate_packed(self):
        """Grids were replaced (upsample / shrink): force a re-pack and a fresh TvmModel."""
        self._packed_versions = None
        self._model_struct = None

    def _model(self):
        """The TvmModel descriptor (host POD) for the current parameters / mask."""
        self._pack()
        mask_key = (self._mask_version, self.empty_space_skipping, self.mlp_mode, self.app_planes_bf16)
        if getattr(self, "_model_struct", None) is not None and self._model_mask_key == mask_key \
                and not (self._tc_stale and self.mlp_mode != "fp32"):
            return self._model_struct
        s = self._struct_for(self._packed, L.TvmModel)
        a = self.aabb.numpy().astype(np.float32)
        for i in range(3):
            s.aabb[i] = float(a[0, i])
            s.aabb[3 + i] = float(a[1, i])
            s.inv_aabb_size[i] = float(self.invaabbSize[i])
            s.grid[i] = int(self.gridSize[i])
        s.step_size = float(self.stepSize)
        s.near_, s.far_ = float(self.near_far[0]), float(self.near_far[1])
        s.density_shift = float(self.density_shift)
        s.distance_scale = float(self.distance_scale)
        s.weight_thres = float(self.rayMarch_weight_thres)
        s.act = L.ACT_SOFTPLUS if self.fea2denseAct == "softplus" else L.ACT_RELU
        s.n_density, s.n_app, s.app_dim = self.density_n_comp[0], self.app_n_comp[0], self.app_dim
        s.view_pe, s.fea_pe, s.feature_c = self.view_pe, self.fea_pe, self.featureC
        if self.alphaMask is not None:
            am = self.alphaMask
            s.alpha_bits = am.bits.data_ptr()
            s.alpha_bricks = am.bricks.data_ptr() if self.empty_space_skipping else None
            s.alpha_dilated = am.dilated.data_ptr() if self.empty_space_skipping else None
            s.alpha_bricks3 = am.bricks3.data_ptr() if (self.empty_space_skipping and not os.environ.get("TVM_NO_BRICKS3")) else None
            a0 = am.aabb.numpy().astype(np.float32)
            for i in range(3):
                s.alpha_grid[i] = int(am.gridSize[i])
                s.alpha_aabb_min[i] = float(a0[0, i])
                s.alpha_inv_size[i] = float(am.invgridSize[i])
        else:
            s.alpha_bits = None
            s.alpha_bricks = None
            s.alpha_dilated = None
            s.alpha_bricks3 = None
        s.tc_weights = None
        s.tc_weights_bwd = None
        if self.mlp_mode != "fp32":
            lib = L.load()
            nbytes = lib.tvm_tc_weights_bytes(C.byref(s))
            if nbytes == 0:
                raise L.TvmError("tensor-core appearance head unavailable in this libtvmrender.so")
            if self._tc is None or self._tc.numel() < nbytes:
                self._tc = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            L.check(lib.tvm_pack_mlp_tc(C.byref(s), _ptr(self._tc), _MLP_FLAGS[self.mlp_mode], _stream_ptr()), "tvm_pack_mlp_tc")
            s.tc_weights = self._tc.data_ptr()
            if self.mlp_mode == "fp16":
                # fp16 forward, bf16 tensor-core backward: the backward kernel reads its own bf16 operand image
                if getattr(self, "_tc_bwd", None) is None or self._tc_bwd.numel() < nbytes:
                    self._tc_bwd = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                L.check(lib.tvm_pack_mlp_tc(C.byref(s), _ptr(self._tc_bwd), L.MLP_BF16, _stream_ptr()), "tvm_pack_mlp_tc")
                s.tc_weights_bwd = self._tc_bwd.data_ptr()
            self._tc_stale = False
        for k in range(3):
            s.app_plane_pair[k] = None
            s.app_line_pair[k] = None
        if self.mlp_mode in ("bf16", "fp16") and self.app_planes_bf16:
            # 16-bit pair records of the appearance planes and lines (tvmrender.h: TvmModel.app_plane_pair): 2 values per texel
            lib = L.load()
            items, _ = self._layout()
            G = [int(g) for g in self.gridSize]
            Ca = self.app_n_comp[0]
            jobs = []
            for k in range(3):
                m0, m1 = MAT_MODE[k]
                jobs.append((items[f"ap{k}"][0], G[m1], G[m0], s.app_plane_pair, k))
            for k in range(3):
                jobs.append((items[f"al{k}"][0], 1, G[VEC_MODE[k]], s.app_line_pair, k))
            n_tot = sum((2 * rows * W * Ca + 7) // 8 * 8 for _, rows, W, _, _ in jobs)
            if getattr(self, "_app16", None) is None or self._app16.numel() != n_tot:
                self._app16 = torch.empty(n_tot, dtype=torch.bfloat16, device=self.device)
            off = 0
            for o, rows, W, field, k in jobs:
                dst = self._app16.data_ptr() + 2 * off
                L.check(lib.tvm_pack_pair16(C.c_void_p(self._packed.data_ptr() + 4 * o), rows, W, Ca, C.c_void_p(dst),
                                            _MLP_FLAGS[self.mlp_mode], _stream_ptr()), "tvm_pack_pair16")
                field[k] = dst
                off += (2 * rows * W * Ca + 7) // 8 * 8
        s.sampling, s.radii = L.SAMPLING_UNIFORM, 0.0
        self._finish_model(s)
        self._model_struct, self._model_mask_key = s, mask_key
        return s

    def _finish_model(self, s):
        """Hook for variants to fill further TvmModel fields."""
        pass

    # ---- workspace ------------------------------------------------------------------------------
    def workspace_bytes(self, n, S):
        out = C.c_size_t(0)
        L.check(L.load().tvm_workspace_bytes(int(n), int(S), C.byref(out)), "tvm_workspace_bytes")
        return out.value

    def max_rays_per_launch(self, S):
        """Rays of one TRAINING launch: its workspace is the worst case (every sample weighted, 44 B x n x S -- tvm_backward
        reads the stash), bounded by `train_ws_budget_bytes` (16 GiB: 380 k rays at S = 1036; the reference trains on 4096)."""
        per_ray = self.workspace_bytes(1024, S) / 1024.0
        return max(1024, int(self.train_ws_budget_bytes / per_ray) // 1024 * 1024)

    def _workspace(self, n, S, slot=0, nbytes=None):
        """Caller-owned scratch of tvm_forward.  Slot 0 is THE workspace (what tvm_backward and workspace_view read); slot 1
        is the second buffer of the two-stream evaluation pipeline.  `nbytes`: a bounded workspace of that size instead of
        the worst case for (n, S)."""
        need = self.workspace_bytes(n, S) if nbytes is None else int(nbytes)
        if slot == 0:
            if self._ws is None or self._ws.numel() < need:
                self._ws = None
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            return self._ws
        if getattr(self, "_ws2", None) is None or self._ws2.numel() < need:
            self._ws2 = None
            self._ws2 = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws2

    # ---- evaluation renders through bounded workspaces (tvmrender.h: "Bounded workspaces") ------------
    def _plan_launch(self, n, S):
        """(rays per launch, workspace bytes or None) of an evaluation render.  None: the worst case fits, nothing to check.
        Otherwise the entry list is sized from what the last renders needed per ray (`_epr_hint`, + 30 %; S / 8 before the
        first one) and the launch is followed by an overflow check."""
        half = self.ws_budget_bytes // 2
        if self.workspace_bytes(n, S) <= half:
            return n, None
        lib, out = L.load(), C.c_size_t(0)
        L.check(lib.tvm_workspace_bytes_bounded(1024, int(S), 0, C.byref(out)), "tvm_workspace_bytes_bounded")
        per_ray = out.value / 1024.0 + 44.0 * (self._epr_hint or S / 8.0)   # per-ray tables + entries
        nr = max(1024, int(half / per_ray) // 1024 * 1024)
        if nr < n:                                               # equal pieces rather than full ones and a remainder
            pieces = -(-n // nr)
            nr = min(nr, -(-(-(-n // pieces)) // 1024) * 1024)
        nr = min(n, nr)
        if self.workspace_bytes(nr, S) <= half:
            return nr, None
        return nr, half

    def _render_bounded(self, n, S, launch, copy_out=None, max_rays=None, ranges=None):
        """Run `launch(s, e, slot, nbytes)` (one tvm_forward* over rays [s, e) on the current stream with workspace `slot`)
        over [0, n): chunks alternate between the caller's stream and a side stream with one workspace each, so that the
        tail of one chunk's kernels overlaps the next chunk's march.  With bounded workspaces each launch leaves the number
        of entries it WANTED in a status word; one synchronisation at the end compares them with the capacities and renders
        the (rare) overflowed ranges again with the hint corrected.  Rays are independent and compositing is per ray, so the
        pixels do not depend on the chunking (tests/test_gpu_forward.py::test_full_frame_properties).
        `copy_out(s, e, done_event)`: per-chunk hook of the streamed renderer (device -> host copies; a range that overflowed
        is copied again after its second render); `max_rays`: upper bound of a chunk (its pipeline stages)."""
        main = torch.cuda.current_stream()
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        side = self._side_stream
        todo = [(0, n)] if ranges is None else list(ranges)
        while todo:
            jobs = []
            for (s0, e0) in todo:
                nr, nbytes = self._plan_launch(e0 - s0, S)
                if max_rays is not None and nr > max_rays:      # the streamed renderer's pipeline stages
                    nr = max_rays
                    if self.workspace_bytes(nr, S) <= self.ws_budget_bytes // 2:
                        nbytes = None
                jobs += [(s, min(e0, s + nr), nbytes) for s in range(s0, e0, nr)]
            bounded = [j for j in jobs if j[2] is not None]
            status = torch.zeros(max(1, len(bounded)), dtype=torch.int32, device=self.device) if bounded else None
            if len(jobs) > 1:
                side.wait_stream(main)
            k = 0
            for i, (s, e, nbytes) in enumerate(jobs):
                stream = side if (i & 1) else main
                with torch.cuda.stream(stream):
                    ws = launch(s, e, i & 1, nbytes)
                    if ws is None:              # the outputs of a deferred render are gone: nothing to repair
                        continue
                    if nbytes is not None:
                        status[k:k + 1].copy_(ws[:4].view(torch.int32), non_blocking=True)
                        k += 1
                    if copy_out is not None:
                        ev = torch.cuda.Event()
                        ev.record(stream)
                        copy_out(s, e, ev)
            if len(jobs) > 1:
                main.wait_stream(side)
            todo = []
            if bounded and self.defer_overflow_check:
                # the caller verifies later (verify_renders): hand the status words and the way to render a range again over
                self._pending_checks.append((status, bounded, S, launch, copy_out, max_rays))
            elif bounded:
                todo = self._overflowed(status, bounded, S)

    def _overflowed(self, status, bounded, S):
        """Ranges of `bounded` launches whose entry lists overflowed (reads the status words: synchronises); updates the hint."""
        lib = L.load()
        wanted = status.cpu().numpy().astype("int64") & 0xFFFFFFFF
        todo, epr = [], 0.0
        for (s, e, nbytes), w in zip(bounded, wanted):
            cap = C.c_uint32(0)
            L.check(lib.tvm_workspace_capacity(e - s, int(S), nbytes, C.byref(cap)), "tvm_workspace_capacity")
            epr = max(epr, float(w) / (e - s))
            if int(w) > cap.value:
                todo.append((s, e))
        # entries per ray of the densest chunk, + 30 % (+ 60 % after an overflow): what the next render sizes its entry lists
        # from.  The hint never drops by more than 10 % per render: smaller chunks have denser maxima than the ones measured.
        want = (1.6 if todo else 1.3) * epr + 1.0
        self._epr_hint = max(want, 0.9 * self._epr_hint) if self._epr_hint else want
        self.ws_overflows += len(todo)
        return todo

    def verify_renders(self):
        """With `defer_overflow_check = True` evaluation renders return without reading their overflow status (the host can
        enqueue the next frame while this one runs: back-to-back frames keep the GPU busy); call this before consuming the
        pixels of those renders.  It synchronises, renders every overflowed range again into the SAME output tensors, and
        returns the number of ranges it had to repair."""
        repaired = 0
        pending, self._pending_checks = self._pending_checks, []
        if pending:
            torch.cuda.synchronize(self.device)       # every stream of the deferred renders (compute, side, copy)
        for status, bounded, S, launch, copy_out, max_rays in pending:
            todo = self._overflowed(status, bounded, S)
            repaired += len(todo)
            if todo:
                defer, self.defer_overflow_check = self.defer_overflow_check, False
                try:
                    self._render_bounded(0, S, launch, copy_out=copy_out, max_rays=max_rays, ranges=todo)
                    torch.cuda.synchronize(self.device)
                finally:
                    self.defer_overflow_check = defer
        return repaired

    def _forward_chunks(self, rays, jitter, flags, S, rgb, depth):
        # a deferred check keeps `launch` alive until verify_renders(): it must not keep the outputs alive with it (a caller
        # that has dropped them needs no repair, and held blocks would make the allocator grow frame after frame)
        out_ref = (weakref.ref(rgb), weakref.ref(depth))

        def launch(s, e, slot, nbytes):
            o_rgb, o_depth = out_ref[0](), out_ref[1]()
            if o_rgb is None or o_depth is None:
                return None
            self._forward_raw(rays[s:e], None if jitter is None else jitter[s:e], flags, S, out=(o_rgb[s:e], o_depth[s:e]),
                              ws_slot=slot, ws_bytes=nbytes)
            return self._ws if slot == 0 else self._ws2
        self._render_bounded(rays.shape[0], S, launch)

    def workspace_view(self, n, S):
        """What the last tvm_forward over (n rays, S samples) left in the workspace (tvmrender.h: TvmWorkspaceLayout):
        the app_mask bits per 32-sample block, the compacted (ray, sample) entries in ray-major order with their weights
        and colours, and acc_map -- the per-sample record of the PRODUCTION march (no TvmAux, skipping and ERT on).
        After a TVM_EVAL_ONLY launch (evaluation renders with `fused_composite`) only n_entries, ent, ent_w and acc are defined."""
        lay = L.TvmWorkspaceLayout()
        L.check(L.load().tvm_workspace_layout(int(n), int(S), C.byref(lay)), "tvm_workspace_layout")
        ws = self._ws
        assert ws is not None and ws.numel() >= lay.bytes

        def view(off, count, dtype):
            return ws[off:off + count * torch.empty(0, dtype=dtype).element_size()].view(dtype)
        n_ent = int(view(lay.n_entries, 1, torch.int32).item())
        NB = lay.n_blocks
        return dict(n_entries=n_ent, n_blocks=NB,
                    blk_mask=view(lay.blk_mask, n * NB, torch.int32).view(n, NB),
                    blk_base=view(lay.blk_base, n * NB, torch.int32).view(n, NB),
                    ent=view(lay.ent, 2 * n_ent, torch.int32).view(n_ent, 2),
                    ent_w=view(lay.ent_w, n_ent, torch.float32),
                    ent_rgb=view(lay.ent_rgb, 3 * n_ent, torch.float32).view(n_ent, 3),
                    acc=view(lay.acc, n, torch.float32))

    # ---- autograd bookkeeping ---------------------------------------------------------------------
    def _forward_stamp(self):
        """What tvm_backward relies on staying as the matching tvm_forward left it (tvmrender.h: "must follow a tvm_forward
        with the same arguments on the same workspace"): the workspace (any later forward on the model overwrites its
        entry list), the parameters and the mask."""
        return (self._fwd_gen, None if self._ws is None else self._ws.data_ptr(),
                tuple((p.data_ptr(), p._version) for p in self._param_list()),
                self._mask_version, tuple(int(g) for g in self.gridSize))

    def _check_stamp(self, stamp):
        if stamp != self._forward_stamp():
            what = ("another forward ran on this model" if stamp[0] != self._fwd_gen or stamp[1] != (None if self._ws is None else self._ws.data_ptr())
                    else "the parameters, the grid or the alpha mask changed")
            raise RuntimeError("tvm_backward must directly follow its tvm_forward on the model's workspace, but " + what +
                               " between this graph's forward and its backward (e.g. a second micro-batch, an evaluation "
                               "render, compute_alpha / filtering_rays / updateAlphaMask, an optimizer step): call backward() "
                               "before the next forward, or re-run the forward")

    @property
    def alphaMask(self):
        return self._alpha_mask

    @alphaMask.setter
    def alphaMask(self, mask):
        # a monotonically increasing version keys the TvmModel cache (id() of a dropped mask can be reused by the next one)
        self._alpha_mask = mask
        self._mask_version = getattr(self, "_mask_version", 0) + 1
        self._model_struct = None

    # ---- raw engine calls -----------------------------------------------------------------------
    def _flags(self, white_bg):
        f = _MLP_FLAGS[self.mlp_mode]
        if white_bg:
            f |= L.WHITE_BG
        if not self.early_termination:
            f |= L.NO_ERT
        return f

    def _forward_raw(self, rays, jitter, flags, S, aux=None, out=None, ws_slot=0, ws_bytes=None):
        lib = L.load()
        n = rays.shape[0]
        assert rays.is_cuda and rays.dtype == torch.float32 and rays.is_contiguous() and rays.shape[1] == 6
        model = self._model()
        ws = self._workspace(n, S, ws_slot, ws_bytes)
        self._fwd_gen += 1
        if out is None:
            rgb = torch.empty((n, 3), dtype=torch.float32, device=rays.device)
            depth = torch.empty((n,), dtype=torch.float32, device=rays.device)
        else:
            rgb, depth = out
        L.check(lib.tvm_forward(C.byref(model), _ptr(rays), n, int(S), _ptr(jitter), flags, _ptr(rgb), _ptr(depth),
                                C.byref(aux) if aux is not None else None,
                                _ptr(self.counters) if self.collect_counters else None,
                                _ptr(ws), ws.numel() if ws_bytes is None else int(ws_bytes), _stream_ptr()), "tvm_forward")
        return rgb, depth

    def enable_peer_allreduce(self, group=None, n_ctas=64):
        """Data-parallel training without NCCL on the step: the flat gradient buffer moves into symmetric memory and is
        summed by tvm_allreduce_sum (two-shot over NVLink peer memory / NVLS multicast).  Collective: every rank calls it,
        after dist.init_from_env, with the grids at their current size (call it again after upsample_volume_grid)."""
        from .dist import PeerComm
        _, total = self._layout()
        self._peer_comm = PeerComm(total, self.device, group, n_ctas)
        self._peer_total = total
        self._grads_packed = self._peer_comm.buf[:total]
        self.grad_sync, self.grad_sync_group = True, group
        self.grad_sync_kind = ("libtvmrender two-shot all-reduce over NVLink peer memory (" +
                               ("NVLS multimem.ld_reduce / multimem.st" if self._peer_comm.multicast else "peer loads / stores") +
                               f", {n_ctas} CTAs) of the flat packed fp32 gradient buffer")
        return self._peer_comm

    def _backward_raw(self, rays, jitter, flags, S, rgb, d_rgb, d_penalty=None, pipelined=False):
        """-> gradients in the order of _param_list() (reference shapes).  pipelined=True (peer all-reduce only): returns
        (flat buffer, layout, event: appearance half exchanged, event: density half exchanged) instead, for
        _unpack_part."""
        lib = L.load()
        n = rays.shape[0]
        model = self._model()
        items, total = self._layout()
        if getattr(self, "_grads_packed", None) is None or self._grads_packed.numel() != total:
            if self._peer_comm is not None:
                if self._peer_total != total:
                    raise L.TvmError("the parameter layout changed under enable_peer_allreduce(): call it again (collective)")
                self._grads_packed = self._peer_comm.buf[:total]       # the gradient buffer lives in symmetric memory
            else:
                self._grads_packed = torch.empty(total, dtype=torch.float32, device=self.device)
        gp = self._grads_packed
        gp.zero_()
        gs = self._struct_for(gp, L.TvmGrads)
        ws = self._workspace(n, S)
        if self.grad_sync:
            # data-parallel training: ONE all-reduce (sum) over the flat packed gradient buffer.  The average
            # comes for free: every gradient is linear in d_rgb (and d_penalty), so those [n,3] / [1] inputs are scaled by
            # 1/world instead of dividing 3.2 M (128^3) .. 17.4 M (300^3) reduced floats afterwards.
            import torch.distributed as dist
            world = dist.get_world_size(self.grad_sync_group) if dist.is_initialized() else 1
            if world > 1:
                d_rgb = d_rgb * (1.0 / world)
                d_penalty = None if d_penalty is None else d_penalty * (1.0 / world)
        if self.grad_sync and self._peer_comm is not None and os.environ.get("TVM_AR_OVERLAP", "2") != "0":
            # exchange fused into the backward schedule: the appearance part of the buffer is all-reduced on a side stream
            # while k_march_bwd runs, the density part right after it (tvm_backward_dp)
            if getattr(self, "_side_stream", None) is None:
                self._side_stream = torch.cuda.Stream(device=self.device)
            x = L.TvmGradExchange()
            x.comm = C.pointer(self._peer_comm.struct)
            x.split_floats, x.total_floats = items["ap0"][0], (total + 3) // 4 * 4
            x.n_ctas, x.n_ctas_overlapped = self._peer_comm.n_ctas, int(os.environ.get("TVM_AR_CTAS_OVERLAP", "16"))
            x.side_stream = self._side_stream.cuda_stream
            call = lambda: L.check(lib.tvm_backward_dp(C.byref(model), _ptr(rays), n, int(S), _ptr(jitter), flags, _ptr(rgb), _ptr(d_rgb),
                                                       _ptr(d_penalty), C.byref(gs), _ptr(ws), ws.numel(), C.byref(x), _stream_ptr()),
                                   "tvm_backward_dp")
            if not pipelined:
                x.phase = 0
                call()
                return self._unpack_grads(gp, items)
            # two halves, both exchanges left on the side stream: the caller (TrainStepGraph._body) runs the appearance tail
            # of the step while the density half is on the wire
            x.phase = 1
            call()
            ev_app = torch.cuda.Event()
            ev_app.record(self._side_stream)
            x.phase = 2
            call()
            ev_den = torch.cuda.Event()
            ev_den.record(self._side_stream)
            return gp, items, ev_app, ev_den
        L.check(lib.tvm_backward(C.byref(model), _ptr(rays), n, int(S), _ptr(jitter), flags, _ptr(rgb), _ptr(d_rgb),
                                 _ptr(d_penalty), C.byref(gs), _ptr(ws), ws.numel(), _stream_ptr()), "tvm_backward")
        if self.grad_sync:
            if self._peer_comm is not None:
                self._peer_comm.allreduce_(0, (total + 3) // 4 * 4)
            else:
                from .dist import allreduce_flat_
                allreduce_flat_(gp, group=self.grad_sync_group, average=False)
        return self._unpack_grads(gp, items)

    def _unpack_grads(self, gp, items, part=None):
        """part None: every gradient, list in _param_list() order; 'density' / 'app': only that half is unpacked (the other
        entries of the list are None)."""
        lib = L.load()
        st = _stream_ptr()
        base = gp.data_ptr()
        at = lambda name: C.c_void_p(base + 4 * items[name][0])
        Cd, Ca, F = self.density_n_comp[0], self.app_n_comp[0], self.featureC
        in_c = self.renderModule.in_mlpC
        out, jobs = {}, []
        for k in range(3):
            for pref, plist, llist, c in (("d", self.density_plane, self.density_line, Cd),
                                          ("a", self.app_plane, self.app_line, Ca)):
                if part is not None and pref != part[0]:
                    out[f"{pref}p{k}"], out[f"{pref}l{k}"] = None, None
                    continue
                gpl, gl = torch.empty_like(plist[k]), torch.empty_like(llist[k])
                # channels-last [H*W][C] -> NCHW [C][H*W] (tvm_unpack_grid), all 12 gradients in one launch
                jobs.append(L.TvmTransposeJob(at(f"{pref}p{k}").value, gpl.data_ptr(), gpl.shape[2] * gpl.shape[3], c))
                jobs.append(L.TvmTransposeJob(at(f"{pref}l{k}").value, gl.data_ptr(), gl.shape[2], c))
                out[f"{pref}p{k}"], out[f"{pref}l{k}"] = gpl, gl
        if part == "density":
            L.check(lib.tvm_transpose_batch((L.TvmTransposeJob * len(jobs))(*jobs), len(jobs), st), "tvm_transpose_batch")
            return [*(out[f"dp{k}"] for k in range(3)), *(out[f"dl{k}"] for k in range(3))] + [None] * (len(self._param_list()) - 6)
        m = self.renderModule.mlp
        J = L.TvmTransposeJob
        g_basis = torch.empty_like(self.basis_mat.weight)
        g_w1, g_w2 = torch.empty_like(m[0].weight), torch.empty_like(m[2].weight)
        g_b1, g_b2 = torch.empty_like(m[0].bias), torch.empty_like(m[2].bias)
        g_w3, g_b3 = torch.empty_like(m[4].weight), torch.empty_like(m[4].bias)
        # [in][out_pad] -> [out][in] (tvm_unpack_linear) and the bias / last-layer slices, in the same launch as the grids
        jobs += [J(at("basis_t").value, g_basis.data_ptr(), 3 * Ca, self.app_dim, self.head_dim(), 0),
                 J(at("w1_t").value, g_w1.data_ptr(), in_c, F, F, 0), J(at("w2_t").value, g_w2.data_ptr(), F, F, F, 0),
                 J(at("b1").value, g_b1.data_ptr(), 1, F), J(at("b2").value, g_b2.data_ptr(), 1, F),
                 J(at("w3").value, g_w3.data_ptr(), 1, 3 * F), J(at("b3").value, g_b3.data_ptr(), 1, 3)]
        L.check(lib.tvm_transpose_batch((J * len(jobs))(*jobs), len(jobs), st), "tvm_transpose_batch")
        return [*(out[f"dp{k}"] for k in range(3)), *(out[f"dl{k}"] for k in range(3)),
                *(out[f"ap{k}"] for k in range(3)), *(out[f"al{k}"] for k in range(3)), g_basis,
                g_w1, g_b1, g_w2, g_b2, g_w3, g_b3, *self._unpack_head_grads(gp, items)]

    def _unpack_head_grads(self, gp, items):
        """Gradients of the parameters `_param_list` appends after the MLP (none for TensorVMSplit)."""
        return []

    def _unpack_part(self, gp, items, part):
        """Unpack one half of the flat gradient buffer straight into .grad: part 'density' = density planes / lines,
        'app' = appearance planes / lines, basis_mat, MLP (and the REFTensoRF heads).  Returns the parameters it filled."""
        grads = self._unpack_grads(gp, items, part=part)
        params = self._param_list()
        sel = range(0, 6) if part == "density" else range(6, len(params))
        out = []
        for i in sel:
            params[i].grad = grads[i]
            out.append(params[i])
        return out

    # ---- reference API: the per-chunk call ----------------------------------------------------
    def forward(self, rays_chunk, white_bg=True, is_train=False, ndc_ray=False, N_samples=-1,
                additional_output=False, jitter=None):
        """TensorBase.execute (tensorBase.py:476-536): returns (rgb_map [n,3], depth_map [n])."""
        if ndc_ray:
            raise NotImplementedError("ndc_ray sampling is outside the hot path (SURVEY.md §2.1)")
        if additional_output:
            raise NotImplementedError("additional_output is only consumed by the NeRF++ variant")
        S = int(N_samples) if N_samples > 0 else self.nSamples
        rays = rays_chunk.contiguous()
        if is_train and jitter is None:
            jitter = torch.rand(rays.shape[0], dtype=torch.float32, device=rays.device)   # tensorBase.py:353
        if not is_train:
            jitter = None
        flags = self._flags(white_bg)
        n, nmax = rays.shape[0], self.max_rays_per_launch(S)
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._param_list())
        if needs_grad:
            if n > nmax:
                raise ValueError(f"training chunk of {n} rays exceeds the workspace budget ({nmax} rays)")
            rgb, depth, penalty = _RenderFn.apply(self, rays, jitter, flags, S, *self._param_list())
            if self.VARIANT == L.VARIANT_REF:
                self.penalty = penalty          # train.py:253-257 reads tensorf.penalty and adds it to the loss
            return rgb, depth
        if self.fused_composite:
            flags |= L.EVAL_ONLY       # no backward follows: the appearance head composites as it goes (tvmrender.h)
        if self._plan_launch(n, S) == (n, None):
            return self._forward_raw(rays, jitter, flags, S)
        rgb = torch.empty((n, 3), dtype=torch.float32, device=rays.device)
        depth = torch.empty((n,), dtype=torch.float32, device=rays.device)
        self._forward_chunks(rays, jitter, flags, S, rgb, depth)
        return rgb, depth

    execute = forward

    # ---- parity instrumentation -------------------------------------------------------------------
    def forward_with_aux(self, rays, white_bg=True, N_samples=-1, jitter=None, want_rgb=True):
        """tvm_forward with every per-sample output requested (tests only; disables ERT)."""
        S = int(N_samples) if N_samples > 0 else self.nSamples
        rays = rays.contiguous()
        n, NB = rays.shape[0], (S + 31) // 32
        dev = rays.device
        bits = lambda: torch.empty((n, NB), dtype=torch.int32, device=dev)
        o = dict(bbox_bits=bits(), valid_bits=bits(), app_bits=bits(),
                 sigma=torch.empty((n, S), dtype=torch.float32, device=dev),
                 weight=torch.empty((n, S), dtype=torch.float32, device=dev),
                 rgb=torch.empty((n, S, 3), dtype=torch.float32, device=dev) if want_rgb else None,
                 acc_map=torch.empty((n,), dtype=torch.float32, device=dev))
        aux = L.TvmAux()
        for k, v in o.items():
            setattr(aux, k, v.data_ptr() if v is not None else None)
        if self.VARIANT == L.VARIANT_REF:
            self._penalty_buf.zero_()
            aux.penalty = self._penalty_buf.data_ptr()
            self.penalty = self._penalty_buf
        rgb_map, depth_map = self._forward_raw(rays, jitter, self._flags(white_bg), S, aux=aux)
        o.update(rgb_map=rgb_map, depth_map=depth_map)
        return o

    def compute_alpha(self, xyz_locs, length=1):
        """tensorBase.py:451-473."""
        xyz = xyz_locs.contiguous().view(-1, 3)
        out = torch.empty(xyz.shape[0], dtype=torch.float32, device=xyz.device)
        model = self._model()
        L.check(L.load().tvm_density_alpha(C.byref(model), _ptr(xyz), xyz.shape[0], float(length), _ptr(out),
                                           _stream_ptr()), "tvm_density_alpha")
        return out.view(xyz_locs.shape[:-1])


class REFTensoRF(TensorVMSplit):
    """REFTensoRF (models/REFTensoRF.py:64-256): normal / diffuse / specular-tint / rho heads on the
    144-vector, reflected direction into MLPRender_Fea_Ref, side output `penalty` (train.py:253-257).
    Backward: k_app_bwd<48> (fp32) or k_app_bwd_tc<REF> (tensor cores, bf16 / fp16 modes)."""
    VARIANT = L.VARIANT_REF

    def init_render_func(self, shadingMode, pos_pe, view_pe, fea_pe, featureC, device):
        self.renderModule = MLPRender_Fea(self.app_dim, view_pe, fea_pe, featureC, extra_in=1).to(device)   # MLPRender_Fea_Ref
        self._penalty_buf = torch.zeros(1, dtype=torch.float32, device=device)     # accumulated by the kernels
        self.penalty = self._penalty_buf                                           # what train.py:253-257 reads

    def head_dim(self):
        return L.REF_HEAD_LD

    def init_svd_volume(self, res, device):
        super().init_svd_volume(res, device)
        k = sum(self.app_n_comp)
        self.normal_linear = _Linear(k, 3).to(device)
        self.diffuse_linear = _Linear(k, 3).to(device)
        self.specular_linear = _Linear(k, 1).to(device)
        self.rho_linear = _Linear(k, 1).to(device)

    def _heads(self):
        return [self.normal_linear, self.diffuse_linear, self.specular_linear, self.rho_linear]

    def get_optparam_groups(self, lr_init_spatialxyz=0.02, lr_init_network=0.001):
        g = super().get_optparam_groups(lr_init_spatialxyz, lr_init_network)
        return g + [{'params': h.parameters(), 'lr': lr_init_network} for h in
                    (self.normal_linear, self.diffuse_linear, self.rho_linear, self.specular_linear)]

    def _param_list(self):
        extra = []
        for h in self._heads():
            extra += [h.weight, h.bias]
        return super()._param_list() + extra

    def load_numpy_params(self, p):
        super().load_numpy_params(p)
        with torch.no_grad():
            for n in ("normal", "diffuse", "specular", "rho"):
                getattr(self, n + "_linear").weight.copy_(torch.from_numpy(p.extra[n + "_w"]))
                getattr(self, n + "_linear").bias.copy_(torch.from_numpy(p.extra[n + "_b"]))

    def _pack_heads(self, lib, at, items, st):
        Ca = self.app_n_comp[0]
        w = torch.cat([self.basis_mat.weight.detach()] + [h.weight.detach() for h in self._heads()], 0).contiguous()
        self._heads_w = w          # keep alive until the pack kernel has run
        L.check(lib.tvm_pack_linear(_ptr(w), w.shape[0], 3 * Ca, L.REF_HEAD_LD, at("basis_t"), st), "tvm_pack_linear")
        b = torch.cat([torch.zeros(self.app_dim, device=w.device)] + [h.bias.detach() for h in self._heads()])
        off = items["head_bias"][0]
        self._packed[off:off + 64].zero_()
        self._packed[off:off + b.numel()].copy_(b)
        return []

    def _unpack_head_grads(self, gp, items):
        """[3*Ca][48] packed head gradients -> normal / diffuse / specular / rho weights and biases (order of _param_list)."""
        lib, st = L.load(), _stream_ptr()
        Ca = self.app_n_comp[0]
        n_out = self.app_dim + 8
        full = torch.empty((n_out, 3 * Ca), dtype=torch.float32, device=gp.device)
        L.check(lib.tvm_unpack_linear(C.c_void_p(gp.data_ptr() + 4 * items["basis_t"][0]), n_out, 3 * Ca, L.REF_HEAD_LD,
                                      _ptr(full), st), "tvm_unpack_linear")
        hb = gp[items["head_bias"][0]:items["head_bias"][0] + n_out]
        out, o = [], self.app_dim
        for h in self._heads():
            k = h.weight.shape[0]
            out += [full[o:o + k].clone(), hb[o:o + k].clone()]
            o += k
        return out

    def _forward_raw(self, rays, jitter, flags, S, aux=None, out=None, ws_slot=0, ws_bytes=None):
        if aux is None:
            aux = L.TvmAux()
            if ws_slot == 0:        # chunked renders: the first chunk of a pair zeroes, both accumulate (atomics)
                self._penalty_buf.zero_()
            aux.penalty = self._penalty_buf.data_ptr()
            self.penalty = self._penalty_buf
        return super()._forward_raw(rays, jitter, flags, S, aux=aux, out=out, ws_slot=ws_slot, ws_bytes=ws_bytes)


class _Seq(torch.nn.ModuleList):
    """Positional container so that parameter names match the reference's nn.Sequential indices."""


class _BgNet(torch.nn.Module):
    """Parameter container mirroring MLPNet (models/nerfplusplus.py:66-113) for D=3, W=128, skips=[1]."""

    def __init__(self, pos_dim, dir_dim, D=3, W=128):
        super().__init__()
        skips = [int(D / 2)]
        layers, dim = [], pos_dim
        for i in range(D):
            layers.append(_Seq([_Linear(dim, W), torch.nn.Identity()]))
            dim = W
            if i in skips and i != D - 1:
                dim += pos_dim
        self.base_layers = torch.nn.ModuleList(layers)
        self.sigma_layers = _Seq([_Linear(dim, 1)])
        self.base_remap_layers = _Seq([_Linear(dim, 256)])
        self.rgb_layers = _Seq([_Linear(256 + dir_dim, W // 2), torch.nn.Identity(), _Linear(W // 2, 3), torch.nn.Identity()])


class NerfPlusPlus(TensorVMSplit):
    """NerfPlusPlus (models/nerfplusplus.py:143-318): sphere-bounded always-jittered foreground sampling plus a
    512-sample inverted-sphere background MLP (tvm_forward_npp / tvm_backward_npp); the U[0,1) draws of
    perturb_samples come from torch.rand on the device unless `fg_rand` / `bg_rand` are injected."""

    def set_nerfplusplus(self, bg_freq=4, bg_view_freq=2, bg_D=4, radii=20):
        if (bg_freq, bg_view_freq, bg_D) != (2, 2, 3):
            raise NotImplementedError("the background kernels are built for bg_freq=2, bg_view_freq=2, bg_D=3 "
                                      "(configs/Scarf.txt:12-15)")
        self.bg_freq, self.bg_view_freq, self.bg_D, self.radii = bg_freq, bg_view_freq, bg_D, radii
        self.bg_net = _BgNet(4 + 4 * bg_freq * 2, 3 + 3 * bg_view_freq * 2, bg_D).to(self.device)
        self._bg_packed = None
        self._bg_versions = None
        self._model_struct = None

    def get_optparam_groups(self, lr_init_spatialxyz=0.02, lr_init_network=0.001):
        return super().get_optparam_groups(lr_init_spatialxyz, lr_init_network) + \
            [{'params': self.bg_net.parameters(), 'lr': lr_init_network}]

    def load_numpy_params(self, p):
        super().load_numpy_params(p)
        e = p.extra
        if not hasattr(self, "bg_net"):
            self.set_nerfplusplus(e["bg_freq"], e["bg_view_freq"], e["bg_D"], e["radii"])
        with torch.no_grad():
            def cp(lin, wb):
                lin.weight.copy_(torch.from_numpy(wb[0]))
                lin.bias.copy_(torch.from_numpy(wb[1]))
            for i, wb in enumerate(e["bg_base"]):
                cp(self.bg_net.base_layers[i][0], wb)
            cp(self.bg_net.sigma_layers[0], e["bg_sigma"])
            cp(self.bg_net.base_remap_layers[0], e["bg_remap"])
            cp(self.bg_net.rgb_layers[0], e["bg_rgb0"])
            cp(self.bg_net.rgb_layers[2], e["bg_rgb1"])

    def _finish_model(self, s):
        s.sampling, s.radii = L.SAMPLING_NPP, float(self.radii)

    _BG_SIZES = dict(w0_t=20 * 128, b0=128, w1_t=128 * 128, b1=128, w2_t=148 * 128, b2=128, w_sigma=128, b_sigma=64,
                     wf_t=128 * 64, bf=64, wv_t=15 * 64, w_rgb=3 * 64, b_rgb=64)

    def _bg_layout(self):
        """Offsets (floats) of the packed background tensors; the gradient buffer uses the same layout."""
        offs, off = {}, 0
        for k, v in self._BG_SIZES.items():
            offs[k] = off
            off += (v + 63) // 64 * 64
        return self._BG_SIZES, offs, off

    def _bg_param_list(self):
        n = self.bg_net
        lins = [n.base_layers[0][0], n.base_layers[1][0], n.base_layers[2][0], n.sigma_layers[0], n.base_remap_layers[0],
                n.rgb_layers[0], n.rgb_layers[2]]
        return [t for lin in lins for t in (lin.weight, lin.bias)]

    def _forward_npp_raw(self, rays, fg_rand, bg_rand, flags, S, out=None, aux=None, ws_slot=0, ws_bytes=None):
        lib = L.load()
        n = rays.shape[0]
        model, bg = self._model(), self._bg_struct()
        ws = self._workspace(n, S, ws_slot, ws_bytes)
        self._fwd_gen += 1
        if out is None:
            out = (torch.empty((n, 3), dtype=torch.float32, device=rays.device),
                   torch.empty((n,), dtype=torch.float32, device=rays.device))
        rgb, depth = out
        L.check(lib.tvm_forward_npp(C.byref(model), C.byref(bg), _ptr(rays), n, S, _ptr(fg_rand), _ptr(bg_rand), flags,
                                    _ptr(rgb), _ptr(depth), C.byref(aux) if aux is not None else None,
                                    _ptr(self.counters) if self.collect_counters else None, _ptr(ws),
                                    ws.numel() if ws_bytes is None else int(ws_bytes), _stream_ptr()), "tvm_forward_npp")
        return rgb, depth

    def _backward_npp_raw(self, rays, fg_rand, bg_rand, flags, S, rgb, d_rgb):
        """tvm_backward_npp -> gradients in the order of _param_list() + _bg_param_list(), reference shapes."""
        lib = L.load()
        n = rays.shape[0]
        model, bg = self._model(), self._bg_struct()
        items, total = self._layout()
        sizes, offs, bg_total = self._bg_layout()
        if getattr(self, "_grads_packed", None) is None or self._grads_packed.numel() != total + bg_total:
            self._grads_packed = torch.empty(total + bg_total, dtype=torch.float32, device=self.device)
        gp = self._grads_packed
        gp.zero_()
        gs = self._struct_for(gp, L.TvmGrads)
        bgp = gp[total:]
        bgs = L.TvmBgGrads()
        for k in sizes:
            setattr(bgs, k, bgp.data_ptr() + 4 * offs[k])
        ws = self._workspace(n, S)
        L.check(lib.tvm_backward_npp(C.byref(model), C.byref(bg), _ptr(rays), n, int(S), _ptr(fg_rand), _ptr(bg_rand), flags,
                                     _ptr(rgb), _ptr(d_rgb), C.byref(gs), C.byref(bgs), _ptr(ws), ws.numel(),
                                     _stream_ptr()), "tvm_backward_npp")
        if self.grad_sync:
            from .dist import allreduce_flat_
            allreduce_flat_(gp, group=self.grad_sync_group, average=True)
        fg = self._unpack_grads(gp, items)
        st = _stream_ptr()
        at = lambda k: C.c_void_p(bgp.data_ptr() + 4 * offs[k])
        nb = self.bg_net
        out = []
        for i, (wk, bk) in enumerate((("w0_t", "b0"), ("w1_t", "b1"), ("w2_t", "b2"))):
            w = nb.base_layers[i][0].weight
            gw = torch.empty_like(w)
            L.check(lib.tvm_unpack_linear(at(wk), w.shape[0], w.shape[1], w.shape[0], _ptr(gw), st), "tvm_unpack_linear")
            out += [gw, bgp[offs[bk]:offs[bk] + 128].clone()]
        out += [bgp[offs["w_sigma"]:offs["w_sigma"] + 128].clone().reshape(1, 128), bgp[offs["b_sigma"]:offs["b_sigma"] + 1].clone()]
        r, g0 = nb.base_remap_layers[0], nb.rgb_layers[0]
        d_rw, d_rb = torch.empty_like(r.weight), torch.empty_like(r.bias)
        d_gw, d_gb = torch.empty_like(g0.weight), torch.empty_like(g0.bias)
        L.check(lib.tvm_bg_fold_bwd(_ptr(r.weight.detach()), _ptr(r.bias.detach()), _ptr(g0.weight.detach()), at("wf_t"),
                                    at("bf"), at("wv_t"), _ptr(d_rw), _ptr(d_rb), _ptr(d_gw), _ptr(d_gb), st),
                "tvm_bg_fold_bwd")
        out += [d_rw, d_rb, d_gw, d_gb, bgp[offs["w_rgb"]:offs["w_rgb"] + 192].clone().reshape(3, 64),
                bgp[offs["b_rgb"]:offs["b_rgb"] + 3].clone()]
        return [*fg, *out]

    def _bg_struct(self, force=False):
        """TvmBgNet over a packed fp32 buffer; re-packed (and re-folded) when a bg parameter changed (force: always -- the
        captured training step re-packs at its end, because a replay runs no host-side version check)."""
        lib = L.load()
        n = self.bg_net
        params = list(n.parameters())
        versions = tuple((p.data_ptr(), p._version) for p in params)
        sizes, offs, off = self._bg_layout()
        if self._bg_packed is None or versions != self._bg_versions or force:
            if self._bg_packed is None:
                self._bg_packed = torch.zeros(off, dtype=torch.float32, device=self.device)
            buf, st = self._bg_packed, _stream_ptr()
            at = lambda k: C.c_void_p(buf.data_ptr() + 4 * offs[k])
            for i, (wk, bk) in enumerate((("w0_t", "b0"), ("w1_t", "b1"), ("w2_t", "b2"))):
                lin = n.base_layers[i][0]
                w = lin.weight.detach()
                L.check(lib.tvm_pack_linear(_ptr(w), w.shape[0], w.shape[1], w.shape[0], at(wk), st), "tvm_pack_linear")
                buf[offs[bk]:offs[bk] + 128].copy_(lin.bias.detach())
            buf[offs["w_sigma"]:offs["w_sigma"] + 128].copy_(n.sigma_layers[0].weight.detach().reshape(-1))
            buf[offs["b_sigma"]:offs["b_sigma"] + 1].copy_(n.sigma_layers[0].bias.detach())
            r, g0, g1 = n.base_remap_layers[0], n.rgb_layers[0], n.rgb_layers[2]
            L.check(lib.tvm_bg_fold(_ptr(r.weight.detach()), _ptr(r.bias.detach()), _ptr(g0.weight.detach()),
                                    _ptr(g0.bias.detach()), at("wf_t"), at("bf"), at("wv_t"), st), "tvm_bg_fold")
            buf[offs["w_rgb"]:offs["w_rgb"] + 192].copy_(g1.weight.detach().reshape(-1))
            buf[offs["b_rgb"]:offs["b_rgb"] + 3].copy_(g1.bias.detach())
            s = L.TvmBgNet()
            for k in sizes:
                setattr(s, k, buf.data_ptr() + 4 * offs[k])
            s.tc_weights = None
            self._bg_struct_c, self._bg_versions = s, versions
            self._bg_tc_stale = True
        s = self._bg_struct_c
        if self.mlp_mode != "fp32" and self._bg_tc_stale:
            # bf16 operand image of the background network for the tcgen05 kernel (k_bg_tc)
            if getattr(self, "_bg_tc", None) is None:
                self._bg_tc = torch.empty(lib.tvm_bg_tc_bytes(), dtype=torch.uint8, device=self.device)
            L.check(lib.tvm_pack_bg_tc(C.byref(s), _ptr(self._bg_tc), _stream_ptr()), "tvm_pack_bg_tc")
            s.tc_weights = self._bg_tc.data_ptr()
            self._bg_tc_stale = False
        return s

    def forward(self, rays_chunk, white_bg=False, is_train=False, ndc_ray=False, N_samples=-1,
                additional_output=True, fg_rand=None, bg_rand=None, aux=None):
        if ndc_ray:
            raise NotImplementedError("ndc_ray sampling is outside the hot path")
        S = int(N_samples) if N_samples > 0 else self.nSamples
        rays = rays_chunk.contiguous()
        n, nmax = rays.shape[0], self.max_rays_per_launch(S)
        dev = rays.device
        if fg_rand is None:
            fg_rand = torch.rand((n, S), dtype=torch.float32, device=dev)       # perturb_samples, :204
        if bg_rand is None:
            bg_rand = torch.rand((n, 512), dtype=torch.float32, device=dev)
        fg_rand, bg_rand = fg_rand.contiguous(), bg_rand.contiguous()
        flags = self._flags(False)
        params = [*self._param_list(), *self._bg_param_list()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            if n > nmax:
                raise ValueError(f"training chunk of {n} rays exceeds the workspace budget ({nmax} rays)")
            assert aux is None
            return _RenderNppFn.apply(self, rays, fg_rand, bg_rand, flags, S, *params)
        rgb = torch.empty((n, 3), dtype=torch.float32, device=dev)
        depth = torch.empty((n,), dtype=torch.float32, device=dev)
        if aux is not None:           # parity instrumentation: per-sample outputs of ONE launch
            assert self._plan_launch(n, S) == (n, None)
            self._forward_npp_raw(rays, fg_rand, bg_rand, flags, S, out=(rgb, depth), aux=aux)
            return rgb, depth

        out_ref = (weakref.ref(rgb), weakref.ref(depth))        # see TensorVMSplit._forward_chunks

        def launch(s, e, slot, nbytes):
            o_rgb, o_depth = out_ref[0](), out_ref[1]()
            if o_rgb is None or o_depth is None:
                return None
            self._forward_npp_raw(rays[s:e], fg_rand[s:e], bg_rand[s:e], flags, S, out=(o_rgb[s:e], o_depth[s:e]),
                                  ws_slot=slot, ws_bytes=nbytes)
            return self._ws if slot == 0 else self._ws2
        self._render_bounded(n, S, launch)
        return rgb, depth

    execute = forward

    def forward_with_aux(self, rays, N_samples=-1, fg_rand=None, bg_rand=None, **_):
        """Parity instrumentation: per-sample masks of the foreground plus bg_lambda / bg_rgb_map."""
        S = int(N_samples) if N_samples > 0 else self.nSamples
        n, NB, dev = rays.shape[0], (S + 31) // 32, rays.device
        bits = lambda: torch.empty((n, NB), dtype=torch.int32, device=dev)
        o = dict(bbox_bits=bits(), valid_bits=bits(), app_bits=bits(),
                 sigma=torch.empty((n, S), dtype=torch.float32, device=dev),
                 weight=torch.empty((n, S), dtype=torch.float32, device=dev),
                 bg_lambda=torch.zeros(n, dtype=torch.float32, device=dev),
                 bg_rgb_map=torch.zeros((n, 3), dtype=torch.float32, device=dev))
        aux = L.TvmAux()
        for k, v in o.items():
            setattr(aux, k, v.data_ptr())
        with torch.no_grad():
            rgb_map, depth_map = self.forward(rays, N_samples=S, fg_rand=fg_rand, bg_rand=bg_rand, aux=aux)
        o.update(rgb_map=rgb_map, depth_map=depth_map)
        return o


def model_from_params(p, device="cuda:0", alpha_volume=None, alpha_aabb=None, mlp_mode="fp32"):
    """TensorVMSplit / REFTensoRF from a parameter record with the reference's shapes (e.g. synthetic.ModelParams)."""
    dev = torch.device(device)
    cls = {"ref": REFTensoRF, "npp": NerfPlusPlus}.get(getattr(p, "extra", {}).get("variant"), TensorVMSplit)
    m = cls(p.aabb, p.gridSize, dev, density_n_comp=list(p.density_n_comp),
                      appearance_n_comp=list(p.app_n_comp), app_dim=p.app_dim, near_far=list(p.near_far),
                      shadingMode="MLP_Fea", density_shift=p.density_shift, distance_scale=p.distance_scale,
                      rayMarch_weight_thres=p.rayMarch_weight_thres, view_pe=p.view_pe, fea_pe=p.fea_pe,
                      featureC=p.featureC, step_ratio=p.step_ratio, fea2denseAct=p.fea2denseAct)
    m.load_numpy_params(p)
    if alpha_volume is not None:
        m.alphaMask = AlphaGridMask(dev, alpha_aabb if alpha_aabb is not None else p.aabb, alpha_volume)
    m.mlp_mode = mlp_mode
    return m


def unpack_bits(bits, S: int) -> np.ndarray:
    """[n, NB] int32 words (device tensor or numpy) -> [n, S] bool numpy (bit j of word b = sample 32*b + j)."""
    w = (bits.detach().cpu().numpy() if torch.is_tensor(bits) else np.ascontiguousarray(bits)).view(np.uint32)
    b = np.unpackbits(w.view(np.uint8).reshape(w.shape[0], -1), axis=1, bitorder="little")
    return b[:, :S].astype(bool)
