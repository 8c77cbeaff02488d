"""ctypes binding of libtvmrender.so (include/tvmrender.h).

The shared library is the product; this file only marshals pointers.  There is deliberately no
fallback: if the library is missing or no CUDA device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TVM_LIB: developer override (kernel-tuning experiments build variant libraries next to the default one)
LIB_PATH = os.environ.get("TVM_LIB") or os.path.join(_HERE, "libtvmrender.so")
ABI_VERSION = 29

# flags (tvmrender.h)
WHITE_BG = 0x1
NO_ERT = 0x2
EVAL_ONLY = 0x4
MLP_FP32 = 0x0
MLP_BF16 = 0x10
MLP_FP16 = 0x20
ACT_SOFTPLUS, ACT_RELU = 0, 1
VARIANT_VM, VARIANT_REF, REF_HEAD_LD = 0, 1, 48
CNT_M_IN, CNT_M_V, CNT_M_A, CNT_RAYS, CNT_WORDS = 0, 1, 2, 3, 8
CNT_BG_RAYS, CNT_BG_SAMPLES = 4, 5
STAGE_NAMES = ["march", "app", "composite", "bwd_app", "bwd_march", "bg", "bwd_bg"]
SAMPLING_UNIFORM, SAMPLING_NPP = 0, 1
STAGE_COUNT = 8

_f3 = C.c_float * 3
_f6 = C.c_float * 6
_i3 = C.c_int32 * 3
_p3 = C.c_void_p * 3


class TvmModel(C.Structure):
    _fields_ = [
        ("aabb", _f6), ("inv_aabb_size", _f3), ("grid", _i3),
        ("step_size", C.c_float), ("near_", C.c_float), ("far_", C.c_float),
        ("density_shift", C.c_float), ("distance_scale", C.c_float), ("weight_thres", C.c_float),
        ("act", C.c_int32), ("n_density", C.c_int32), ("n_app", C.c_int32), ("app_dim", C.c_int32),
        ("view_pe", C.c_int32), ("fea_pe", C.c_int32), ("feature_c", C.c_int32),
        ("density_plane", _p3), ("density_line", _p3), ("app_plane", _p3), ("app_line", _p3),
        ("variant", C.c_int32), ("basis_t", C.c_void_p), ("head_bias", C.c_void_p), ("w1_t", C.c_void_p), ("b1", C.c_void_p), ("w2_t", C.c_void_p),
        ("b2", C.c_void_p), ("w3", C.c_void_p), ("b3", C.c_void_p),
        ("alpha_bits", C.c_void_p), ("alpha_grid", _i3), ("alpha_aabb_min", _f3), ("alpha_inv_size", _f3),
        ("alpha_bricks", C.c_void_p), ("alpha_dilated", C.c_void_p), ("tc_weights", C.c_void_p), ("sampling", C.c_int32), ("radii", C.c_float),
        ("app_plane_pair", _p3), ("app_line_pair", _p3), ("tc_weights_bwd", C.c_void_p), ("alpha_bricks3", C.c_void_p),
    ]


class TvmAux(C.Structure):
    _fields_ = [("bbox_bits", C.c_void_p), ("valid_bits", C.c_void_p), ("app_bits", C.c_void_p),
                ("sigma", C.c_void_p), ("weight", C.c_void_p), ("rgb", C.c_void_p), ("acc_map", C.c_void_p),
                ("bg_lambda", C.c_void_p), ("bg_rgb_map", C.c_void_p), ("penalty", C.c_void_p)]


class TvmWorkspaceLayout(C.Structure):
    _fields_ = [("n_entries", C.c_size_t), ("blk_mask", C.c_size_t), ("blk_base", C.c_size_t), ("ent", C.c_size_t),
                ("ent_w", C.c_size_t), ("ent_rgb", C.c_size_t), ("acc", C.c_size_t), ("rgb_sum", C.c_size_t),
                ("capacity", C.c_uint32), ("n_blocks", C.c_int32), ("bytes", C.c_size_t)]


AR_MAX_WORLD, AR_MAX_CTAS = 16, 128


class TvmPeerComm(C.Structure):
    _fields_ = [("bufs", C.c_void_p * AR_MAX_WORLD), ("signals", C.c_void_p * AR_MAX_WORLD), ("multicast", C.c_void_p),
                ("epoch_dev", C.c_void_p), ("rank", C.c_int32), ("world", C.c_int32)]


class TvmGradExchange(C.Structure):
    _fields_ = [("comm", C.POINTER(TvmPeerComm)), ("split_floats", C.c_size_t), ("total_floats", C.c_size_t),
                ("n_ctas", C.c_int32), ("n_ctas_overlapped", C.c_int32), ("side_stream", C.c_void_p), ("phase", C.c_int32)]


class TvmBgNet(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w0_t", "b0", "w1_t", "b1", "w2_t", "b2", "w_sigma", "b_sigma", "wf_t", "bf",
                                           "wv_t", "w_rgb", "b_rgb", "tc_weights")]


BG_FIELDS = ("w0_t", "b0", "w1_t", "b1", "w2_t", "b2", "w_sigma", "b_sigma", "wf_t", "bf", "wv_t", "w_rgb", "b_rgb")


class TvmBgGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in BG_FIELDS]


class TvmTransposeJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32),
                ("src_ld", C.c_int32), ("dst_ld", C.c_int32)]


class TvmTvJob(C.Structure):
    _fields_ = [("plane_nchw", C.c_void_p), ("grad_nchw", C.c_void_p), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("weight", C.c_float), ("weight_dev", C.c_void_p), ("overwrite", C.c_int32)]


class TvmAdamTensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_size_t), ("lr", C.c_float), ("lr_index", C.c_int32)]


ADAM_MAX_TENSORS = 32


class TvmGrads(C.Structure):
    _fields_ = [("density_plane", _p3), ("density_line", _p3), ("app_plane", _p3), ("app_line", _p3),
                ("basis_t", C.c_void_p), ("head_bias", C.c_void_p), ("w1_t", C.c_void_p), ("b1", C.c_void_p), ("w2_t", C.c_void_p),
                ("b2", C.c_void_p), ("w3", C.c_void_p), ("b3", C.c_void_p)]


EXPORTS = [
    "tvm_last_error", "tvm_abi_version", "tvm_device_count", "tvm_pack_grid", "tvm_unpack_grid",
    "tvm_pack_linear", "tvm_unpack_linear", "tvm_transpose_batch", "tvm_pack_pair16", "tvm_pack_alpha", "tvm_pack_alpha_bricks", "tvm_pack_alpha_bricks3", "tvm_pack_alpha_dilated", "tvm_tc_weights_bytes", "tvm_pack_mlp_tc",
    "tvm_workspace_bytes", "tvm_workspace_bytes_bounded", "tvm_workspace_capacity", "tvm_forward_entries", "tvm_workspace_layout", "tvm_forward", "tvm_forward_npp", "tvm_bg_fold", "tvm_bg_tc_bytes", "tvm_pack_bg_tc", "tvm_backward", "tvm_backward_npp", "tvm_bg_fold_bwd",
    "tvm_density_alpha", "tvm_mse_loss",
    "tvm_profile_enable", "tvm_profile_collect",
    "tvm_dense_alpha", "tvm_alpha_mask_from_dense", "tvm_filter_rays", "tvm_generate_rays", "tvm_upsample_grid", "tvm_upsample_grids",
    "tvm_tv_loss", "tvm_tv_loss_batch", "tvm_l1_loss", "tvm_vector_diffs", "tvm_adam_step", "tvm_selftest_umma", "tvm_bench_gather",
    "tvm_allreduce_signal_words", "tvm_allreduce_sum", "tvm_backward_dp",
]


class TvmError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """dlopen libtvmrender.so; raises (never falls back) when it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TvmError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       f"or `make -C {os.path.join(_HERE, 'csrc')}`; there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise TvmError(f"libtvmrender.so does not export {name}")
    vp, i32, u32, f32 = C.c_void_p, C.c_int, C.c_uint32, C.c_float
    lib.tvm_last_error.restype = C.c_char_p
    lib.tvm_last_error.argtypes = []
    lib.tvm_abi_version.restype = i32
    lib.tvm_device_count.restype = i32
    lib.tvm_pack_grid.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.tvm_unpack_grid.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.tvm_pack_linear.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.tvm_transpose_batch.argtypes = [C.POINTER(TvmTransposeJob), i32, vp]
    lib.tvm_unpack_linear.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.tvm_pack_alpha.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.tvm_pack_pair16.argtypes = [vp, i32, i32, i32, vp, u32, vp]
    lib.tvm_pack_alpha_bricks.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.tvm_pack_alpha_dilated.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.tvm_pack_alpha_bricks3.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.tvm_tc_weights_bytes.restype = C.c_size_t
    lib.tvm_tc_weights_bytes.argtypes = [C.POINTER(TvmModel)]
    lib.tvm_bg_tc_bytes.restype = C.c_size_t
    lib.tvm_bg_tc_bytes.argtypes = []
    lib.tvm_pack_mlp_tc.argtypes = [C.POINTER(TvmModel), vp, u32, vp]
    lib.tvm_workspace_bytes.argtypes = [i32, i32, C.POINTER(C.c_size_t)]
    lib.tvm_workspace_layout.argtypes = [i32, i32, C.POINTER(TvmWorkspaceLayout)]
    lib.tvm_workspace_bytes_bounded.argtypes = [i32, i32, u32, C.POINTER(C.c_size_t)]
    lib.tvm_workspace_capacity.argtypes = [i32, i32, C.c_size_t, C.POINTER(u32)]
    lib.tvm_forward_entries.argtypes = [vp, vp, C.POINTER(u32)]
    lib.tvm_forward.argtypes = [C.POINTER(TvmModel), vp, i32, i32, vp, u32, vp, vp, C.POINTER(TvmAux), vp, vp,
                                C.c_size_t, vp]
    lib.tvm_forward_npp.argtypes = [C.POINTER(TvmModel), C.POINTER(TvmBgNet), vp, i32, i32, vp, vp, u32, vp, vp,
                                    C.POINTER(TvmAux), vp, vp, C.c_size_t, vp]
    lib.tvm_bg_fold.argtypes = [vp] * 8
    lib.tvm_pack_bg_tc.argtypes = [C.POINTER(TvmBgNet), vp, vp]
    i3 = C.POINTER(C.c_int32)
    lib.tvm_dense_alpha.argtypes = [C.POINTER(TvmModel), i3, f32, vp, vp]
    lib.tvm_alpha_mask_from_dense.argtypes = [vp, i3, f32, vp, vp, vp, vp, vp]
    lib.tvm_filter_rays.argtypes = [C.POINTER(TvmModel), vp, i32, i32, i32, vp, vp, vp]
    lib.tvm_generate_rays.argtypes = [C.POINTER(C.c_float), i32, i32, f32, f32, f32, f32, i32, i32, vp, vp]
    lib.tvm_upsample_grid.argtypes = [vp, i32, i32, i32, vp, i32, i32, vp]
    lib.tvm_upsample_grids.argtypes = [i32, C.POINTER(C.c_void_p), i3, C.POINTER(C.c_void_p), i3, vp]
    lib.tvm_tv_loss.argtypes = [vp, i32, i32, i32, f32, vp, vp, vp, vp]
    lib.tvm_tv_loss_batch.argtypes = [C.POINTER(TvmTvJob), i32, vp, vp]
    lib.tvm_l1_loss.argtypes = [vp, C.c_size_t, f32, vp, vp, vp, vp]
    lib.tvm_vector_diffs.argtypes = [vp, i32, i32, f32, vp, vp, vp, vp]
    lib.tvm_selftest_umma.argtypes = [vp] * 7
    lib.tvm_bench_gather.argtypes = [vp, C.c_size_t, i32, i32, vp, vp]
    lib.tvm_adam_step.argtypes = [C.POINTER(TvmAdamTensor), i32, f32, f32, f32, i32, vp, vp]
    lib.tvm_backward.argtypes = [C.POINTER(TvmModel), vp, i32, i32, vp, u32, vp, vp, vp, C.POINTER(TvmGrads), vp,
                                 C.c_size_t, vp]
    lib.tvm_backward_dp.argtypes = [C.POINTER(TvmModel), vp, i32, i32, vp, u32, vp, vp, vp, C.POINTER(TvmGrads), vp,
                                    C.c_size_t, C.POINTER(TvmGradExchange), vp]
    lib.tvm_backward_npp.argtypes = [C.POINTER(TvmModel), C.POINTER(TvmBgNet), vp, i32, i32, vp, vp, u32, vp, vp,
                                     C.POINTER(TvmGrads), C.POINTER(TvmBgGrads), vp, C.c_size_t, vp]
    lib.tvm_bg_fold_bwd.argtypes = [vp] * 11
    lib.tvm_density_alpha.argtypes = [C.POINTER(TvmModel), vp, i32, f32, vp, vp]
    lib.tvm_mse_loss.argtypes = [vp, vp, i32, f32, vp, vp, vp]
    lib.tvm_allreduce_signal_words.argtypes = [i32, C.POINTER(C.c_size_t)]
    lib.tvm_allreduce_sum.argtypes = [C.POINTER(TvmPeerComm), C.c_size_t, C.c_size_t, i32, vp]
    lib.tvm_profile_enable.argtypes = [i32]
    lib.tvm_profile_collect.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int)]
    for name in EXPORTS:
        if name not in ("tvm_last_error", "tvm_tc_weights_bytes", "tvm_bg_tc_bytes"):
            getattr(lib, name).restype = i32
    if lib.tvm_abi_version() != ABI_VERSION:
        raise TvmError(f"libtvmrender.so ABI {lib.tvm_abi_version()} != binding ABI {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().tvm_last_error().decode("utf-8", "replace")
        raise TvmError(f"{what} failed (rc={rc}): {msg}")


def require_cuda():
    """Fail loudly when there is no GPU: the product path has no CPU implementation."""
    lib = load()
    n = lib.tvm_device_count()
    if n <= 0:
        raise TvmError("no CUDA device visible to libtvmrender.so "
                       f"({lib.tvm_last_error().decode('utf-8', 'replace')}); there is no CPU fallback")
    return n


def profile_enable(on: bool):
    check(load().tvm_profile_enable(1 if on else 0), "tvm_profile_enable")


def profile_collect():
    """-> ({stage: ms}, {stage: launches}) accumulated since the last collect."""
    ms = (C.c_float * STAGE_COUNT)()
    cnt = (C.c_int * STAGE_COUNT)()
    check(load().tvm_profile_collect(ms, cnt), "tvm_profile_collect")
    return ({n: float(ms[i]) for i, n in enumerate(STAGE_NAMES)}, {n: int(cnt[i]) for i, n in enumerate(STAGE_NAMES)})
