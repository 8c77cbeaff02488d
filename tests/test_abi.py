"""The C-ABI library loads and exports every symbol include/tvmrender.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "tvmrender.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tvm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    syms = _declared_symbols()
    assert len(syms) >= 14
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(built_lib._lib.EXPORTS) == syms


def test_abi_version_and_struct_sizes(built_lib):
    lib = built_lib._lib.load()
    assert lib.tvm_abi_version() == built_lib._lib.ABI_VERSION
    # the ctypes mirrors must match the C layout: compile a tiny probe with the real header
    import subprocess, tempfile
    probe = r'''
#include <stdio.h>
#include "tvmrender.h"
int main(){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(TvmModel), sizeof(TvmAux), sizeof(TvmGrads), sizeof(TvmBgNet),
                   sizeof(TvmBgGrads), sizeof(TvmTransposeJob), sizeof(TvmTvJob), sizeof(TvmAdamTensor),
                   sizeof(TvmWorkspaceLayout)); printf("%zu %zu\n", sizeof(TvmPeerComm), sizeof(TvmGradExchange)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(probe)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "p"),
                               os.path.join(d, "p.c")])
        out = subprocess.check_output([os.path.join(d, "p")]).split()
    L = built_lib._lib
    assert [int(x) for x in out] == [ctypes.sizeof(t) for t in (L.TvmModel, L.TvmAux, L.TvmGrads, L.TvmBgNet, L.TvmBgGrads,
                                                                  L.TvmTransposeJob, L.TvmTvJob, L.TvmAdamTensor,
                                                                  L.TvmWorkspaceLayout, L.TvmPeerComm, L.TvmGradExchange)]


def test_fails_loudly_without_gpu(built_lib):
    """On a box without a CUDA device the product path must raise, not fall back."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built_lib.TvmError):
        built_lib.TensorVMSplit([[-1, -1, -1], [1, 1, 1]], [8, 8, 8], "cuda", shadingMode="MLP_Fea")


def test_workspace_bytes(built_lib):
    lib = built_lib._lib.load()
    out = ctypes.c_size_t(0)
    assert lib.tvm_workspace_bytes(4096, 440, ctypes.byref(out)) == 0
    assert out.value >= 4096 * 440 * 24
    assert lib.tvm_workspace_bytes(0, 440, ctypes.byref(out)) != 0
    assert b"bad arguments" in lib.tvm_last_error()
    # tvm_workspace_layout: the members a caller may read back lie inside the workspace, 256-byte aligned, in carve order
    lay = built_lib._lib.TvmWorkspaceLayout()
    assert lib.tvm_workspace_layout(4096, 440, ctypes.byref(lay)) == 0
    assert lib.tvm_workspace_bytes(4096, 440, ctypes.byref(out)) == 0 and lay.bytes == out.value
    assert lay.capacity == 4096 * 440 and lay.n_blocks == 14 and lay.n_entries == 0
    offs = [lay.n_entries, lay.blk_mask, lay.blk_base, lay.ent, lay.ent_w, lay.ent_rgb, lay.acc, lay.rgb_sum]
    assert all(o % 256 == 0 and o < lay.bytes for o in offs) and sorted(offs[:5]) == offs[:5]
    assert lay.blk_base - lay.blk_mask >= 4096 * 14 * 4 and lay.ent_w - lay.ent >= lay.capacity * 8
    # bounded workspaces: bytes(entries) and capacity(bytes) are inverse up to the 256-byte padding; the worst case is the cap
    cap = ctypes.c_uint32(0)
    assert lib.tvm_workspace_capacity(4096, 440, out.value, ctypes.byref(cap)) == 0 and cap.value == 4096 * 440
    assert lib.tvm_workspace_capacity(4096, 440, out.value * 2, ctypes.byref(cap)) == 0 and cap.value == 4096 * 440
    for entries in (4096, 100000, 4096 * 440 - 1):
        b = ctypes.c_size_t(0)
        assert lib.tvm_workspace_bytes_bounded(4096, 440, entries, ctypes.byref(b)) == 0 and b.value <= out.value
        assert lib.tvm_workspace_capacity(4096, 440, b.value, ctypes.byref(cap)) == 0
        assert entries <= cap.value <= entries + 64
        assert lib.tvm_workspace_capacity(4096, 440, b.value - 256 * 6, ctypes.byref(cap)) == 0 and cap.value < entries
    assert lib.tvm_workspace_bytes_bounded(4096, 440, 0, ctypes.byref(b)) == 0        # floor: one entry per ray
    assert lib.tvm_workspace_capacity(4096, 440, b.value, ctypes.byref(cap)) == 0 and 4096 <= cap.value <= 4096 + 64
    assert lib.tvm_workspace_capacity(4096, 440, 1024, ctypes.byref(cap)) == 0 and cap.value == 0
