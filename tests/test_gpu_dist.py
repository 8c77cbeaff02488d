"""Multi-GPU checks on real devices (one rank per GPU under torch.distributed.run): sharded render, data-parallel gradients
through NCCL and through libtvmrender's own peer-memory all-reduce, the captured DP step.  Skips below two devices (the
driver's single-GPU `pytest -m gpu` box); the host-side logic is covered on CPU by tests/test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dp_check_on_all_visible_gpus(built_lib):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 CUDA devices")
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "scripts", "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(out.stdout[-2000:], out.stderr[-3000:])
    assert out.returncode == 0 and "dp_check OK" in out.stdout
