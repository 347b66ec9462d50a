"""Multi-GPU parity of the fused tensor-parallel exchange (csrc/tp_allreduce.cu): needs >= 2 GPUs
on the box (skipped otherwise).  Runs tools/tp_check.py under torchrun: integer-valued inputs, so
the cross-rank sum must equal NCCL's exactly and the residual exactly (the reference's
tests/distributed/test_custom_all_reduce.py:30-51 uses the same trick); the normed output is
checked against the oracle's fused_add_rms_norm."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_tp_exchange(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
           os.path.join(ROOT, "tools", "tp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "TP CHECK PASS" in r.stdout, r.stdout[-3000:]
