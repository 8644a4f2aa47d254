"""Multi-GPU parity: the row-partitioned path (NCCL ghost rows, halo, allreduce) against the single-GPU
path on the same inputs, launched with torchrun.  Needs >= 2 GPUs on the box; on a 1-GPU box the
row-partitioned HOST logic is still covered by tests/test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_partitioned_matches_single_gpu(iife):
    n = iife.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (host logic covered by test_dist_gloo.py)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "dist_check.py"), "12"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "dist_check ok" in out.stdout
