"""N-rank (NCCL) run of the hot path against the oracle; needs >= 2 visible GPUs (skipped on the 1-GPU test box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_hot_path_matches_oracle():
    try:
        n = int(subprocess.check_output(["nvidia-smi", "-L"], text=True).count("GPU "))
    except Exception:
        n = 0
    if n < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29611",
           os.path.join(ROOT, "tests", "mp", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTI_GPU_CHECK world=2 OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
