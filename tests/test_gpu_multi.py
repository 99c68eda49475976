"""Multi-GPU (one process per GPU, NCCL halo exchange) parity: runs only when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_over_nccl_equals_single():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "run_sharded_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "SHARDED_NCCL_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
