"""Multi-GPU (one process per GPU, NCCL halo exchange) parity: runs only when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("fuse", [False, True])
def test_sharded_over_nccl_equals_single(fuse):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "run_sharded_nccl.py")]
    env = dict(os.environ)
    if fuse:        # fused E+H stepping on every shard (double-buffered halo planes)
        env.update(ALISB_TEST_FUSE="1", ALISB_FUSE_MIN="1", ALISB_SMALL="1", ALISB_FUSE_FRAC="0.01")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert "SHARDED_NCCL_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
