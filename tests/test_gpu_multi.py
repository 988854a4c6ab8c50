"""-m gpu multi-GPU test (needs >= 2 GPUs; skipped otherwise): a 2-rank ZeRO-2 data-parallel step equals
the 1-GPU step on the same global batch (loss and updated weights <= 1e-3 relative)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_zero2_step_matches_single_gpu():
    port = str(29500 + os.getpid() % 500)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", port, os.path.join(REPO, "tools", "gpu_check_zero2.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "ZERO2_OK" in r.stdout
