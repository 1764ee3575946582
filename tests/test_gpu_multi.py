"""World-size-2 NCCL check of the sharded loss and the library-owned communicator (needs two GPUs on
the box; skipped otherwise -- the gloo world-2 test in test_host.py covers the host logic on CPU)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_loss_nccl_world2_library_communicator():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.join(HERE, "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    text = out.stdout + out.stderr
    assert out.returncode == 0, text[-2000:]
    assert text.count("-> OK") == 2 and "MISMATCH" not in text, text[-2000:]
    assert "lib comm True" in text, text[-2000:]
