"""Multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box).  Spawns torchrun on tools/sharded_check.py,
which compares ShardedSimulator (CUDA IPC peer swaps and the NCCL fallback) with the CPU oracle."""
import os
import subprocess
import sys

import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_gpu_parity(exchange):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(H.ROOT, "tools", "sharded_check.py"), exchange, "24"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "sharded check ok" in out.stdout
