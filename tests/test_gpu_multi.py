"""-m gpu: the NCCL path, one process per GPU (skipped on single-GPU boxes, where the same
kernels and message layout are covered by the loopback transport in test_gpu_parity.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("n", [2, 4, 8])
def test_nccl_exchange_parity(n):
    if _ngpus() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + n), os.path.join(ROOT, "tests", "mp_parity.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
