"""-m gpu: column-panel pipelining and the deferred / background scatter-add policies on the
loopback transport (rank threads on one GPU) and, where the box has the GPUs, one process per GPU over real
NVLink. The same scenarios run on the kernel-logic library in the CPU suite (tests/test_kernel_logic_emu.py)."""
import numpy as np
import pytest

import cm_b200 as cm
from scenario import Cfg, assert_blocks_match, run_checker, run_cuda, slice_steps

# a disagreement between rank threads would be a hang, not a failure: bound every test (pytest-timeout)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.mark.parametrize("grid,mode,npanel", [((2, 2, 64, 64, 1024, 2000, 1000), 2, 4), ((2, 4, 16, 16, 256, 500, 900), 2, 4),
                                              ((3, 2, 8, 4, 96, 250, 150), 2, 3)])
def test_column_panels_wire_bit_exact(grid, mode, npanel):
    """commits cut into column panels (thresholds lowered through the environment so
    that small commits are cut too): local blocks match the oracle and, index sets being sorted, the
    per-owner wire streams are the reference's element for element"""
    from emu_cases import panels_wire_and_blocks
    panels_wire_and_blocks(grid, mode, npanel)


@pytest.mark.parametrize("mode,budget,policy", [(2, 0, 1), (2, 40000, 1), (2, 0, 2), (2, 40000, 2)])
def test_deferred_scatter_add(mode, budget, policy):
    """cm_set_apply_policy(CM_APPLY_DEFERRED / CM_APPLY_BACKGROUND): commits deliver, the log is applied on
    read / over budget / when an arena grows (policy 2: also in the background, on the apply stream,
    while later commits pack); same local blocks as the eager oracle"""
    from emu_cases import deferred_apply
    deferred_apply(mode, budget, policy)


def test_deferred_scatter_add_api():
    from emu_cases import deferred_apply_api
    deferred_apply_api()


# ---- the same variants over real NVLink / NCCL, one process per GPU (skipped on single-GPU boxes) ----
MP_VARIANTS = {
    "batch": {"CM_B200_PIPELINE": "0"},
    "panels": {"CM_B200_PIPELINE": "1", "CM_B200_PIPELINE_MIN_ELEMS": "1"},
    "deferred": {"CM_TEST_APPLY_POLICY": "1"},
    "background": {"CM_TEST_APPLY_POLICY": "2", "CM_B200_BG_MIN_ELEMS": "1"},
    "nccl": {"CM_TEST_EXCHANGE_MODE": "1"},
    # threshold flushes inside update() with several ranks: flush regions mapped with CUDA IPC, stored into over NVLink
    "flush": {"CM_TEST_FLUSH_THRESHOLD": "50000"},
    "flush-nccl": {"CM_TEST_FLUSH_THRESHOLD": "50000", "CM_TEST_EXCHANGE_MODE": "1"},
    "flush-deferred": {"CM_TEST_FLUSH_THRESHOLD": "50000", "CM_TEST_APPLY_POLICY": "1"},
}


@pytest.mark.parametrize("variant", sorted(MP_VARIANTS))
@pytest.mark.parametrize("n", [2, 8])
def test_multi_gpu_parity_of_variants(n, variant):
    import os
    import subprocess
    import sys
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29750 + n * 10 + sorted(MP_VARIANTS).index(variant)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "tests", "mp_parity.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **MP_VARIANTS[variant]))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
