"""-m gpu: the opt-in variants built after round 1's GPU time ran out, on the
loopback transport (rank threads on one GPU). Kept in a file that sorts after the established GPU
suites, so that `pytest -x` reaches those first. The same scenarios run on the kernel-logic library in
the CPU suite (tests/test_kernel_logic_emu.py)."""
import numpy as np
import pytest

import cm_b200 as cm
from scenario import Cfg, assert_blocks_match, run_checker, run_cuda, slice_steps

# a disagreement between rank threads would be a hang, not a failure: bound every test (pytest-timeout)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.mark.parametrize("grid,mode,npanel", [((2, 2, 64, 64, 1024, 2000, 1000), 2, 4), ((2, 4, 16, 16, 256, 500, 900), 2, 4),
                                              ((3, 2, 8, 4, 96, 250, 150), 3, 3)])
def test_column_panels_wire_bit_exact(grid, mode, npanel):
    """commits cut into column panels (thresholds lowered through the environment so
    that small commits are cut too): local blocks match the oracle and, index sets being sorted, the
    per-owner wire streams are the reference's element for element"""
    from emu_cases import panels_wire_and_blocks
    panels_wire_and_blocks(grid, mode, npanel)


def test_flag_barrier_is_refused_between_rank_threads():
    """CM_B200_FLAG_BARRIER=1 applies to one-process-per-GPU runs only: rank threads sharing one GPU keep
    the transport's barrier (mutually waiting streams of one device could deadlock), results unchanged"""
    from emu_cases import _Env
    cfg = Cfg(cm.F64, 600, 600, 2, 2, 64, 64, 512)
    rng = np.random.default_rng(72)
    steps = slice_steps(cfg, rng, 3, 100, 100, commit_every=2)
    with _Env(CM_B200_FLAG_BARRIER=1):
        got = run_cuda(cfg, steps, exchange_mode=2)
    assert_blocks_match(cfg, got, run_checker(cfg, steps))


@pytest.mark.parametrize("mode,budget,policy", [(2, 0, 1), (3, 0, 1), (2, 40000, 1), (2, 0, 2), (2, 40000, 2)])
def test_deferred_scatter_add(mode, budget, policy):
    """cm_set_apply_policy(CM_APPLY_DEFERRED / CM_APPLY_BACKGROUND): commits deliver, the log is applied on
    read / over budget / when an arena grows (policy 2: also in the background, on the apply stream,
    while later commits pack); same local blocks as the eager oracle"""
    from emu_cases import deferred_apply
    deferred_apply(mode, budget, policy)


def test_deferred_scatter_add_api():
    from emu_cases import deferred_apply_api
    deferred_apply_api()


def test_device_layout_speculative_pack():
    """CM_B200_DEVICE_LAYOUT=1: layout on the device, pack launched before the host has read the counts"""
    from emu_cases import device_layout_speculative_pack
    device_layout_speculative_pack()


def test_segment_wise_scatter_add():
    """CM_B200_SEG_APPLY=1 on tiny updates (K4-shaped: 32 x 32 on a 4 x 2 grid) and on one rank"""
    from emu_cases import _Env
    for cfg in (Cfg(cm.F32, 2048, 2048, 4, 2, 64, 64, 512), Cfg(cm.F64, 1024, 1024, 1, 1, 64, 64, 1024)):
        rng = np.random.default_rng(181)
        steps = slice_steps(cfg, rng, 6, 32, 32, commit_every=3)
        with _Env(CM_B200_SEG_APPLY=1):
            got = run_cuda(cfg, steps)
        assert_blocks_match(cfg, got, run_checker(cfg, steps))


# ---- the same variants over real NVLink / NCCL, one process per GPU (skipped on single-GPU boxes) ----
MP_VARIANTS = {
    "panels": {"CM_B200_PIPELINE": "1", "CM_B200_PIPELINE_MIN_ELEMS": "1"},
    "panels+flags": {"CM_B200_PIPELINE": "1", "CM_B200_PIPELINE_MIN_ELEMS": "1", "CM_B200_FLAG_BARRIER": "1"},
    "flags+devlayout": {"CM_B200_FLAG_BARRIER": "1", "CM_B200_DEVICE_LAYOUT": "1"},
    "deferred": {"CM_TEST_APPLY_POLICY": "1"},
    "background+flags": {"CM_TEST_APPLY_POLICY": "2", "CM_B200_BG_MIN_ELEMS": "1", "CM_B200_FLAG_BARRIER": "1"},
    "dma+panels": {"CM_TEST_EXCHANGE_MODE": "3", "CM_B200_PIPELINE_MIN_ELEMS": "1"},
    "segapply+threads0": {"CM_B200_SEG_APPLY": "1", "CM_B200_PACK_THREADS": "0", "CM_B200_PINNED_DESCS": "1"},
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


# The two inline-PTX variants below have not had their first GPU run yet: they only run when asked for
# (CM_B200_TEST_PTX_VARIANTS=1, scripts/measure_variants.sh does), in child processes, so that an untested
# instruction sequence cannot take the established suite down with it.
PTX_VARIANTS = pytest.mark.skipif(not __import__("os").environ.get("CM_B200_TEST_PTX_VARIANTS"),
                                  reason="unmeasured inline-PTX variant: set CM_B200_TEST_PTX_VARIANTS=1")


@PTX_VARIANTS
def test_tile_write_back_by_bulk_reduce_add():
    """CM_B200_TILE_BULKRED=1: a tile is added to its column by ONE bulk reduce-add of the TMA unit
    instead of 16-byte load/add/store steps (read once per process: child process)"""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, 'tests')\n"
        "import cm_b200 as cm\n"
        "from scenario import Cfg, assert_blocks_match, run_checker, run_cuda, slice_steps\n"
        "for dtype, m in [(cm.F64, 2048), (cm.F32, 1000), (cm.I32, 1024), (cm.F64, 20000)]:\n"
        "    cfg = Cfg(dtype, m, 512, 1, 1, 64, 64, m)\n"
        "    steps = slice_steps(cfg, np.random.default_rng(4), 12, min(m, 600) // 4 * 4, 100, commit_every=6)\n"
        "    got = run_cuda(cfg, steps, apply_kernel=2)\n"
        "    assert got['stats'][0]['tile_launches'] > 0\n"
        "    assert_blocks_match(cfg, got, run_checker(cfg, steps))\n"
        "print('VARIANT_OK')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, CM_B200_TILE_BULKRED="1"), capture_output=True,
                       text=True, timeout=600)
    assert p.returncode == 0 and "VARIANT_OK" in p.stdout, p.stdout[-1500:] + p.stderr[-3000:]


@PTX_VARIANTS
def test_pack_with_bulk_stores_child_process():
    """test_pack_with_bulk_stores in a child process with a timeout"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys\nsys.path.insert(0, 'tests')\nimport test_gpu_variants as V\n"
            "V.test_pack_with_bulk_stores()\nprint('VARIANT_OK')\n")
    p = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "VARIANT_OK" in p.stdout, p.stdout[-1500:] + p.stderr[-3000:]


def test_pack_with_bulk_stores():
    """CM_B200_PACK_BULK=1: a column's run of rows for one owner row leaves the SM as one bulk copy;
    wire streams bit-exact, local blocks within tolerance (runs of >= 64 rows take the bulk path,
    shorter ones and odd byte counts the per-thread stores). On a GPU only through the child-process
    wrapper above; the kernel-logic suite calls it directly."""
    import os
    if not os.environ.get("CM_EMU_LIB") and not cm.LIB_PATH.endswith("_emu.so") and not os.environ.get("CM_B200_TEST_PTX_VARIANTS"):
        pytest.skip("unmeasured inline-PTX variant: set CM_B200_TEST_PTX_VARIANTS=1")
    from emu_cases import _Env
    from scenario import assert_wire_match
    for cfg, m_u, n_u in ((Cfg(cm.F64, 2000, 1000, 2, 2, 64, 64, 1024), 500, 300), (Cfg(cm.F32, 1500, 900, 2, 3, 16, 32, 768), 391, 200),
                          (Cfg(cm.F64, 600, 600, 1, 2, 64, 64, 640), 90, 70)):
        rng = np.random.default_rng(201)
        steps = slice_steps(cfg, rng, 4, m_u, n_u, commit_every=2)
        want = run_checker(cfg, steps, wire=True, threshold=1 << 40)
        with _Env(CM_B200_PACK_BULK=1):
            got = run_cuda(cfg, steps, wire=True, exchange_mode=2)
        assert_wire_match(cfg, got, want)
        assert_blocks_match(cfg, got, want)
