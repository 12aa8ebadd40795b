"""-m gpu: parity at the GEOMETRY of the configurations BASELINE.json names, against the
reference itself (oracle/_ref: the unmodified reference header over the in-repo UPC++ shim; the C restatement
where that library is missing), with U(-1,1) payloads, local blocks within the north_star tolerance and the
per-owner wire streams bit for bit.

  K1  <double,1,1,64,64,8192>, 8192 x 8192, 256 x 256 updates: full size, one flush of 2048 updates
  K2  <double,2,4,64,64,16384>: LLD 16384 = two row bands per column tile, 512 x 512 updates, loopback 2 x 4
  K3  <double,2,4,64,64,32768>: four row bands, clustered index sets (runs of 64 consecutive indices at block
      boundaries, blocks of grid row / column 0 drawn with weight 4), loopback 2 x 4
  K4  <float,4,2,64,64,8192>: 32 x 32 updates, commit every 1024 updates per rank, loopback 4 x 2
The multi-rank cases keep the configuration's template parameters and global row count and narrow the global
column count, so that eight local blocks fit one GPU and one host: local indices do not depend on N.
"""
import numpy as np
import pytest

import cm_b200 as cm
from oracle_bindings import RefShim, ref_available
from scenario import Cfg, assert_blocks_match, assert_wire_match, rand_payload, run_checker, run_cuda, slice_steps, sorted_unique

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _kind(cfg):
    ok = ref_available() and RefShim.have_config(cfg.dtype, cfg.nprow, cfg.npcol, cfg.mb, cfg.nb, cfg.lld)
    return "ref" if ok else "oracle"


def test_k1_full_size_one_flush_of_2048_updates():
    cfg = Cfg(cm.F64, 8192, 8192, 1, 1, 64, 64, 8192)
    rng = np.random.default_rng(101)
    steps = slice_steps(cfg, rng, 2048, 256, 256)  # 134 Mi elements, ~2 hits per element of the block: one tile sweep
    got = run_cuda(cfg, steps, device_api=True)
    assert got["stats"][0]["flushes"] == 1 and got["stats"][0]["kernel_launches"] > 0
    want = run_checker(cfg, steps, kind=_kind(cfg))
    assert_blocks_match(cfg, got, want)
    # wire streams (the pack path on one rank) on the first 256 updates
    sub = steps[:256] + [("commit",)]
    want = run_checker(cfg, sub, kind=_kind(cfg), wire=True, threshold=2 ** 31 - 1)
    got = run_cuda(cfg, sub, wire=True)
    assert_wire_match(cfg, got, want)
    assert_blocks_match(cfg, got, want)


@pytest.mark.parametrize("apply_kernel", [0, 1, 2])
def test_k2_geometry_two_row_bands(apply_kernel):
    cfg = Cfg(cm.F64, 32768, 4096, 2, 4, 64, 64, 16384)
    rng = np.random.default_rng(102)
    steps = slice_steps(cfg, rng, 40, 512, 512, commit_every=20)  # 0.31 hits per element and commit: tile path (auto)
    want = run_checker(cfg, steps, kind=_kind(cfg), wire=True, threshold=2 ** 31 - 1)
    got = run_cuda(cfg, steps, wire=True, apply_kernel=apply_kernel)
    assert_wire_match(cfg, got, want)
    assert_blocks_match(cfg, got, want)
    if apply_kernel != 1:
        assert all(st["tile_launches"] > 0 for st in got["stats"])


def _clustered(rng, hi, k, blk, np_):
    nb = hi // blk
    w = np.ones(nb)
    w[::np_] = 4.0
    blocks = np.sort(rng.choice(nb, size=k // blk, replace=False, p=w / w.sum()))
    return (blocks[:, None] * blk + np.arange(blk)[None, :]).reshape(-1).astype(np.int64)


@pytest.mark.parametrize("pipelined", [False, True])
def test_k3_geometry_clustered_index_sets(pipelined, monkeypatch):
    cfg = Cfg(cm.F64, 65536, 4096, 2, 4, 64, 64, 32768)
    rng = np.random.default_rng(103)
    steps = []
    for u in range(10):
        for r in range(cfg.P):
            ix = _clustered(rng, cfg.m, 512, cfg.mb, cfg.nprow)
            jx = _clustered(rng, cfg.n, 512, cfg.nb, cfg.npcol)
            steps.append(("slice", r, rand_payload(rng, cfg.dtype, (512, 512)), 512, 512, 512, False, ix, jx))
        if u == 4:
            steps.append(("commit",))
    steps.append(("commit",))
    if pipelined:
        monkeypatch.setenv("CM_B200_PIPELINE_MIN_ELEMS", "1")
    else:
        monkeypatch.setenv("CM_B200_PIPELINE", "0")
    want = run_checker(cfg, steps, kind=_kind(cfg), wire=True, threshold=2 ** 31 - 1)
    got = run_cuda(cfg, steps, wire=True)
    assert_wire_match(cfg, got, want)
    assert_blocks_match(cfg, got, want)
    # the hot owner (0,0) really is hot: well above the balanced share
    applied = [st["elements_applied"] for st in got["stats"]]
    assert applied[0] > 2.5 * sum(applied) / cfg.P


def test_k4_geometry_tiny_updates_many_commits():
    cfg = Cfg(cm.F32, 32768, 4096, 4, 2, 64, 64, 8192)
    rng = np.random.default_rng(104)
    # 1 Mi elements per rank and commit, as in K4, on a block of 8192 x 2048: 0.06 hits per element (K4 itself: 0.008)
    steps = slice_steps(cfg, rng, 2048, 32, 32, commit_every=1024)
    want = run_checker(cfg, steps, kind=_kind(cfg), wire=True, threshold=2 ** 31 - 1)
    got = run_cuda(cfg, steps, wire=True)
    assert_wire_match(cfg, got, want)
    assert_blocks_match(cfg, got, want)
    assert all(st["tile_launches"] == 0 for st in got["stats"])  # sparse flushes: RED path


@pytest.mark.parametrize("env", [{"CM_B200_PIPELINE": "0"}, {"CM_B200_PIPELINE_MIN_ELEMS": "1"}])
def test_deterministic_mode_run_to_run_on_2x4(env, monkeypatch):
    """deterministic-order mode on eight ranks: two runs give bit-identical local blocks (batch and pipelined
    commits), and they match the reference within tolerance"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    cfg = Cfg(cm.F64, 4096, 4096, 2, 4, 64, 64, 2048)
    rng = np.random.default_rng(105)
    steps = slice_steps(cfg, rng, 20, 512, 512, commit_every=10)  # ~2.5 hits per element: the order matters
    for r in range(cfg.P):
        steps.insert(-1, ("elems", r, rand_payload(rng, cfg.dtype, 500), rng.integers(0, cfg.m, 500), rng.integers(0, cfg.n, 500)))
    a = run_cuda(cfg, steps, deterministic=True)
    b = run_cuda(cfg, steps, deterministic=True)
    for r in range(cfg.P):
        assert a["local"][r].tobytes() == b["local"][r].tobytes(), f"rank {r}: not bit-exact run to run"
    assert_blocks_match(cfg, a, run_checker(cfg, steps, kind=_kind(cfg)))


def test_more_than_2_26_elements_per_rank_between_commits_on_2x2():
    """bin_flush_threshold with several ranks (reference :330-375, :614): every rank of a
    2 x 2 grid issues 71 Mi elements (> 2^26) through the HOST API between two commits, twice. A rank whose staged
    elements exceed the threshold (2^22) packs them at once and releases its staging: the arenas stay at one
    threshold's worth instead of growing with what is issued. Round 1 holds the flushed messages at the source (no
    flush regions yet), round 2 stores them into the owners' regions at flush time. Blocks match the reference."""
    cfg = Cfg(cm.F64, 16384, 4096, 2, 2, 64, 64, 8192)
    rng = np.random.default_rng(106)
    payloads = [rand_payload(rng, cfg.dtype, (512, 512)) for _ in range(8)]
    U, thr = 272, 1 << 22
    steps = []
    for rnd in range(2):
        for u in range(U):
            for r in range(cfg.P):
                steps.append(("slice", r, payloads[(u + r) % 8], 512, 512, 512, False, sorted_unique(rng, cfg.m, 512),
                              sorted_unique(rng, cfg.n, 512)))
        steps.append(("commit",))
    assert U * 512 * 512 > 1 << 26
    got = run_cuda(cfg, steps, threshold=thr)
    for st in got["stats"]:
        assert st["elements_staged"] == 2 * U * 512 * 512
        assert st["flushes"] >= 2 * (U * 512 * 512 // (thr + 512 * 512)) - 2
        assert st["flushed_held"] > 0 and st["flushed_shipped"] > 0
        assert st["staging_peak_bytes"] < 2 * (thr + 512 * 512) * 8, st  # bounded by the threshold, not by 2 x 570 MB issued
    assert_blocks_match(cfg, got, run_checker(cfg, steps, kind=_kind(cfg)))
