"""-m gpu parity tests: CUDA path (through the C ABI) vs the oracle on seeded inputs."""
import os

import numpy as np
import pytest

import cm_b200 as cm
from scenario import (Cfg, assert_blocks_match, assert_wire_match, rand_payload, run_checker, run_cuda,
                      slice_steps, sorted_unique)

from oracle_bindings import ref_available  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1200)]  # a hang must become a failure (pytest-timeout)


@pytest.mark.parametrize("dtype", [cm.F64, cm.F32, cm.I32])
def test_single_rank_slices(dtype):
    cfg = Cfg(dtype, 2048, 2048, 1, 1, 64, 64, 2048)
    rng = np.random.default_rng(11)
    steps = slice_steps(cfg, rng, 40, 200, 150, commit_every=16)
    assert_blocks_match(cfg, run_cuda(cfg, steps), run_checker(cfg, steps))


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("geom", [(1, 1, 2048, 2048, 2048), (2, 2, 2000, 1000, 1024), (1, 1, 20000, 300, 20000)])
def test_both_scatter_add_kernels(kernel, geom):
    """RED.ADD kernel (1) and column-tile kernel (2) forced, incl. a local block taller than one
    shared-memory band (20000 rows of fp64 > 8192)"""
    nprow, npcol, m, n, lld = geom
    cfg = Cfg(cm.F64, m, n, nprow, npcol, 64, 64, lld)
    rng = np.random.default_rng(14)
    steps = slice_steps(cfg, rng, 6, min(m, 700), min(n, 90), commit_every=3)
    got = run_cuda(cfg, steps, apply_kernel=kernel)
    assert all((st["tile_launches"] > 0) == (kernel == 2) for st in got["stats"])
    assert_blocks_match(cfg, got, run_checker(cfg, steps))


def test_single_rank_threshold_flush_and_unsorted_sweep():
    cfg = Cfg(cm.F64, 2048, 2048, 1, 1, 64, 64, 2048)
    rng = np.random.default_rng(12)
    steps = slice_steps(cfg, rng, 30, 128, 128)
    want = run_checker(cfg, steps)
    got = run_cuda(cfg, steps, threshold=50000)
    assert got["stats"][0]["flushes"] > 5
    assert_blocks_match(cfg, got, want)
    assert_blocks_match(cfg, run_cuda(cfg, steps, sorted_sweep=False, apply_kernel=1), want)


def test_single_rank_device_api():
    cfg = Cfg(cm.F64, 2048, 2048, 1, 1, 64, 64, 2048)
    rng = np.random.default_rng(13)
    steps = slice_steps(cfg, rng, 20, 256, 256)
    got = run_cuda(cfg, steps, device_api=True)
    assert got["stats"][0]["h2d_bytes"] < 64 * 1024  # payloads never crossed PCIe
    assert_blocks_match(cfg, got, run_checker(cfg, steps))


@pytest.mark.parametrize("dtype", [cm.F64, cm.F32, cm.I32])
def test_k0_grid_2x2_slices(dtype):
    """the reference's randomSliceUpdates geometry (test/cm_test.cpp:160-218), seeded, non-unit payloads"""
    cfg = Cfg(dtype, 2000, 1000, 2, 2, 64, 64, 1024)
    rng = np.random.default_rng(21)
    steps = slice_steps(cfg, rng, 3, 800, 500, commit_every=1)
    assert_blocks_match(cfg, run_cuda(cfg, steps), run_checker(cfg, steps))


@pytest.mark.parametrize("mode", [1, 2])
def test_both_exchange_modes(mode):
    """1 = all-to-all-v through the transport from a send arena, 2 = pack kernel stores straight into
    the owners' receive arenas (peer-to-peer), 3 = copy-engine pushes; same messages, same result"""
    cfg = Cfg(cm.F64, 2000, 1000, 2, 2, 64, 64, 1024)
    rng = np.random.default_rng(22)
    steps = slice_steps(cfg, rng, 4, 500, 300, commit_every=1)  # growing then steady arenas
    for r in range(cfg.P):
        steps.append(("elems", r, rand_payload(rng, cfg.dtype, 300), rng.integers(0, cfg.m, 300), rng.integers(0, cfg.n, 300)))
    steps.append(("commit",))
    want = run_checker(cfg, steps)
    got = run_cuda(cfg, steps, exchange_mode=mode)
    assert_blocks_match(cfg, got, want)
    assert all(st["bytes_sent"] > 0 for st in got["stats"])


def test_pipelined_commit_chunks():
    """a commit big enough to be cut into column panels and pipelined (scatter-add of panel k on the
    apply stream while panel k+1 is packed across the exchange): > 32 Mi elements staged per rank, and the owners'
    bar (the busiest one receives 128 Mi by default) lowered to the same value"""
    import torch
    mode = 2
    cfg = Cfg(cm.F32, 1024, 1024, 2, 2, 64, 64, 512)
    old_env = os.environ.get("CM_B200_PIPELINE_MIN_ELEMS")
    os.environ["CM_B200_PIPELINE_MIN_ELEMS"] = str(32 << 20)
    world = cm.LoopbackWorld(4, 0)
    try:
        mats = world.run(lambda r: cm.Matrix(*cfg.args()))
        world.run(lambda r: mats[r].set_exchange_mode(mode))
        g = torch.Generator(device="cuda")
        g.manual_seed(3)
        U, m_u, n_u = 140, 512, 512  # 36.7 Mi elements per rank
        keep, sums = [], []
        for r in range(4):
            ix = torch.rand((U, cfg.m), generator=g, device="cuda").topk(m_u, dim=1).indices.sort(dim=1).values.contiguous()
            jx = torch.rand((U, cfg.n), generator=g, device="cuda").topk(n_u, dim=1).indices.sort(dim=1).values.contiguous()
            data = torch.randint(-3, 4, (U, n_u, m_u), generator=g, device="cuda").to(torch.float32).contiguous()
            keep.append((ix, jx, data))
            dense = torch.zeros((cfg.n, cfg.m), dtype=torch.float32, device="cuda")
            for u in range(U):
                dense[jx[u][:, None], ix[u][None, :]] += data[u]
            sums.append(dense)
        want = sum(sums).cpu().numpy()  # want[j, i]; small integers: exact in fp32 in any order
        torch.cuda.synchronize()

        def issue(r):
            ix, jx, data = keep[r]
            for u in range(U):
                mats[r].update_device(data[u].data_ptr(), m_u, n_u, m_u, False, ix[u].data_ptr(), jx[u].data_ptr())
            mats[r].commit()
            return mats[r].local(), mats[r].stats()
        out = world.run(issue)
        for r, (blk, st) in enumerate(out):
            blk = blk.reshape(-1, cfg.lld)
            pr, pc = r // 2, r % 2
            gi = np.array([(l // 64) * 128 + pr * 64 + l % 64 for l in range(512)])
            gj = np.array([(l // 64) * 128 + pc * 64 + l % 64 for l in range(512)])
            assert np.array_equal(blk[:512, :512], want[np.ix_(gj, gi)])
        # 4 column panels => 4 scatter-add launches per rank for the second commit
        world.run(lambda r: (mats[r].reset_stats(), mats[r].set_profiling(True)))
        out2 = world.run(issue)
        assert all(st["apply_launches"] == 4 for _, st in out2), [st["apply_launches"] for _, st in out2]
        world.run(lambda r: mats[r].destroy())
    finally:
        world.close()
        if old_env is None:
            os.environ.pop("CM_B200_PIPELINE_MIN_ELEMS", None)
        else:
            os.environ["CM_B200_PIPELINE_MIN_ELEMS"] = old_env


@pytest.mark.parametrize("grid", [(2, 3, 4, 8, 64, 100, 190), (3, 2, 8, 4, 96, 250, 150), (2, 4, 64, 64, 1024, 2000, 3000),
                                  (1, 2, 64, 64, 1024, 1000, 2000), (2, 1, 64, 64, 1024, 2000, 1000)])
def test_nonsquare_grids_wire_bit_exact(grid):
    nprow, npcol, mb, nb, lld, m, n = grid
    dtype = cm.F32 if (nprow, npcol) == (3, 2) else cm.F64
    cfg = Cfg(dtype, m, n, nprow, npcol, mb, nb, lld)
    rng = np.random.default_rng(31)
    steps = slice_steps(cfg, rng, 3, min(m, 90), min(n, 70), commit_every=2)
    want = run_checker(cfg, steps, wire=True, threshold=1 << 40)  # one message per commit, like ours
    got = run_cuda(cfg, steps, wire=True)
    assert_wire_match(cfg, got, want)
    assert_blocks_match(cfg, got, want)


def test_single_rank_wire_bit_exact():
    cfg = Cfg(cm.F64, 1000, 1000, 1, 1, 64, 64, 1024)
    rng = np.random.default_rng(32)
    steps = slice_steps(cfg, rng, 5, 100, 80)
    want = run_checker(cfg, steps, wire=True, threshold=1 << 40)
    got = run_cuda(cfg, steps, wire=True)
    assert_wire_match(cfg, got, want)
    assert_blocks_match(cfg, got, want)


def test_edge_cases_duplicates_unsorted_trans_ld():
    cfg = Cfg(cm.F64, 500, 300, 2, 2, 64, 64, 1024)
    rng = np.random.default_rng(41)
    steps = []
    # duplicates + unsorted indices must accumulate (Appendix A item 4)
    ix = rng.integers(0, cfg.m, 70).astype(np.int64)
    jx = rng.integers(0, cfg.n, 50).astype(np.int64)
    ix[5] = ix[6] = ix[7]
    jx[1] = jx[2]
    steps.append(("slice", 0, rand_payload(rng, cfg.dtype, (50, 70)), 70, 50, 70, False, ix, jx))
    # padded leading dimension
    ix = sorted_unique(rng, cfg.m, 33)
    jx = sorted_unique(rng, cfg.n, 21)
    steps.append(("slice", 1, rand_payload(rng, cfg.dtype, (21, 40)), 33, 21, 40, False, ix, jx))
    # transposed view (LocalMatrix::trans, include/local_matrix.hpp:58-72): storage 21 x 33, ld 25
    steps.append(("slice", 2, rand_payload(rng, cfg.dtype, (33, 25)), 33, 21, 25, True, ix, jx))
    # 1 x 1, single row, single column
    one = np.array([3], dtype=np.int64)
    steps.append(("slice", 3, rand_payload(rng, cfg.dtype, (1, 1)), 1, 1, 1, False, one, one))
    steps.append(("slice", 3, rand_payload(rng, cfg.dtype, (21, 1)), 1, 21, 1, False, one, jx))
    steps.append(("slice", 0, rand_payload(rng, cfg.dtype, (1, 33)), 33, 1, 33, False, ix, one))
    steps.append(("commit",))
    steps.append(("commit",))  # empty commit
    want = run_checker(cfg, steps, wire=True, threshold=1 << 40)
    got = run_cuda(cfg, steps, wire=True)
    assert_wire_match(cfg, got, want)
    assert_blocks_match(cfg, got, want)
    # the same on one rank through the direct (no-pack) path
    cfg1 = Cfg(cm.F64, 500, 300, 1, 1, 64, 64, 512)
    steps1 = [s if s[0] != "slice" else (s[0], 0) + s[2:] for s in steps]
    assert_blocks_match(cfg1, run_cuda(cfg1, steps1), run_checker(cfg1, steps1))


@pytest.mark.parametrize("P", [1, 4])
def test_elemental_updates(P):
    """randomElementUpdates (test/cm_test.cpp:44-101), seeded"""
    cfg = Cfg(cm.F32, 2000, 1000, 2, 2, 64, 64, 1024) if P == 4 else Cfg(cm.F32, 2000, 1000, 1, 1, 64, 64, 2048)
    rng = np.random.default_rng(51)
    steps = []
    for it in range(3):
        for r in range(cfg.P):
            n = 5000
            steps.append(("elems", r, rand_payload(rng, cfg.dtype, n), rng.integers(0, cfg.m, n), rng.integers(0, cfg.n, n)))
        steps.append(("commit",))
    want = run_checker(cfg, steps)
    assert_blocks_match(cfg, run_cuda(cfg, steps), want)


@pytest.mark.parametrize("P", [1, 4])
def test_symmetric_update_and_fill_lower(P):
    """randomSliceUpdatesSymmetricFill (test/cm_test.cpp:220-282), seeded, non-unit payloads"""
    cfg = Cfg(cm.F64, 2000, 2000, 2, 2, 64, 64, 1024) if P == 4 else Cfg(cm.F64, 2000, 2000, 1, 1, 64, 64, 2048)
    rng = np.random.default_rng(61)
    steps = []
    for it in range(3):
        for r in range(cfg.P):
            ix = sorted_unique(rng, cfg.m, 300)
            steps.append(("sym", r, rand_payload(rng, cfg.dtype, (300, 300)), 300, 300, False, ix))
        steps.append(("commit",))
    steps += [("fill_lower",), ("commit",)]
    want = run_checker(cfg, steps)
    got = run_cuda(cfg, steps)
    assert_blocks_match(cfg, got, want)


def test_reset_and_multiple_rounds():
    cfg = Cfg(cm.F64, 600, 600, 2, 2, 64, 64, 512)
    rng = np.random.default_rng(71)
    steps = slice_steps(cfg, rng, 4, 100, 100, commit_every=2)
    steps.append(("reset",))
    steps += slice_steps(cfg, rng, 2, 64, 64)
    assert_blocks_match(cfg, run_cuda(cfg, steps), run_checker(cfg, steps))


def test_deterministic_mode_bit_exact_run_to_run():
    cfg = Cfg(cm.F64, 1024, 1024, 1, 1, 64, 64, 1024)
    rng = np.random.default_rng(81)
    steps = slice_steps(cfg, rng, 60, 256, 256)  # ~3.7 hits per element: order matters
    a = run_cuda(cfg, steps, deterministic=True)
    b = run_cuda(cfg, steps, deterministic=True)
    assert a["local"][0].tobytes() == b["local"][0].tobytes()
    assert_blocks_match(cfg, a, run_checker(cfg, steps))


# ---- against the committed fixtures generated from the reference itself -----------------
import os  # noqa: E402

from golden_scenarios import SCENARIOS, assert_wire_matches_golden, build_steps, load_golden  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_cuda_matches_reference_golden(name):
    spec = SCENARIOS[name]
    cfg, golden = load_golden(os.path.join(GOLDEN, name + ".npz"))
    got = run_cuda(cfg, build_steps(spec), wire=True, threshold=spec.get("ours_threshold"))
    if "ours_threshold" in spec:
        assert all(st["flushes"] > 0 for st in got["stats"])
    if cfg.nprow != cfg.npcol:
        golden = dict(golden, m_local=got["m_local"])  # reference ctor defect, SURVEY.md Appendix C-1
    assert_blocks_match(cfg, got, golden)
    if "wire_inds" in golden:
        assert_wire_matches_golden(cfg, got, golden, multiset=(spec["kind"] == "symmetric"))
    # and through the direct / non-captured path
    assert_blocks_match(cfg, run_cuda(cfg, build_steps(spec), threshold=spec.get("ours_threshold")), golden)


def test_accessors_descriptor_host_mirror_and_knobs():
    cfg = Cfg(cm.F64, 2000, 1000, 1, 1, 64, 64, 2048)
    cm.init(0, 1, 0)
    try:
        M = cm.Matrix(*cfg.args())
        assert (M.m, M.n, M.m_local, M.n_local, M.lld, M.pgrid_row, M.pgrid_col) == (2000, 1000, 2000, 1000, 2048, 0, 0)
        assert M.descinit(3) == [1, 3, 2000, 1000, 64, 64, 0, 0, 2048]
        assert M.flush_threshold() == 1 << 30 and M.progress_interval() == 1
        M.set_flush_threshold(12345)
        M.set_progress_interval(0)
        assert M.flush_threshold() == 12345 and M.progress_interval() == 0
        rng = np.random.default_rng(5)
        ix, jx = sorted_unique(rng, cfg.m, 100), sorted_unique(rng, cfg.n, 80)
        data = rand_payload(rng, cfg.dtype, (80, 100))
        M.update(data, ix, jx)
        M.commit()
        host = M.local_host_view()  # get_local_data(): host mirror of the HBM block
        assert np.array_equal(host, M.local())
        assert host[jx[3] * 2048 + ix[7]] == data[3, 7] == M.get(int(ix[7]), int(jx[3]))
        M.reset()
        assert not M.local().any()
        # contract violations fail loudly (the reference asserts, :417-428)
        with pytest.raises(cm.CmError):
            cm.Matrix(cm.F64, 2000, 1000, 1, 1, 64, 64, 1024)  # m_local > LLD
        with pytest.raises(cm.CmError):
            cm.Matrix(cm.F64, 2000, 1000, 2, 2, 64, 64, 1024)  # grid != rank_n()
        M.destroy()
    finally:
        cm.finalize()


def test_cpp_reference_suite_port():
    """cpp_tests/cm_test: the reference's own four scenarios (test/cm_test.cpp) through the C++
    drop-in headers, 4 ranks as threads on this GPU, T in {int, float, double}"""
    import subprocess
    exe = os.path.join(os.path.dirname(GOLDEN), "..", "cpp_tests", "cm_test")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe)])
    p = subprocess.run([exe, "3"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "PASS" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]


def test_full_size_properties_k1():
    """BASELINE K1 geometry at full size (N = 8192, 256 x 256 updates): properties that need no
    oracle run — linearity (A then -A cancels exactly... within rounding of one add) and the
    column/row sums of an all-ones update set"""
    import torch
    cfg = Cfg(cm.F64, 8192, 8192, 1, 1, 64, 64, 8192)
    cm.init(0, 1, 0)
    try:
        M = cm.Matrix(*cfg.args())
        g = torch.Generator(device="cuda")
        g.manual_seed(1)
        U = 512
        ones = torch.ones((256, 256), dtype=torch.float64, device="cuda")
        ix = torch.rand((U, 8192), generator=g, device="cuda").topk(256, dim=1).indices.sort(dim=1).values.contiguous()
        jx = torch.rand((U, 8192), generator=g, device="cuda").topk(256, dim=1).indices.sort(dim=1).values.contiguous()
        torch.cuda.synchronize()
        for u in range(U):
            M.update_device(ones.data_ptr(), 256, 256, 256, False, ix[u].data_ptr(), jx[u].data_ptr())
        M.commit()
        blk = torch.from_numpy(M.local()).reshape(8192, 8192)  # [col, row]
        assert blk.sum().item() == U * 256 * 256  # every element landed exactly once
        rows = torch.zeros(8192, dtype=torch.float64)
        rows.index_add_(0, ix.reshape(-1).cpu(), torch.full((U * 256,), 256.0, dtype=torch.float64))
        assert torch.equal(blk.sum(dim=0), rows)  # per-row hit counts
        neg = -ones
        for u in range(U):
            M.update_device(neg.data_ptr(), 256, 256, 256, False, ix[u].data_ptr(), jx[u].data_ptr())
        M.commit()
        assert not M.local().any()  # integers: exact cancellation
        M.destroy()
    finally:
        cm.finalize()


def test_empty_ragged_and_lopsided_inputs():
    """zero-sized updates are no-ops, ranks may stage nothing at all, every element of an update may
    belong to a single owner (all other owners receive empty segments), commits may be empty"""
    cfg = Cfg(cm.F64, 512, 512, 2, 2, 64, 64, 256)
    rng = np.random.default_rng(91)
    empty = np.zeros(0, dtype=np.int64)
    steps = [("commit",)]  # nothing staged anywhere
    # rank 0: rows and columns all inside blocks owned by grid position (1,1) = rank 3
    ix = np.concatenate([np.arange(64, 128), np.arange(192, 230)]).astype(np.int64)
    jx = np.concatenate([np.arange(64, 100), np.arange(448, 512)]).astype(np.int64)
    steps.append(("slice", 0, rand_payload(rng, cfg.dtype, (len(jx), len(ix))), len(ix), len(jx), len(ix), False, ix, jx))
    # rank 1: zero rows / zero columns
    steps.append(("slice", 1, np.zeros(0), 0, 5, 1, False, empty, np.arange(5, dtype=np.int64)))
    steps.append(("slice", 1, np.zeros(0), 7, 0, 7, False, np.arange(7, dtype=np.int64), empty))
    # rank 2 stages nothing; rank 3 a ragged pair of updates (1 x 300 and 300 x 1)
    one = np.array([511], dtype=np.int64)
    many = sorted_unique(rng, 512, 300)
    steps.append(("slice", 3, rand_payload(rng, cfg.dtype, (300, 1)), 1, 300, 1, False, one, many))
    steps.append(("slice", 3, rand_payload(rng, cfg.dtype, (1, 300)), 300, 1, 300, False, many, one))
    steps += [("commit",), ("commit",)]
    want = run_checker(cfg, steps, wire=True, threshold=1 << 40)
    got = run_cuda(cfg, steps, wire=True)
    assert_wire_match(cfg, got, want)
    assert_blocks_match(cfg, got, want)
    assert len(got["wire"][(0, 3)][0]) == len(ix) * len(jx) and len(got["wire"][(0, 0)][0]) == 0


def test_contract_violations_fail_loudly():
    """what the reference asserts on (or leaves undefined) is an error code here, never a silent fallback"""
    cm.init(0, 1, 0)
    try:
        M = cm.Matrix(cm.F64, 256, 256, 1, 1, 64, 64, 256)
        d = np.zeros((4, 4))
        ix = np.arange(4, dtype=np.int64)
        with pytest.raises(cm.CmError):
            M.update(d, ix, ix, m_u=-1, n_u=4)
        with pytest.raises(cm.CmError):
            M.update(d, ix, ix, m_u=4, n_u=4, ld=2)  # leading dimension smaller than the rows stored
        with pytest.raises(cm.CmError):
            M.update_device(0, 4, 4, 4, False, 0, 0)  # null device pointers
        with pytest.raises(cm.CmError):
            M.set_apply_kernel(7)
        with pytest.raises(cm.CmError):
            M.set_exchange_mode(9)
        M.update(d + 1.0, ix, ix)  # the matrix is still usable afterwards
        M.commit()
        assert M.local().sum() == 16.0
        M.destroy()
    finally:
        cm.finalize()


@pytest.mark.parametrize("P", [1, 4])
def test_save_load_reference_file_format(P, tmp_path):
    """randomElementUpdatesMPIIO (test/cm_test.cpp:286-349) without MPI: save() writes the dense global
    matrix column-major with no header (what the reference's MPI darray view writes), load() into a
    second instance reproduces every local block"""
    cfg = Cfg(cm.F64, 2000, 1000, 2, 2, 64, 64, 1024) if P == 4 else Cfg(cm.F64, 2000, 1000, 1, 1, 64, 64, 2048)
    rng = np.random.default_rng(95)
    steps = slice_steps(cfg, rng, 2, 300, 200)
    for r in range(cfg.P):
        steps.insert(-1, ("elems", r, rand_payload(rng, cfg.dtype, 2000), rng.integers(0, cfg.m, 2000), rng.integers(0, cfg.n, 2000)))
    want = run_checker(cfg, steps)
    dense = np.zeros((cfg.n, cfg.m))  # dense[j, i] = column-major global matrix
    for r in range(cfg.P):
        blk = want["local"][r].reshape(-1, cfg.lld)
        gi = np.array([(l // 64) * 64 * cfg.nprow + (r // cfg.npcol) * 64 + l % 64 for l in range(want["m_local"][r])])
        gj = np.array([(l // 64) * 64 * cfg.npcol + (r % cfg.npcol) * 64 + l % 64 for l in range(want["n_local"][r])])
        dense[np.ix_(gj, gi)] = blk[:len(gj), :len(gi)]
    path = str(tmp_path / "dist_mat1")
    world = None
    if P == 1:
        cm.init(0, 1, 0)
        run, on = (lambda fn: [fn(0)]), (lambda r, fn: fn(r))
    else:
        world = cm.LoopbackWorld(P, 0)
        run, on = world.run, world.on
    try:
        mats = run(lambda r: cm.Matrix(*cfg.args()))
        for st in steps:
            if st[0] == "slice":
                on(st[1], lambda rr: mats[rr].update(st[2], st[7], st[8], st[3], st[4], st[5], st[6]))
            elif st[0] == "elems":
                on(st[1], lambda rr: mats[rr].update_elems(st[2], st[3], st[4]))
            else:
                run(lambda r: mats[r].commit())
        run(lambda r: mats[r].save(path))
        on_disk = np.fromfile(path, dtype=np.float64)
        assert on_disk.size == cfg.m * cfg.n
        err = np.linalg.norm(on_disk - dense.reshape(-1)) / np.linalg.norm(dense)
        assert err <= 1e-12
        mats2 = run(lambda r: cm.Matrix(*cfg.args()))
        run(lambda r: mats2[r].load(path))
        a, b = run(lambda r: mats[r].local()), run(lambda r: mats2[r].local())
        for r in range(cfg.P):
            assert np.array_equal(a[r], b[r])  # bit-exact round trip, padding rows stay zero
        run(lambda r: (mats[r].destroy(), mats2[r].destroy()))
    finally:
        if world is not None:
            world.close()
        else:
            cm.finalize()


@pytest.mark.parametrize("mode", [2, 1])
def test_multi_rank_threshold_flush(mode):
    """bin_flush_threshold with several ranks (reference :346, :614, flush(thresh) inside update()): a rank whose
    staged elements exceed the threshold maps and packs them at once, NOT collectively — into the owners' flush
    regions over the peer-to-peer path once those exist (mode 2, from the second round on; the fifth flush of a
    round no longer fits and waits at the source), into buffers of the source otherwise (round 1; mode 1:
    all-to-all-v) — and the owners apply them inside the next commit. Per-(source, owner) streams stay
    bit-exact with the reference's, blocks within tolerance, staging is released at every flush."""
    cfg = Cfg(cm.F64, 2000, 1000, 2, 2, 64, 64, 1024)
    rng = np.random.default_rng(77)
    steps, steps_e = [], []  # steps_e: the same with elemental updates in between (they stay in host bins until the commit,
    for rnd in range(3):     # so their place in the stream differs from the reference's: blocks only)
        for u in range(12):
            for r in range(cfg.P):
                if r == 3 and rnd == 1:
                    continue  # a rank that flushes nothing in a round
                ix, jx = sorted_unique(rng, cfg.m, 200), sorted_unique(rng, cfg.n, 150)
                steps.append(("slice", r, rand_payload(rng, cfg.dtype, (150, 200)), 200, 150, 200, False, ix, jx))
                steps_e.append(steps[-1])
            if u == 5:
                v = rand_payload(rng, cfg.dtype, (40,))
                steps_e.append(("elems", 1, v, rng.integers(0, cfg.m, 40), rng.integers(0, cfg.n, 40)))
        steps.append(("commit",))
        steps_e.append(("commit",))
    kind = "ref" if ref_available() else "oracle"
    want = run_checker(cfg, steps, kind=kind, wire=True)
    got = run_cuda(cfg, steps, wire=True, threshold=50000, exchange_mode=mode)  # 30000 elements per update: a flush every 2
    assert_wire_match(cfg, got, want)
    assert_blocks_match(cfg, got, want)
    assert_blocks_match(cfg, run_cuda(cfg, steps_e, threshold=50000, exchange_mode=mode), run_checker(cfg, steps_e, kind=kind))
    for r, st in enumerate(got["stats"]):
        assert st["flushes"] >= 12, st
        if mode == 2:
            assert st["flushed_shipped"] > 0 and st["flushed_held"] > 0, st
        else:
            assert st["flushed_shipped"] == 0 and st["flushed_held"] > 0, st
    # same through the device API (borrowed payloads are released by the flush) and in deterministic mode
    got2 = run_cuda(cfg, steps, threshold=50000, exchange_mode=mode, device_api=True, deterministic=True, free_after_flush=True)
    assert_blocks_match(cfg, got2, want)
    got3 = run_cuda(cfg, steps, threshold=50000, exchange_mode=mode, device_api=True, deterministic=True, free_after_flush=True)
    assert all(a.tobytes() == b.tobytes() for a, b in zip(got2["local"], got3["local"]))


def test_flush_call_and_staging_limit_several_ranks():
    """cm_flush() on several ranks packs what is staged (it used to be a no-op there), and a rank that reaches the
    limit of updates one scatter-add can address flushes by itself instead of refusing the update; the owners
    then apply more segments than one scatter-add addresses in several launches. Deferred policy: flushed
    messages join the log."""
    cfg = Cfg(cm.F32, 1000, 600, 2, 2, 64, 64, 512)
    rng = np.random.default_rng(78)
    steps = slice_steps(cfg, rng, 23, 48, 40, commit_every=12)
    want = run_checker(cfg, steps)
    for policy in (0, 1):
        got = run_cuda(cfg, steps, staging_limit=5, apply_policy=policy)
        assert all(st["flushes"] >= 4 for st in got["stats"])
        assert_blocks_match(cfg, got, want)
    got = run_cuda(cfg, steps, flush_every=3)
    assert all(st["flushes"] >= 6 for st in got["stats"])
    assert_blocks_match(cfg, got, want)


def test_scatter_add_choice_follows_row_contiguity():
    """The owner picks the scatter-add per flush from what the sources report: hits per element of the block AND the
    share of elements whose row directly follows the previous one. Runs of consecutive indices (K3's clustered sets)
    make RED cheap (whole sectors), so the same volume that goes to the tile kernel with scattered rows stays on RED."""
    cfg = Cfg(cm.F64, 2048, 2048, 2, 2, 64, 64, 1024)
    rng = np.random.default_rng(79)

    def runs(hi, k):  # k / 64 runs of 64 consecutive indices at block boundaries
        blocks = np.sort(rng.choice(hi // 64, size=k // 64, replace=False))
        return (blocks[:, None] * 64 + np.arange(64)[None, :]).reshape(-1).astype(np.int64)

    for clustered in (True, False):
        steps = []
        for u in range(16):
            for r in range(cfg.P):
                ix = runs(cfg.m, 256) if clustered else sorted_unique(rng, cfg.m, 256)
                jx = runs(cfg.n, 256) if clustered else sorted_unique(rng, cfg.n, 256)
                steps.append(("slice", r, rand_payload(rng, cfg.dtype, (256, 256)), 256, 256, 256, False, ix, jx))
        steps.append(("commit",))
        got = run_cuda(cfg, steps)  # ~1 hit per element of every block
        assert_blocks_match(cfg, got, run_checker(cfg, steps))
        for st in got["stats"]:
            assert (st["tile_launches"] == 0) == clustered, (clustered, st)
