"""Extra cases for tests/emu/run_cases.py: the GPU tests that stage device-resident inputs with
torch (which needs a real GPU) restated with numpy — under the CUDA-runtime test double a "device"
pointer is a host pointer. They run ONLY against tests/emu/libcm_b200_emu.so."""
import os

import numpy as np

import cm_b200 as cm
from scenario import Cfg, assert_blocks_match, rand_payload, run_checker, slice_steps, sorted_unique


def _ptr(a):
    return a.ctypes.data


def device_api_single_rank():
    """cm_update_slice_device on one rank (borrowed payloads, no staging copies), threshold flushes"""
    cfg = Cfg(cm.F64, 2048, 2048, 1, 1, 64, 64, 2048)
    rng = np.random.default_rng(13)
    steps = slice_steps(cfg, rng, 20, 256, 256)
    cm.init(0, 1, 0)
    try:
        M = cm.Matrix(*cfg.args())
        M.set_flush_threshold(300000)
        keep = []
        for st in steps:
            if st[0] == "slice":
                _, _, data, m_u, n_u, ld, trans, ix, jx = st
                d, i, j = np.ascontiguousarray(data), np.ascontiguousarray(ix), np.ascontiguousarray(jx)
                keep.append((d, i, j))
                M.update_device(_ptr(d), m_u, n_u, ld, trans, _ptr(i), _ptr(j))
            else:
                M.commit()
        got = {"local": [M.local()], "m_local": [M.m_local], "n_local": [M.n_local], "stats": [M.stats()]}
        M.destroy()
    finally:
        cm.finalize()
    assert got["stats"][0]["flushes"] > 2
    assert_blocks_match(cfg, got, run_checker(cfg, steps))


def _pipelined(mode, env):
    """a commit large enough to be cut into chunks (scatter-add of chunk k while chunk k+1 is packed):
    the thresholds are lowered through the environment so that it stays small enough for fibers"""
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    cfg = Cfg(cm.F32, 256, 256, 2, 2, 16, 16, 128)  # 128 local columns = 8 blocks of NB: four 2-block panels
    world = cm.LoopbackWorld(4, 0)
    try:
        mats = world.run(lambda r: cm.Matrix(*cfg.args()))
        world.run(lambda r: mats[r].set_exchange_mode(mode))
        rng = np.random.default_rng(3)
        U, m_u, n_u = 24, 128, 128
        keep, want = [], np.zeros((cfg.n, cfg.m), dtype=np.float32)
        for r in range(4):
            ups = []
            for u in range(U):
                ix, jx = sorted_unique(rng, cfg.m, m_u), sorted_unique(rng, cfg.n, n_u)
                data = rng.integers(-3, 4, size=(n_u, m_u)).astype(np.float32)
                want[np.ix_(jx, ix)] += data
                ups.append((data, ix, jx))
            keep.append(ups)

        def issue(r):
            for data, ix, jx in keep[r]:
                mats[r].update_device(_ptr(data), m_u, n_u, m_u, False, _ptr(ix), _ptr(jx))
            mats[r].commit()
            return mats[r].local(), mats[r].stats()
        world.run(lambda r: (mats[r].reset_stats(), mats[r].set_profiling(True)))
        out = world.run(issue)
        for r, (blk, st) in enumerate(out):
            blk = blk.reshape(-1, cfg.lld)
            pr, pc = r // 2, r % 2
            gi = np.array([(l // 16) * 32 + pr * 16 + l % 16 for l in range(128)])
            gj = np.array([(l // 16) * 32 + pc * 16 + l % 16 for l in range(128)])
            assert np.array_equal(blk[:128, :128], want[np.ix_(gj, gi)]), f"rank {r}"
            assert st["apply_launches"] == 4, st["apply_launches"]  # one scatter-add per chunk
        world.run(lambda r: mats[r].destroy())
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        world.close()


def pipelined_commit_chunks(mode):
    _pipelined(mode, {"CM_B200_PIPELINE": "1", "CM_B200_PIPELINE_MIN_ELEMS": "1000"})


CASES = [
    ("device_api_single_rank", [()]),
    ("pipelined_commit_chunks", [(2,)]),
]


# ---- column panels (PanelPlan, csrc/cm_common.h) ---------------------------------------------
class _Env:
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def panels_wire_and_blocks(grid, mode, npanel):
    """a commit cut into column panels: same local blocks as the oracle, and — index sets being
    sorted — the very same per-owner wire stream as the reference (update-major over the panels)"""
    from scenario import assert_wire_match, run_cuda
    nprow, npcol, mb, nb, lld, m, n = grid
    cfg = Cfg(cm.F64, m, n, nprow, npcol, mb, nb, lld)
    rng = np.random.default_rng(131)
    steps = slice_steps(cfg, rng, 5, min(m, 120), min(n, 90), commit_every=3)
    want = run_checker(cfg, steps, wire=True, threshold=1 << 40)
    with _Env(CM_B200_PIPELINE=1, CM_B200_PIPELINE_MIN_ELEMS=1, CM_B200_PANELS=npanel):
        got = run_cuda(cfg, steps, wire=True, exchange_mode=mode)
    assert_wire_match(cfg, got, want)
    assert_blocks_match(cfg, got, want)
    # panels really were in use: more than one scatter-add launch per commit on some rank
    with _Env(CM_B200_PIPELINE=1, CM_B200_PIPELINE_MIN_ELEMS=1, CM_B200_PANELS=npanel):
        got2 = run_cuda(cfg, steps, exchange_mode=mode)
    assert_blocks_match(cfg, got2, want)


def panels_unsorted_masked_and_mixed_sources():
    """panels with unsorted / duplicate index sets, the symmetric update and fill_lower, elemental
    updates in the same commit, and sources that split next to sources that do not"""
    from scenario import run_cuda
    cfg = Cfg(cm.F64, 700, 700, 2, 2, 32, 32, 512)
    rng = np.random.default_rng(132)
    steps = []
    for r in range(cfg.P):
        ix = rng.integers(0, cfg.m, 150).astype(np.int64)  # unsorted, with duplicates
        jx = rng.integers(0, cfg.n, 110).astype(np.int64)
        steps.append(("slice", r, rand_payload(rng, cfg.dtype, (110, 150)), 150, 110, 150, False, ix, jx))
        sx = sorted_unique(rng, cfg.m, 90)
        steps.append(("sym", r, rand_payload(rng, cfg.dtype, (90, 90)), 90, 90, False, sx))
        steps.append(("elems", r, rand_payload(rng, cfg.dtype, 50), rng.integers(0, cfg.m, 50), rng.integers(0, cfg.n, 50)))
    # rank 3 stages one tiny extra update only in the second round: it will not split, the others do
    steps.append(("commit",))
    for r in range(3):
        ix, jx = sorted_unique(rng, cfg.m, 200), sorted_unique(rng, cfg.n, 200)
        steps.append(("slice", r, rand_payload(rng, cfg.dtype, (200, 200)), 200, 200, 200, False, ix, jx))
    steps.append(("slice", 3, rand_payload(rng, cfg.dtype, (3, 5)), 5, 3, 5, False, sorted_unique(rng, cfg.m, 5),
                  sorted_unique(rng, cfg.n, 3)))
    steps += [("commit",), ("fill_lower",), ("commit",)]
    want = run_checker(cfg, steps)
    for mode in (2,):
        with _Env(CM_B200_PIPELINE=1, CM_B200_PIPELINE_MIN_ELEMS=1000, CM_B200_PANELS=4):
            got = run_cuda(cfg, steps, exchange_mode=mode)
        assert_blocks_match(cfg, got, want)
        with _Env(CM_B200_PIPELINE=1, CM_B200_PIPELINE_MIN_ELEMS=1000, CM_B200_PANELS=3):
            got = run_cuda(cfg, steps, exchange_mode=mode, deterministic=True)
        assert_blocks_match(cfg, got, want)


CASES += [
    ("panels_wire_and_blocks", [((2, 2, 64, 64, 1024, 2000, 1000), 2, 4), ((2, 4, 16, 16, 256, 500, 900), 2, 4),
                                ((2, 4, 16, 16, 256, 500, 900), 2, 8), ((2, 2, 64, 64, 1024, 2000, 1000), 2, 7),
                                ((3, 2, 8, 4, 96, 250, 150), 2, 3), ((1, 2, 64, 64, 1024, 1000, 130), 2, 4),
                                ((2, 1, 64, 64, 1024, 2000, 40), 2, 2)]),
    ("panels_unsorted_masked_and_mixed_sources", [()]),
]


# ---- deferred scatter-add (cm_set_apply_policy) ---------------------------------
def deferred_apply(mode, budget, policy=1):
    """commit() only delivers; the log is applied when the block is read, when it outgrows its budget
    (budget small: every other commit) or when an arena has to grow; several rounds, elemental updates,
    the symmetric update, fill_lower and reset in between — same local blocks as the eager oracle"""
    from scenario import run_cuda
    cfg = Cfg(cm.F64, 900, 900, 2, 2, 32, 32, 512)
    rng = np.random.default_rng(151)
    steps = []
    for rnd in range(5):
        for r in range(cfg.P):
            k = 60 + 40 * rnd  # growing messages: the arenas have to grow while a log is pending
            ix, jx = sorted_unique(rng, cfg.m, k), sorted_unique(rng, cfg.n, k - 10)
            steps.append(("slice", r, rand_payload(rng, cfg.dtype, (k - 10, k)), k, k - 10, k, False, ix, jx))
            steps.append(("elems", r, rand_payload(rng, cfg.dtype, 30), rng.integers(0, cfg.m, 30), rng.integers(0, cfg.n, 30)))
            sx = sorted_unique(rng, cfg.m, 50)
            steps.append(("sym", r, rand_payload(rng, cfg.dtype, (50, 50)), 50, 50, False, sx))
        steps.append(("commit",))
    steps += [("fill_lower",), ("commit",), ("reset",)]
    steps += slice_steps(cfg, rng, 3, 80, 70, commit_every=1)
    want = run_checker(cfg, steps)
    with _Env(CM_B200_BG_MIN_ELEMS=1 if policy == 2 else 1 << 24):
        got = run_cuda(cfg, steps, exchange_mode=mode, apply_policy=policy, defer_budget=budget)
    assert_blocks_match(cfg, got, want)


def deferred_apply_api():
    """the log is really deferred (no scatter-add launches at commit), accessors materialise, one-sided
    reads of a rank with a pending log are refused until cm_materialize, switching back to eager works"""
    cfg = Cfg(cm.F64, 512, 512, 2, 2, 64, 64, 256)
    rng = np.random.default_rng(152)
    world = cm.LoopbackWorld(4, 0)
    try:
        mats = world.run(lambda r: cm.Matrix(*cfg.args()))
        world.run(lambda r: (mats[r].set_exchange_mode(2), mats[r].set_apply_policy(1), mats[r].set_profiling(True)))
        assert world.on(0, lambda r: mats[r].apply_policy()) == 1
        dense = np.zeros((cfg.n, cfg.m))
        for rnd in range(3):
            for r in range(4):
                ix, jx = sorted_unique(rng, cfg.m, 100), sorted_unique(rng, cfg.n, 90)
                data = rand_payload(rng, cfg.dtype, (90, 100))
                dense[np.ix_(jx, ix)] += data
                world.on(r, lambda rr: mats[rr].update(data, ix, jx))
            world.run(lambda r: mats[r].commit())
        st = world.run(lambda r: mats[r].stats())
        assert all(s["apply_launches"] == 0 and s["elements_applied"] == 0 and s["commits"] == 3 for s in st), st
        # element (0, 0) lives on rank 0: its owner may read it, the others may not yet
        v = world.on(0, lambda r: mats[r].get(0, 0))
        assert np.isclose(v, dense[0, 0], rtol=1e-12, atol=1e-12)
        try:
            world.on(1, lambda r: mats[r].get(0, 0))
            raise AssertionError("one-sided read of a rank with a pending log must be refused")
        except cm.CmError:
            pass
        world.run(lambda r: mats[r].materialize())
        assert np.isclose(world.on(1, lambda r: mats[r].get(0, 0)), dense[0, 0], rtol=1e-12, atol=1e-12)
        st = world.run(lambda r: mats[r].stats())
        assert all(s["apply_launches"] == 1 for s in st), [s["apply_launches"] for s in st]  # ONE sweep for three commits
        blocks = world.run(lambda r: mats[r].local())
        for r in range(4):
            pr, pc = r // 2, r % 2
            gi = np.array([(l // 64) * 128 + pr * 64 + l % 64 for l in range(256)])
            gj = np.array([(l // 64) * 128 + pc * 64 + l % 64 for l in range(256)])
            got = blocks[r].reshape(-1, cfg.lld)[:256, :256]
            assert np.linalg.norm(got - dense[np.ix_(gj, gi)]) <= 1e-12 * np.linalg.norm(dense)
        world.run(lambda r: mats[r].set_apply_policy(0))
        for r in range(4):
            world.on(r, lambda rr: mats[rr].update(np.ones((4, 4)), np.arange(4, dtype=np.int64), np.arange(4, dtype=np.int64)))
        world.run(lambda r: mats[r].commit())
        assert np.isclose(world.on(3, lambda r: mats[r].get(0, 0)), dense[0, 0] + 4.0)
        world.run(lambda r: mats[r].destroy())
    finally:
        world.close()


CASES += [("deferred_apply", [(2, 0), (2, 40000), (2, 1), (2, 0, 2), (2, 40000, 2)]),
          ("deferred_apply_api", [()])]


def two_matrices_and_batch_api():
    """two instances alive at once with interleaved updates and commits (arenas, flags and logs are per
    instance), fed through the batch forms of the host and device APIs"""
    cfgA = Cfg(cm.F64, 700, 500, 2, 2, 32, 32, 384)
    cfgB = Cfg(cm.F32, 300, 900, 2, 2, 16, 64, 160)
    rng = np.random.default_rng(161)
    world = cm.LoopbackWorld(4, 0)
    try:
        A = world.run(lambda r: cm.Matrix(*cfgA.args()))
        B = world.run(lambda r: cm.Matrix(*cfgB.args()))
        world.run(lambda r: B[r].set_apply_policy(1))
        stepsA, stepsB, keep = [], [], []
        for rnd in range(3):
            for r in range(4):
                for cfg, mats, steps, device in ((cfgA, A, stepsA, False), (cfgB, B, stepsB, True)):
                    cnt, m_u, n_u = 3, 40 + 10 * rnd, 30
                    data = np.ascontiguousarray(rand_payload(rng, cfg.dtype, (cnt, n_u, m_u)))
                    ix = np.ascontiguousarray(np.stack([sorted_unique(rng, cfg.m, m_u) for _ in range(cnt)]))
                    jx = np.ascontiguousarray(np.stack([sorted_unique(rng, cfg.n, n_u) for _ in range(cnt)]))
                    keep.append((data, ix, jx))
                    for u in range(cnt):
                        steps.append(("slice", r, data[u], m_u, n_u, m_u, False, ix[u], jx[u]))
                    batch = cm.SliceBatch(cnt, m_u, n_u, data.ctypes.data, ix.ctypes.data, jx.ctypes.data, data.itemsize)
                    world.on(r, lambda rr: mats[rr].update_batch(batch, device))
            world.run(lambda r: A[r].commit())
            stepsA.append(("commit",))
            if rnd != 1:  # B commits less often than A
                world.run(lambda r: B[r].commit())
                stepsB.append(("commit",))
        world.run(lambda r: B[r].commit())
        stepsB.append(("commit",))
        for cfg, mats, steps in ((cfgA, A, stepsA), (cfgB, B, stepsB)):
            got = {"local": world.run(lambda r: mats[r].local()), "m_local": world.run(lambda r: mats[r].m_local),
                   "n_local": world.run(lambda r: mats[r].n_local)}
            assert_blocks_match(cfg, got, run_checker(cfg, steps))
        world.run(lambda r: (A[r].destroy(), B[r].destroy()))
    finally:
        world.close()


CASES += [("two_matrices_and_batch_api", [()])]


def deterministic_mode_under_variants():
    """deterministic mode stays bit-exact run to run under column panels, the deferred and the background
    policy (what is applied in which scatter-add depends only on the sequence of calls)"""
    from scenario import run_cuda
    cfg = Cfg(cm.F64, 640, 640, 2, 2, 32, 32, 320)
    rng = np.random.default_rng(191)
    steps = slice_steps(cfg, rng, 12, 200, 200, commit_every=3)  # ~2.3 hits per element per rank: order matters
    want = run_checker(cfg, steps)
    for env, kw in (({"CM_B200_PIPELINE": 1, "CM_B200_PIPELINE_MIN_ELEMS": 1}, {}),
                    ({}, {"apply_policy": 1}), ({"CM_B200_BG_MIN_ELEMS": 1}, {"apply_policy": 2})):
        with _Env(**env):
            a = run_cuda(cfg, steps, exchange_mode=2, deterministic=True, **kw)
            b = run_cuda(cfg, steps, exchange_mode=2, deterministic=True, **kw)
        for r in range(cfg.P):
            assert a["local"][r].tobytes() == b["local"][r].tobytes(), (env, kw, r)
        assert_blocks_match(cfg, a, want)


CASES += [("deterministic_mode_under_variants", [()])]


def k1_full_size_properties():
    """BASELINE K1 geometry at full size (N = 8192, 256 x 256 all-ones updates, device API, asynchronous
    threshold flushes through both scatter-add kernels): every element lands exactly once, per-row hit
    counts, exact cancellation by the negated updates (tests/test_gpu_parity.py::test_full_size_properties_k1
    restated with numpy)"""
    cm.init(0, 1, 0)
    try:
        M = cm.Matrix(cm.F64, 8192, 8192, 1, 1, 64, 64, 8192)
        rng = np.random.default_rng(1)
        U = 512
        ones = np.ones((256, 256))
        ix = np.stack([np.sort(rng.choice(8192, 256, replace=False)) for _ in range(U)]).astype(np.int64)
        jx = np.stack([np.sort(rng.choice(8192, 256, replace=False)) for _ in range(U)]).astype(np.int64)
        M.set_flush_threshold(1 << 23)  # 129 updates per threshold flush (density 0.126: tiles), 125 in the last one (0.122: RED)
        for sign in (ones, -ones):
            for u in range(U):
                M.update_device(_ptr(sign), 256, 256, 256, False, _ptr(ix[u]), _ptr(jx[u]))
            M.commit()
            blk = M.local().reshape(8192, 8192)
            if sign is ones:
                assert blk.sum() == U * 256 * 256
                rows = np.zeros(8192)
                np.add.at(rows, ix.reshape(-1), 256.0)
                assert np.array_equal(blk.sum(axis=0), rows)
                st = M.stats()
                assert st["tile_launches"] >= 1 and st["flushes"] >= 2
            else:
                assert not blk.any()
        M.destroy()
    finally:
        cm.finalize()


CASES += [("k1_full_size_properties", [()])]


def bench_step_helpers():
    """bench.py's workload plumbing (make_batches / run_step, eager and deferred with the final
    cm_materialize) on a scaled-down K2-shaped spec, against a dense numpy sum"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    spec = dict(name="mini-K2", dtype=0, m=1024, n=1024, nprow=2, npcol=2, mb=64, nb=64, lld=512, m_u=96, n_u=80, U=12,
                commit_every=4)
    rng = np.random.default_rng(211)
    for deferred in (False, True):
        world = cm.LoopbackWorld(4, 0)
        try:
            mats = world.run(lambda r: cm.Matrix(spec["dtype"], spec["m"], spec["n"], spec["nprow"], spec["npcol"], spec["mb"],
                                                 spec["nb"], spec["lld"]))
            sp = dict(spec, deferred=deferred)
            if deferred:
                world.run(lambda r: mats[r].set_apply_policy(1, 1 << 30))
            dense = np.zeros((spec["n"], spec["m"]))
            inputs = []
            for r in range(4):
                data = np.ascontiguousarray(rng.uniform(-1, 1, size=(spec["U"], spec["n_u"], spec["m_u"])))
                ix = np.ascontiguousarray(np.stack([sorted_unique(rng, spec["m"], spec["m_u"]) for _ in range(spec["U"])]))
                jx = np.ascontiguousarray(np.stack([sorted_unique(rng, spec["n"], spec["n_u"]) for _ in range(spec["U"])]))
                for u in range(spec["U"]):
                    dense[np.ix_(jx[u], ix[u])] += data[u]
                inputs.append((data, ix, jx))
            batches = [bench.make_batches(cm, sp, _ptr(d), _ptr(i), _ptr(j), 8, True) for d, i, j in inputs]
            assert len(batches[0]) == 3  # commit every 4 of 12 updates
            world.run(lambda r: bench.run_step(mats[r], batches[r], sp))
            blocks = world.run(lambda r: mats[r].local())
            for r in range(4):
                pr, pc = r // 2, r % 2
                gi = np.array([(l // 64) * 128 + pr * 64 + l % 64 for l in range(512)])
                gj = np.array([(l // 64) * 128 + pc * 64 + l % 64 for l in range(512)])
                got = blocks[r].reshape(-1, spec["lld"])[:512, :512]
                assert np.linalg.norm(got - dense[np.ix_(gj, gi)]) <= 1e-12 * np.linalg.norm(dense), (deferred, r)
            world.run(lambda r: mats[r].destroy())
        finally:
            world.close()


CASES += [("bench_step_helpers", [()])]
