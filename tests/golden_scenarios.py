"""Seeded scenarios behind tests/golden/*.npz (shared by the generator and the tests)."""
import numpy as np

import cm_b200 as cm
from scenario import Cfg, rand_payload, sorted_unique

# cfg = (dtype, m, n, nprow, npcol, mb, nb, lld). Square grids only where fill_lower / m_local
# are involved: the reference ctor mis-computes the grid row for NPROW != NPCOL.
# threshold: the reference's bin_flush_threshold for the run (None = its default 10000); the
# wire-stream comparison needs one message per (source, owner, commit), i.e. a huge threshold.
SCENARIOS = {
    "k0_slices_f32": dict(cfg=(cm.F32, 160, 120, 2, 2, 64, 64, 1024), seed=101, kind="slices", nupd=2, m_u=50, n_u=40,
                          threshold=1 << 30),
    "k0_slices_f64": dict(cfg=(cm.F64, 160, 120, 2, 2, 64, 64, 1024), seed=102, kind="slices", nupd=2, m_u=50, n_u=40,
                          threshold=1 << 30),
    "k0_slices_i32": dict(cfg=(cm.I32, 160, 120, 2, 2, 64, 64, 1024), seed=103, kind="slices", nupd=2, m_u=50, n_u=40,
                          threshold=1 << 30),
    "default_threshold_f64": dict(cfg=(cm.F64, 160, 120, 2, 2, 64, 64, 1024), seed=104, kind="slices", nupd=3, m_u=150,
                                  n_u=110, threshold=None),
    # both sides at the reference's default threshold: every bin / every rank ships from inside update() (:346, :614),
    # several messages per (source, owner) and commit; the concatenated streams must still be the reference's
    "threshold_flush_f64": dict(cfg=(cm.F64, 160, 120, 2, 2, 64, 64, 1024), seed=110, kind="slices", nupd=6, m_u=150,
                                n_u=110, commit_every=3, threshold=10000, ours_threshold=10000),
    "single_rank_f64": dict(cfg=(cm.F64, 60, 150, 1, 1, 4, 8, 64), seed=105, kind="slices", nupd=4, m_u=30, n_u=40,
                            threshold=1 << 30),
    "nonsquare_grid_f64": dict(cfg=(cm.F64, 100, 190, 2, 3, 4, 8, 64), seed=109, kind="slices", nupd=2, m_u=40, n_u=50,
                               threshold=1 << 30),
    "elements_f32": dict(cfg=(cm.F32, 160, 120, 2, 2, 64, 64, 1024), seed=106, kind="elements", n=1500, threshold=1 << 30),
    "symmetric_fill_f64": dict(cfg=(cm.F64, 150, 150, 2, 2, 64, 64, 1024), seed=107, kind="symmetric", nupd=2, n_u=60,
                               threshold=1 << 30),
    "edge_cases_f64": dict(cfg=(cm.F64, 160, 120, 2, 2, 64, 64, 1024), seed=108, kind="edges", threshold=1 << 30),
}


def build_steps(spec):
    cfg = Cfg(*spec["cfg"])
    rng = np.random.default_rng(spec["seed"])
    steps = []
    kind = spec["kind"]
    if kind == "slices":
        for u in range(spec["nupd"]):
            for r in range(cfg.P):
                m_u, n_u = min(spec["m_u"], cfg.m), min(spec["n_u"], cfg.n)
                ix, jx = sorted_unique(rng, cfg.m, m_u), sorted_unique(rng, cfg.n, n_u)
                steps.append(("slice", r, rand_payload(rng, cfg.dtype, (n_u, m_u)), m_u, n_u, m_u, False, ix, jx))
            if (u + 1) % spec.get("commit_every", 1) == 0 or u + 1 == spec["nupd"]:
                steps.append(("commit",))
    elif kind == "elements":
        for it in range(2):
            for r in range(cfg.P):
                n = spec["n"]
                steps.append(("elems", r, rand_payload(rng, cfg.dtype, n), rng.integers(0, cfg.m, n), rng.integers(0, cfg.n, n)))
            steps.append(("commit",))
    elif kind == "symmetric":
        for u in range(spec["nupd"]):
            for r in range(cfg.P):
                ix = sorted_unique(rng, cfg.m, spec["n_u"])
                # dyadic payloads: sums are exact in any order, so the values fill_lower puts on the
                # wire (read back from the accumulated block) are bit-identical on every implementation
                pay = (rng.integers(-64, 65, size=(spec["n_u"], spec["n_u"])) / 8.0).astype(cm.NP_DTYPE[cfg.dtype])
                steps.append(("sym", r, pay, spec["n_u"], spec["n_u"], False, ix))
            steps.append(("commit",))
        steps += [("fill_lower",), ("commit",)]
    elif kind == "edges":
        ix = rng.integers(0, cfg.m, 70).astype(np.int64)  # unsorted, with duplicates
        jx = rng.integers(0, cfg.n, 50).astype(np.int64)
        ix[5] = ix[6] = ix[7]
        jx[1] = jx[2]
        steps.append(("slice", 0, rand_payload(rng, cfg.dtype, (50, 70)), 70, 50, 70, False, ix, jx))
        ix, jx = sorted_unique(rng, cfg.m, 33), sorted_unique(rng, cfg.n, 21)
        steps.append(("slice", 1, rand_payload(rng, cfg.dtype, (21, 40)), 33, 21, 40, False, ix, jx))  # ld > m_u
        steps.append(("slice", 2, rand_payload(rng, cfg.dtype, (33, 25)), 33, 21, 25, True, ix, jx))   # transposed view
        one = np.array([3], dtype=np.int64)
        steps.append(("slice", 3, rand_payload(rng, cfg.dtype, (1, 1)), 1, 1, 1, False, one, one))
        steps.append(("slice", 3, rand_payload(rng, cfg.dtype, (21, 1)), 1, 21, 1, False, one, jx))
        steps.append(("slice", 0, rand_payload(rng, cfg.dtype, (1, 33)), 33, 1, 33, False, ix, one))
        steps += [("commit",), ("commit",)]
    else:
        raise ValueError(kind)
    return steps


def load_golden(path):
    z = np.load(path)
    cfg = Cfg(*[int(x) for x in z["cfg"]])
    out = {"local": [z[f"local_{r}"] for r in range(cfg.P)],
           "m_local": [int(z[f"m_local_{r}"]) for r in range(cfg.P)],
           "n_local": [int(z[f"n_local_{r}"]) for r in range(cfg.P)]}
    if "wire_inds_0_0" in z.files:
        out["wire_inds"] = {(s, d): z[f"wire_inds_{s}_{d}"].astype(np.int64) for s in range(cfg.P) for d in range(cfg.P)}
        out["wire_sha"] = {(s, d): bytes(z[f"wire_sha_{s}_{d}"]) for s in range(cfg.P) for d in range(cfg.P)}
    return cfg, out


def assert_wire_matches_golden(cfg, got, golden, multiset=False):
    """bit-exact: index streams equal, value streams hash to the reference's SHA-256.
    multiset=True: same elements per (source, owner) but not the same order — used for
    fill_lower, whose elements this implementation emits column-major in the DESTINATION
    orientation (a transposed slice update) while the reference walks its local block
    column-major in the SOURCE orientation (:696-709); the values are covered by the block check."""
    import hashlib
    for s in range(cfg.P):
        for d in range(cfg.P):
            gi, gd = got["wire"].get((s, d), (np.zeros(0, np.int64), np.zeros(0)))
            if multiset:
                assert np.array_equal(np.sort(gi), np.sort(golden["wire_inds"][(s, d)])), \
                    f"wire {s}->{d}: index multiset differs from the reference"
                continue
            assert np.array_equal(gi, golden["wire_inds"][(s, d)]), f"wire {s}->{d}: index stream differs from the reference"
            assert hashlib.sha256(gd.tobytes()).digest() == golden["wire_sha"][(s, d)], \
                f"wire {s}->{d}: value stream differs from the reference"
