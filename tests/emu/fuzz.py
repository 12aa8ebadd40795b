"""Seeded fuzzer for the kernel-logic TEST library (tests/emu/libcm_b200_emu.so, see run_cases.py):
random distribution geometries, update shapes, index-set styles, API mixes and library knobs, every
scenario checked against the oracle (local blocks within the north_star tolerance, wire streams
bit-exact where the index sets make the comparison well defined).

usage: python tests/emu/fuzz.py [--seed S] [--count N] [--verbose]
"""
import argparse
import os
import sys
import time
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import cm_b200 as cm  # noqa: E402

cm.LIB_PATH = os.environ.get("CM_EMU_LIB", os.path.join(HERE, "libcm_b200_emu.so"))  # e.g. a sanitizer build

from scenario import (Cfg, assert_blocks_match, assert_wire_match, rand_payload, run_checker, run_cuda,  # noqa: E402
                      sorted_unique)

ENV_KEYS = ("CM_B200_PIPELINE", "CM_B200_PIPELINE_MIN_ELEMS", "CM_B200_PANELS", "CM_B200_BG_MIN_ELEMS")


def index_set(rng, hi, k, style):
    if style == "sorted":
        return sorted_unique(rng, hi, min(k, hi))
    if style == "runs":  # runs of consecutive indices (clustered, K3-like)
        out = []
        while len(out) < k:
            s = int(rng.integers(0, hi))
            out.extend(range(s, min(hi, s + int(rng.integers(1, 40)))))
        return np.unique(np.array(out[:k], dtype=np.int64))
    return rng.integers(0, hi, size=k).astype(np.int64)  # unsorted, duplicates allowed


TALL_P = 0.15  # share of scenarios whose local block is taller than one shared-memory tile (--tall)


def make_scenario(rng):
    nprow, npcol = [(1, 1), (1, 2), (2, 1), (2, 2), (2, 3), (3, 2), (1, 4), (2, 4), (4, 2)][int(rng.integers(0, 9))]
    mb, nb = int(rng.choice([1, 3, 8, 16, 64])), int(rng.choice([1, 4, 8, 16, 64]))
    dtype = int(rng.choice([cm.F64, cm.F64, cm.F32, cm.I32]))
    tall = rng.random() < TALL_P  # taller than one shared-memory tile (8192 fp64 / 16384 fp32 rows)
    m = int(rng.integers(9000, 20000)) * nprow if tall else int(rng.integers(max(mb * nprow, 2), 1500))
    n = int(rng.integers(max(nb * npcol, 2), max(nb * npcol, 2) + (200 if tall else 1200)))
    # every rank needs a non-empty local block (reference asserts, :422-427)
    m = max(m, mb * nprow)
    n = max(n, nb * npcol)
    m_loc_max = -(-m // (mb * nprow)) * mb
    lld = m_loc_max + int(rng.integers(0, 17))
    cfg = Cfg(dtype, m, n, nprow, npcol, mb, nb, lld)
    square = rng.random() < 0.25 and not tall and m >= nb * npcol
    if square:
        cfg = Cfg(dtype, m, m, nprow, npcol, mb, nb, lld)
    style = str(rng.choice(["sorted", "sorted", "runs", "unsorted"]))
    steps = []
    rounds = int(rng.integers(1, 4))
    for _ in range(rounds):
        for r in range(cfg.P):
            for _ in range(int(rng.integers(0, 4))):
                kind = rng.random()
                if kind < 0.7:
                    m_u = int(rng.integers(1, min(cfg.m, 700 if tall else 320) + 1))
                    n_u = int(rng.integers(1, min(cfg.n, 60 if tall else 200) + 1))
                    ix, jx = index_set(rng, cfg.m, m_u, style), index_set(rng, cfg.n, n_u, style)
                    m_u, n_u = len(ix), len(jx)
                    trans = rng.random() < 0.2
                    rows_s, cols_s = (n_u, m_u) if trans else (m_u, n_u)
                    ld = rows_s + int(rng.integers(0, 5))
                    steps.append(("slice", r, rand_payload(rng, dtype, (cols_s, ld)), m_u, n_u, ld, trans, ix, jx))
                elif kind < 0.85 and square:
                    k = int(rng.integers(1, min(cfg.m, 150) + 1))
                    ix = index_set(rng, cfg.m, k, style)
                    k = len(ix)
                    steps.append(("sym", r, rand_payload(rng, dtype, (k, k)), k, k, False, ix))
                else:
                    k = int(rng.integers(1, 200))
                    steps.append(("elems", r, rand_payload(rng, dtype, k), rng.integers(0, cfg.m, k), rng.integers(0, cfg.n, k)))
        if square and rng.random() < 0.3:
            steps += [("commit",), ("fill_lower",)]
        steps.append(("commit",))
        if rng.random() < 0.15:
            steps.append(("reset",))
    knobs = dict(apply_kernel=int(rng.choice([0, 1, 2])), sorted_sweep=bool(rng.random() < 0.8),
                 deterministic=bool(rng.random() < 0.25), exchange_mode=int(rng.choice([0, 1, 2])))
    if cfg.P > 1 and knobs["exchange_mode"] != 1 and rng.random() < 0.35:  # deferred scatter-add (peer-to-peer modes)
        knobs["apply_policy"] = int(rng.choice([1, 2]))
        knobs["defer_budget"] = int(rng.choice([0, 1, 20000, 400000]))
    if cfg.P == 1:
        knobs["exchange_mode"] = 0
    # threshold flushes on any number of ranks (reference :346, :614), explicit cm_flush calls, and the limit of
    # staged updates at which a rank flushes by itself
    if rng.random() < 0.5:
        knobs["threshold"] = int(rng.integers(1, 60000))
    if rng.random() < 0.15:
        knobs["flush_every"] = int(rng.integers(1, 4))
    if rng.random() < 0.15:
        knobs["staging_limit"] = int(rng.integers(1, 4))
    env = {}
    if cfg.P > 1 and rng.random() < 0.6:
        env = {"CM_B200_PIPELINE": str(int(rng.random() < 0.8)), "CM_B200_PIPELINE_MIN_ELEMS": str(int(rng.choice([1, 2000, 50000]))),
               "CM_B200_PANELS": str(int(rng.integers(1, 9)))}
    # stream comparison is defined for slice-only scenarios with panel-monotone index sets: elemental
    # updates travel in a lane of their own here (the reference interleaves them in the same bin) and
    # the symmetric update / fill_lower are masked on the owner
    if knobs.get("apply_policy") == 2:
        env = dict(env, CM_B200_BG_MIN_ELEMS=str(int(rng.choice([1, 5000]))))
    # (any index style, panels or not: an update whose columns are not panel-monotone stays in panel 0)
    wire = all(s[0] in ("slice", "commit", "reset") for s in steps) and rng.random() < 0.8
    return cfg, steps, knobs, env, wire


def run_one(seed, verbose=False):
    rng = np.random.default_rng(seed)
    cfg, steps, knobs, env, wire = make_scenario(rng)
    if verbose:
        print(f"seed {seed}: {cfg} steps={len(steps)} knobs={knobs} env={env} wire={wire}", flush=True)
    old = {k: os.environ.get(k) for k in ENV_KEYS}
    for k in ENV_KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
    try:
        want = run_checker(cfg, steps, wire=wire, threshold=(1 << 40) if wire else None)
        got = run_cuda(cfg, steps, wire=wire, **knobs)
        if wire:
            assert_wire_match(cfg, got, want)
        assert_blocks_match(cfg, got, want)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return cfg, knobs, env


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--count", type=int, default=50)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--tall", type=float, default=0.15, help="share of tall local blocks")
    a = ap.parse_args()
    global TALL_P
    TALL_P = a.tall
    fails = 0
    t0 = time.time()
    for s in range(a.seed, a.seed + a.count):
        try:
            run_one(s, a.verbose)
        except Exception as e:  # noqa: BLE001
            fails += 1
            print(f"FAIL seed {s}: {type(e).__name__}: {str(e)[:300]}", flush=True)
            if a.verbose:
                traceback.print_exc(limit=8)
    print(f"fuzz: {a.count} scenarios from seed {a.seed}, fails: {fails}, {time.time() - t0:.1f}s", flush=True)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
