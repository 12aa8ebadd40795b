"""Scenario runner shared by the parity tests: the same list of steps is executed by the
CUDA path (through the C ABI, include/cm_b200.h) and by a checker (oracle/), and the local
blocks / wire streams are compared.

A step is a tuple:
  ("slice", rank, data, m_u, n_u, ld, trans, ix, jx)   update(Mat, ix, jx)        :604-615
  ("sym",   rank, data, n_u, ld, trans, ix)            update(Mat, ix)            :641-680
  ("elems", rank, vals, is, js)                        update(T, i, j) x n        :623-629
  ("fill_lower",) ("commit",) ("reset",)               collectives                :694, :723, :491
"""
import numpy as np

import cm_b200 as cm
from oracle_bindings import Oracle, RefShim

TOL = {cm.F64: 1e-12, cm.F32: 1e-5, cm.I32: 0.0}  # north_star: normwise relative


class Cfg:
    def __init__(self, dtype, m, n, nprow, npcol, mb, nb, lld):
        self.dtype, self.m, self.n = dtype, m, n
        self.nprow, self.npcol, self.mb, self.nb, self.lld = nprow, npcol, mb, nb, lld
        self.P = nprow * npcol

    def args(self):
        return (self.dtype, self.m, self.n, self.nprow, self.npcol, self.mb, self.nb, self.lld)

    def __repr__(self):
        return "Cfg" + repr(self.args())


def run_checker(cfg, steps, kind="oracle", wire=False, threshold=None):
    cls = Oracle if kind == "oracle" else RefShim
    o = cls(*cfg.args(), record_wire=wire)
    if threshold is not None:
        o.set_flush_threshold(threshold)
    for st in steps:
        op = st[0]
        if op == "slice":
            _, r, data, m_u, n_u, ld, trans, ix, jx = st
            o.update_slice(r, data, m_u, n_u, ld, trans, ix, jx)
        elif op == "sym":
            _, r, data, n_u, ld, trans, ix = st
            o.update_sym(r, data, n_u, ld, trans, ix)
        elif op == "elems":
            _, r, vals, is_, js = st
            if kind == "oracle":
                for v, i, j in zip(vals, is_, js):
                    o.update_elem(r, v, int(i), int(j))
            else:
                o.update_elems(r, vals, is_, js)
        elif op == "fill_lower":
            o.fill_lower()
        elif op == "commit":
            o.commit()
        elif op == "reset":
            o.reset()
        else:
            raise ValueError(op)
    out = {"local": [o.local(r) for r in range(cfg.P)],
           "m_local": [o.m_local(r) for r in range(cfg.P)],
           "n_local": [o.n_local(r) for r in range(cfg.P)]}
    if wire:
        out["wire"] = {(s, d): o.wire(s, d) for s in range(cfg.P) for d in range(cfg.P)}
    o.close()
    return out


def torch_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


class _HostBuf:
    """stand-in for a device tensor under the kernel-logic test library"""

    def __init__(self, a, dt):
        self.a = np.ascontiguousarray(a, dtype=dt).copy()

    def data_ptr(self):
        return self.a.ctypes.data

    def is_floating_point(self):
        return self.a.dtype.kind == "f"

    def fill_(self, v):
        self.a[...] = v


def run_cuda(cfg, steps, wire=False, threshold=None, sorted_sweep=True, deterministic=False, device_api=False,
             apply_kernel=0, exchange_mode=0, apply_policy=0, defer_budget=0, staging_limit=None, flush_every=0,
             free_after_flush=False):
    """Execute on the GPU. P == 1 uses the process runtime, P > 1 the loopback transport
    (threads-as-ranks on one device). device_api=True feeds slices through
    cm_update_slice_device from torch tensors already resident in HBM."""
    P = cfg.P
    world = None
    if P == 1:
        cm.init(0, 1, 0)
        run = lambda fn: [fn(0)]
        on = lambda r, fn: fn(r)
    else:
        world = cm.LoopbackWorld(P, 0)
        run, on = world.run, world.on
    keep = []  # device tensors borrowed until commit
    keep_r = {r: [] for r in range(P)}  # free_after_flush: ... or until the issuing rank's next flush
    nflush = [0] * P
    nslice = [0] * P
    try:
        mats = run(lambda r: cm.Matrix(*cfg.args()))

        def setup(r):
            M = mats[r]
            if wire:
                M.keep_wire(True)
            if threshold is not None:
                M.set_flush_threshold(threshold)
            M.set_sorted_sweep(sorted_sweep)
            M.set_deterministic(deterministic)
            M.set_apply_kernel(apply_kernel)
            if exchange_mode:
                M.set_exchange_mode(exchange_mode)
            if apply_policy:
                M.set_apply_policy(apply_policy, defer_budget)
            if staging_limit:
                M.set_staging_limit(staging_limit)
        run(setup)
        wires = {}
        for st in steps:
            op = st[0]
            if op == "slice":
                _, r, data, m_u, n_u, ld, trans, ix, jx = st
                if device_api and not torch_cuda():
                    # kernel-logic test library (tests/emu): "device" memory is host memory
                    d, di, dj = _HostBuf(data, cm.NP_DTYPE[cfg.dtype]), _HostBuf(ix, np.int64), _HostBuf(jx, np.int64)
                if device_api and torch_cuda():
                    import torch
                    d = torch.from_numpy(np.ascontiguousarray(data, dtype=cm.NP_DTYPE[cfg.dtype])).cuda()
                    di = torch.from_numpy(np.ascontiguousarray(ix, dtype=np.int64)).cuda()
                    dj = torch.from_numpy(np.ascontiguousarray(jx, dtype=np.int64)).cuda()
                    torch.cuda.synchronize()
                if device_api:
                    (keep_r[r] if free_after_flush else keep).extend([d, di, dj])
                    on(r, lambda rr: mats[rr].update_device(d.data_ptr(), m_u, n_u, ld, trans, di.data_ptr(), dj.data_ptr()))
                    if free_after_flush:
                        # the borrow ends when a flush returns (include/cm_b200.h): scribble over what was lent
                        n = on(r, lambda rr: mats[rr].stats()["flushes"])
                        if n > nflush[r]:
                            nflush[r] = n
                            for t in keep_r[r]:
                                t.fill_(float("nan") if t.is_floating_point() else -7)
                            if torch_cuda():
                                torch.cuda.synchronize()
                            keep_r[r] = []
                else:
                    on(r, lambda rr: mats[rr].update(data, ix, jx, m_u, n_u, ld, trans))
                nslice[r] += 1
                if flush_every and nslice[r] % flush_every == 0:
                    on(r, lambda rr: mats[rr].flush())
            elif op == "sym":
                _, r, data, n_u, ld, trans, ix = st
                on(r, lambda rr: mats[rr].update_sym(data, ix, ld, trans))
            elif op == "elems":
                _, r, vals, is_, js = st
                on(r, lambda rr: mats[rr].update_elems(vals, is_, js))
            elif op == "fill_lower":
                run(lambda r: mats[r].fill_lower())
            elif op == "commit":
                run(lambda r: mats[r].commit())
                keep.clear()
                keep_r = {r: [] for r in range(P)}
                nflush = run(lambda r: mats[r].stats()["flushes"])
                if wire:
                    for s in range(P):
                        for d in range(P):
                            inds, data = on(s, lambda rr: mats[rr].wire(d))
                            if (s, d) in wires:
                                wires[(s, d)] = (np.concatenate([wires[(s, d)][0], inds]),
                                                 np.concatenate([wires[(s, d)][1], data]))
                            else:
                                wires[(s, d)] = (inds, data)
            elif op == "reset":
                run(lambda r: mats[r].reset())
            else:
                raise ValueError(op)
        out = {"local": run(lambda r: mats[r].local()),
               "m_local": run(lambda r: mats[r].m_local),
               "n_local": run(lambda r: mats[r].n_local),
               "stats": run(lambda r: mats[r].stats())}
        if wire:
            out["wire"] = wires
        run(lambda r: mats[r].destroy())
        return out
    finally:
        if world is not None:
            world.close()
        else:
            cm.finalize()


def normwise_rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    num = np.linalg.norm(a - b)
    return num / den if den > 0 else num


def assert_blocks_match(cfg, got, want):
    for r in range(cfg.P):
        assert got["m_local"][r] == want["m_local"][r] and got["n_local"][r] == want["n_local"][r]
        a, b = got["local"][r], want["local"][r]
        assert a.shape == b.shape
        if cfg.dtype == cm.I32:
            assert np.array_equal(a, b), f"rank {r}: integer blocks differ"
        else:
            err = normwise_rel(a, b)
            assert err <= TOL[cfg.dtype], f"rank {r}: normwise rel err {err:.3e} > {TOL[cfg.dtype]:.0e}"


def assert_wire_match(cfg, got, want):
    """bit-exact owner mapping and packing (north_star)"""
    for s in range(cfg.P):
        for d in range(cfg.P):
            gi, gd = got["wire"].get((s, d), (np.zeros(0, np.int64), np.zeros(0)))
            wi, wd = want["wire"][(s, d)]
            assert len(gi) == len(wi), f"wire {s}->{d}: {len(gi)} elements, want {len(wi)}"
            assert np.array_equal(gi, wi), f"wire {s}->{d}: index stream differs"
            assert gd.tobytes() == wd.tobytes(), f"wire {s}->{d}: value stream differs"


# ---- seeded input generators -----------------------------------------------------------
def rand_payload(rng, dtype, shape):
    if dtype == cm.I32:
        return rng.integers(-9, 10, size=shape).astype(np.int32)
    return rng.uniform(-1.0, 1.0, size=shape).astype(cm.NP_DTYPE[dtype])


def sorted_unique(rng, hi, k):
    return np.sort(rng.choice(hi, size=k, replace=False)).astype(np.int64)


def slice_steps(cfg, rng, nupd, m_u, n_u, ranks=None, commit_every=None):
    steps = []
    ranks = list(range(cfg.P)) if ranks is None else ranks
    for u in range(nupd):
        for r in ranks:
            ix = sorted_unique(rng, cfg.m, m_u)
            jx = sorted_unique(rng, cfg.n, n_u)
            data = rand_payload(rng, cfg.dtype, (n_u, m_u))  # column-major m_u x n_u
            steps.append(("slice", r, data, m_u, n_u, m_u, False, ix, jx))
        if commit_every and (u + 1) % commit_every == 0:
            steps.append(("commit",))
    steps.append(("commit",))
    return steps
