"""ctypes binding of libcm_b200.so (include/cm_b200.h) for the test-suite and bench.py.

The product is the C-ABI library plus the C++ headers in include/; Python is only the
harness. Importing this module loads the library (no GPU needed for that); every call that
touches matrix data raises ``CmError`` when the CUDA path is unavailable — there is no CPU
fallback anywhere in this package, and nothing here imports ``oracle/``.

The directory is named after the project (``convergent-matrix_b200``), which is not a valid
Python identifier; ``cm_b200.py`` at the repo root re-exports it as ``import cm_b200``.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcm_b200.so")

F64, F32, I32 = 0, 1, 2
NP_DTYPE = {F64: np.float64, F32: np.float32, I32: np.int32}
MASK_NONE, MASK_UPPER, MASK_STRICT_LOWER = 0, 1, 2

_vp, _l, _i = C.c_void_p, C.c_long, C.c_int
_lp = C.POINTER(C.c_long)


class CmError(RuntimeError):
    pass


class Stats(C.Structure):
    _fields_ = [(n, C.c_long) for n in ("updates", "elements_staged", "elements_applied", "flushes",
                                        "commits", "kernel_launches", "bytes_sent", "bytes_received",
                                        "h2d_bytes", "d2h_bytes")] + \
               [("apply_ms", C.c_double), ("pack_ms", C.c_double), ("exchange_ms", C.c_double),
                ("apply_launches", C.c_long), ("apply_elements", C.c_long), ("tile_launches", C.c_long),
                ("flushed_shipped", C.c_long), ("flushed_held", C.c_long), ("staging_peak_bytes", C.c_long)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/cm_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "cm_version": (C.c_char_p, []),
    "cm_last_error": (C.c_char_p, []),
    "cm_rt_unique_id": (_i, [_vp, C.c_size_t]),
    "cm_rt_init": (_i, [_i, _i, _i, _vp]),
    "cm_rt_init_from_env": (_i, []),
    "cm_rt_init_loopback": (_i, [_i, _i]),
    "cm_rt_bind_rank": (_i, [_i]),
    "cm_rt_finalize": (_i, []),
    "cm_rt_rank_me": (_i, []),
    "cm_rt_rank_n": (_i, []),
    "cm_rt_barrier": (_i, []),
    "cm_rt_device": (_i, []),
    "cm_create": (_i, [C.POINTER(_vp), _i, _l, _l, _l, _l, _l, _l, _l]),
    "cm_destroy": (_i, [_vp]),
    "cm_update_slice": (_i, [_vp, _vp, _l, _l, _l, _i, _vp, _vp]),
    "cm_update_sym": (_i, [_vp, _vp, _l, _l, _i, _vp]),
    "cm_update_elem": (_i, [_vp, _vp, _l, _l]),
    "cm_update_elems": (_i, [_vp, _vp, _vp, _vp, _l]),
    "cm_fill_lower": (_i, [_vp]),
    "cm_update_slice_device": (_i, [_vp, _vp, _l, _l, _l, _i, _vp, _vp, _i]),
    "cm_update_slices_device": (_i, [_vp, _l, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "cm_update_slices": (_i, [_vp, _l, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cm_flush": (_i, [_vp]),
    "cm_commit": (_i, [_vp]),
    "cm_reset": (_i, [_vp]),
    "cm_local_data_device": (_i, [_vp, C.POINTER(_vp)]),
    "cm_local_data_host": (_i, [_vp, C.POINTER(_vp)]),
    "cm_local_data_copy": (_i, [_vp, _vp]),
    "cm_get": (_i, [_vp, _l, _l, _vp]),
    "cm_save": (_i, [_vp, C.c_char_p]),
    "cm_load": (_i, [_vp, C.c_char_p]),
    "cm_m": (_l, [_vp]),
    "cm_n": (_l, [_vp]),
    "cm_m_local": (_l, [_vp]),
    "cm_n_local": (_l, [_vp]),
    "cm_lld": (_l, [_vp]),
    "cm_pgrid_row": (_l, [_vp]),
    "cm_pgrid_col": (_l, [_vp]),
    "cm_dtype": (_i, [_vp]),
    "cm_get_flush_threshold": (_l, [_vp]),
    "cm_set_flush_threshold": (_i, [_vp, _l]),
    "cm_get_progress_interval": (_i, [_vp]),
    "cm_set_progress_interval": (_i, [_vp, _i]),
    "cm_set_deterministic": (_i, [_vp, _i]),
    "cm_set_sorted_sweep": (_i, [_vp, _i]),
    "cm_set_apply_kernel": (_i, [_vp, _i, C.c_double]),
    "cm_set_apply_policy": (_i, [_vp, _i, _l]),
    "cm_get_apply_policy": (_i, [_vp]),
    "cm_materialize": (_i, [_vp]),
    "cm_set_exchange_mode": (_i, [_vp, _i]),
    "cm_get_exchange_mode": (_i, [_vp]),
    "cm_descinit": (_i, [_vp, _i, C.POINTER(_i)]),
    "cm_get_stats": (_i, [_vp, C.POINTER(Stats)]),
    "cm_reset_stats": (_i, [_vp]),
    "cm_set_profiling": (_i, [_vp, _i]),
    "cm_stream": (_i, [_vp, C.POINTER(_vp)]),
    "cm_debug_keep_wire": (_i, [_vp, _i]),
    "cm_debug_set_staging_limit": (_i, [_vp, _l]),
    "cm_debug_wire_count": (_i, [_vp, _i, _lp]),
    "cm_debug_wire_expand": (_i, [_vp, _i, _vp, _vp]),
    "cm_numroc": (_l, [_l, _l, _l, _l]),
    "cm_owner_of": (None, [_l] * 7 + [C.POINTER(_i), _lp]),
    "cm_descinit_raw": (None, [_i, _l, _l, _l, _l, _l, C.POINTER(_i)]),
}

_lib = None


def lib():
    """Load libcm_b200.so; raises CmError (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CmError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise CmError(f"cm_b200 error {rc}: {lib().cm_last_error().decode()}")


def _np_ptr(a):
    return a.ctypes.data_as(_vp)


# ---- runtime -------------------------------------------------------------------------
def unique_id():
    buf = C.create_string_buffer(128)
    _check(lib().cm_rt_unique_id(buf, 128))
    return buf.raw


def init(rank=0, nranks=1, device=0, uid=None):
    _check(lib().cm_rt_init(rank, nranks, device, uid))


def init_loopback(nranks, device=0):
    _check(lib().cm_rt_init_loopback(nranks, device))


def bind_rank(rank):
    _check(lib().cm_rt_bind_rank(rank))


def finalize():
    _check(lib().cm_rt_finalize())


def rank_me():
    return lib().cm_rt_rank_me()


def rank_n():
    return lib().cm_rt_rank_n()


def barrier():
    _check(lib().cm_rt_barrier())


def numroc(m, np_, ip, mb):
    return lib().cm_numroc(m, np_, ip, mb)


def owner_of(i, j, nprow, npcol, mb, nb, lld):
    r, ij = _i(), _l()
    lib().cm_owner_of(i, j, nprow, npcol, mb, nb, lld, C.byref(r), C.byref(ij))
    return r.value, ij.value


def descinit_raw(ictxt, m, n, mb, nb, lld):
    d = (_i * 9)()
    lib().cm_descinit_raw(ictxt, m, n, mb, nb, lld, d)
    return list(d)


# ---- matrix --------------------------------------------------------------------------
class Matrix:
    """One rank's ConvergentMatrix (reference include/convergent_matrix.hpp:298-971)."""

    def __init__(self, dtype, m, n, nprow, npcol, mb, nb, lld):
        self.L = lib()
        self.dtype = dtype
        self.np_dtype = NP_DTYPE[dtype]
        h = _vp()
        _check(self.L.cm_create(C.byref(h), dtype, m, n, nprow, npcol, mb, nb, lld))
        self.h = h

    def destroy(self):
        if self.h:
            h, self.h = self.h, None
            _check(self.L.cm_destroy(h))

    # reference API -------------------------------------------------------------------
    def update(self, data, ix, jx, m_u=None, n_u=None, ld=None, trans=False):
        """update(Mat, ix, jx) with HOST arrays. `data` holds the payload storage
        column-major: element (i,j) = data.flat[j + i*ld] if trans else data.flat[i + ld*j]."""
        ix = np.ascontiguousarray(ix, dtype=np.int64)
        jx = np.ascontiguousarray(jx, dtype=np.int64)
        data = np.ascontiguousarray(data, dtype=self.np_dtype)
        m_u = len(ix) if m_u is None else m_u
        n_u = len(jx) if n_u is None else n_u
        if ld is None:
            ld = n_u if trans else m_u
        _check(self.L.cm_update_slice(self.h, _np_ptr(data), m_u, n_u, ld, int(trans), _np_ptr(ix), _np_ptr(jx)))

    def update_sym(self, data, ix, ld=None, trans=False):
        ix = np.ascontiguousarray(ix, dtype=np.int64)
        data = np.ascontiguousarray(data, dtype=self.np_dtype)
        n_u = len(ix)
        _check(self.L.cm_update_sym(self.h, _np_ptr(data), n_u, n_u if ld is None else ld, int(trans), _np_ptr(ix)))

    def update_elem(self, val, i, j):
        v = np.array([val], dtype=self.np_dtype)
        _check(self.L.cm_update_elem(self.h, _np_ptr(v), i, j))

    def update_elems(self, vals, is_, js):
        vals = np.ascontiguousarray(vals, dtype=self.np_dtype)
        is_ = np.ascontiguousarray(is_, dtype=np.int64)
        js = np.ascontiguousarray(js, dtype=np.int64)
        _check(self.L.cm_update_elems(self.h, _np_ptr(vals), _np_ptr(is_), _np_ptr(js), len(vals)))

    def update_device(self, d_data_ptr, m_u, n_u, ld, trans, d_ix_ptr, d_jx_ptr, mask=MASK_NONE):
        """Device-resident update; arguments are raw device addresses (e.g. tensor.data_ptr())."""
        _check(self.L.cm_update_slice_device(self.h, d_data_ptr, m_u, n_u, ld, int(trans), d_ix_ptr, d_jx_ptr, mask))

    def update_batch(self, batch, device, mask=MASK_NONE):
        """`batch` = SliceBatch (arrays of per-update pointers/dims built once, reused every step)."""
        if device:
            _check(self.L.cm_update_slices_device(self.h, batch.count, _np_ptr(batch.data), _np_ptr(batch.m_u),
                                                  _np_ptr(batch.n_u), _np_ptr(batch.ld), None, _np_ptr(batch.ix),
                                                  _np_ptr(batch.jx), mask))
        else:
            _check(self.L.cm_update_slices(self.h, batch.count, _np_ptr(batch.data), _np_ptr(batch.m_u),
                                           _np_ptr(batch.n_u), _np_ptr(batch.ld), None, _np_ptr(batch.ix),
                                           _np_ptr(batch.jx)))

    def fill_lower(self):
        _check(self.L.cm_fill_lower(self.h))

    def flush(self):
        _check(self.L.cm_flush(self.h))

    def commit(self):
        _check(self.L.cm_commit(self.h))

    def reset(self):
        _check(self.L.cm_reset(self.h))

    def get(self, i, j):
        v = np.zeros(1, dtype=self.np_dtype)
        _check(self.L.cm_get(self.h, i, j, _np_ptr(v)))
        return v[0]

    def save(self, fname):
        _check(self.L.cm_save(self.h, fname.encode()))

    def load(self, fname):
        _check(self.L.cm_load(self.h, fname.encode()))

    # accessors -----------------------------------------------------------------------
    m = property(lambda s: s.L.cm_m(s.h))
    n = property(lambda s: s.L.cm_n(s.h))
    m_local = property(lambda s: s.L.cm_m_local(s.h))
    n_local = property(lambda s: s.L.cm_n_local(s.h))
    lld = property(lambda s: s.L.cm_lld(s.h))
    pgrid_row = property(lambda s: s.L.cm_pgrid_row(s.h))
    pgrid_col = property(lambda s: s.L.cm_pgrid_col(s.h))

    def local(self):
        """get_local_data_copy(): flat column-major copy of the lld x n_local block."""
        out = np.empty(self.lld * self.n_local, dtype=self.np_dtype)
        _check(self.L.cm_local_data_copy(self.h, _np_ptr(out)))
        return out

    def local_device_ptr(self):
        p = _vp()
        _check(self.L.cm_local_data_device(self.h, C.byref(p)))
        return p.value

    def local_host_view(self):
        p = _vp()
        _check(self.L.cm_local_data_host(self.h, C.byref(p)))
        n = self.lld * self.n_local
        buf = (C.c_char * (n * np.dtype(self.np_dtype).itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=self.np_dtype)

    def descinit(self, ictxt=0):
        d = (_i * 9)()
        _check(self.L.cm_descinit(self.h, ictxt, d))
        return list(d)

    # knobs ---------------------------------------------------------------------------
    def set_flush_threshold(self, elements):
        _check(self.L.cm_set_flush_threshold(self.h, elements))

    def flush_threshold(self):
        return self.L.cm_get_flush_threshold(self.h)

    def set_progress_interval(self, k):
        _check(self.L.cm_set_progress_interval(self.h, k))

    def progress_interval(self):
        return self.L.cm_get_progress_interval(self.h)

    def set_deterministic(self, on):
        _check(self.L.cm_set_deterministic(self.h, int(on)))

    def set_sorted_sweep(self, on):
        _check(self.L.cm_set_sorted_sweep(self.h, int(on)))

    def set_apply_kernel(self, kernel, tile_density=0.0):
        _check(self.L.cm_set_apply_kernel(self.h, kernel, tile_density))

    def set_apply_policy(self, policy, budget_bytes=0):
        """0 = eager (reference semantics), 1 = deferred scatter-add (collective; see cm_b200.h)"""
        _check(self.L.cm_set_apply_policy(self.h, policy, budget_bytes))

    def apply_policy(self):
        return self.L.cm_get_apply_policy(self.h)

    def materialize(self):
        _check(self.L.cm_materialize(self.h))

    def set_exchange_mode(self, mode):
        _check(self.L.cm_set_exchange_mode(self.h, mode))

    def exchange_mode(self):
        return self.L.cm_get_exchange_mode(self.h)

    def set_profiling(self, on):
        _check(self.L.cm_set_profiling(self.h, int(on)))

    def stream(self):
        p = _vp()
        _check(self.L.cm_stream(self.h, C.byref(p)))
        return p.value

    def stats(self):
        s = Stats()
        _check(self.L.cm_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        _check(self.L.cm_reset_stats(self.h))

    # parity hooks --------------------------------------------------------------------
    def keep_wire(self, on=True):
        _check(self.L.cm_debug_keep_wire(self.h, int(on)))

    def set_staging_limit(self, max_updates):
        _check(self.L.cm_debug_set_staging_limit(self.h, int(max_updates)))

    def wire(self, dst):
        n = _l()
        _check(self.L.cm_debug_wire_count(self.h, dst, C.byref(n)))
        inds = np.zeros(n.value, dtype=np.int64)
        data = np.zeros(n.value, dtype=self.np_dtype)
        if n.value:
            _check(self.L.cm_debug_wire_expand(self.h, dst, _np_ptr(inds), _np_ptr(data)))
        return inds, data


class SliceBatch:
    """Per-update argument arrays for Matrix.update_batch: `count` updates of m_u x n_u (ld = m_u)
    whose payloads / index sets are consecutive in three base buffers (host or device addresses)."""

    def __init__(self, count, m_u, n_u, data_ptr, ix_ptr, jx_ptr, esz):
        k = np.arange(count, dtype=np.int64)
        self.count = count
        self.data = (data_ptr + k * (m_u * n_u * esz)).astype(np.uint64)
        self.ix = (ix_ptr + k * (m_u * 8)).astype(np.uint64)
        self.jx = (jx_ptr + k * (n_u * 8)).astype(np.uint64)
        self.m_u = np.full(count, m_u, dtype=np.int64)
        self.n_u = np.full(count, n_u, dtype=np.int64)
        self.ld = np.full(count, m_u, dtype=np.int64)


class LoopbackWorld:
    """P ranks as P threads of this process on one GPU (cm_rt_init_loopback): the single-GPU
    stand-in for `torchrun --nproc-per-node P`, mirroring how the reference's tests run 4
    ranks on one host over the GASNet smp conduit. ``run(fn)`` executes fn(rank) on every
    rank's thread concurrently (collectives need that); ``on(rank, fn)`` runs on one."""

    def __init__(self, nranks, device=0):
        init_loopback(nranks, device)
        self.n = nranks
        self._cv = threading.Condition()
        self._tasks = [None] * nranks
        self._results = [None] * nranks
        self._pending = 0
        self._stop = False
        self._threads = [threading.Thread(target=self._worker, args=(r,), daemon=True) for r in range(nranks)]
        for t in self._threads:
            t.start()

    def _worker(self, rank):
        bind_rank(rank)
        while True:
            with self._cv:
                while self._tasks[rank] is None and not self._stop:
                    self._cv.wait()
                if self._tasks[rank] is None and self._stop:
                    return
                fn = self._tasks[rank]
                self._tasks[rank] = None
            try:
                res = (True, fn(rank))
            except BaseException as e:  # noqa: BLE001 - re-raised on the caller's thread
                res = (False, e)
            with self._cv:
                self._results[rank] = res
                self._pending -= 1
                self._cv.notify_all()

    def _dispatch(self, ranks, fn):
        with self._cv:
            for r in ranks:
                self._tasks[r] = fn
            self._pending = len(ranks)
            self._cv.notify_all()
            while self._pending:
                self._cv.wait()
            out = [self._results[r] for r in ranks]
        for ok, v in out:
            if not ok:
                raise v
        return [v for _, v in out]

    def run(self, fn):
        return self._dispatch(list(range(self.n)), fn)

    def on(self, rank, fn):
        return self._dispatch([rank], fn)[0]

    def close(self):
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        for t in self._threads:
            t.join()
        finalize()
