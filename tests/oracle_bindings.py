"""ctypes bindings for the CHECKERS under oracle/ (test infrastructure only).

* ``Oracle``  -> oracle/libcm_oracle.so  (plain-C restatement, oracle/cm_oracle.c)
* ``RefShim`` -> oracle/_ref/libcm_ref.so (the unmodified reference header over the in-repo
  UPC++ shim, oracle/ref_driver.cpp); only present where it was built from /root/reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libcm_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libcm_ref.so")

F64, F32, I32 = 0, 1, 2
NP_DTYPE = {F64: np.float64, F32: np.float32, I32: np.int32}
DTYPE_CODE = {np.dtype(np.float64): F64, np.dtype(np.float32): F32, np.dtype(np.int32): I32}

_vp, _l, _i = C.c_void_p, C.c_long, C.c_int
_lp = C.POINTER(C.c_long)


def _ptr(a):
    return a.ctypes.data_as(_vp)


def _lptr(a):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(_lp)


def _as_long(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


class Oracle:
    """All P ranks of one ConvergentMatrix simulated by oracle/cm_oracle.c."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(ORACLE_SO)
            L.cmo_roc.restype = _l
            L.cmo_roc.argtypes = [_l] * 4
            L.cmo_owner.argtypes = [_l] * 7 + [C.POINTER(_i), _lp]
            L.cmo_create.restype = _vp
            L.cmo_create.argtypes = [_i] + [_l] * 7
            L.cmo_destroy.argtypes = [_vp]
            L.cmo_nranks.argtypes = [_vp]
            L.cmo_m_local.restype = _l
            L.cmo_m_local.argtypes = [_vp, _i]
            L.cmo_n_local.restype = _l
            L.cmo_n_local.argtypes = [_vp, _i]
            L.cmo_local_data.restype = _vp
            L.cmo_local_data.argtypes = [_vp, _i]
            L.cmo_set_flush_threshold.argtypes = [_vp, _l]
            L.cmo_update_slice.argtypes = [_vp, _i, _vp, _l, _l, _l, _i, _lp, _lp]
            L.cmo_update_sym.argtypes = [_vp, _i, _vp, _l, _l, _i, _lp]
            L.cmo_update_elem.argtypes = [_vp, _i, _vp, _l, _l]
            L.cmo_fill_lower.argtypes = [_vp]
            L.cmo_commit.argtypes = [_vp]
            L.cmo_reset.argtypes = [_vp]
            L.cmo_get.argtypes = [_vp, _l, _l, _vp]
            L.cmo_wire_enable.argtypes = [_vp, _i]
            L.cmo_wire_count.restype = _l
            L.cmo_wire_count.argtypes = [_vp, _i, _i]
            L.cmo_wire_copy.argtypes = [_vp, _i, _i, _lp, _vp]
            L.cmo_wire_clear.argtypes = [_vp]
            L.cmo_elements_applied.restype = _l
            L.cmo_elements_applied.argtypes = [_vp]
            L.cmo_messages_sent.restype = _l
            L.cmo_messages_sent.argtypes = [_vp]
            cls._lib = L
        return cls._lib

    def __init__(self, dtype, m, n, nprow, npcol, mb, nb, lld, record_wire=False):
        self.L = self.lib()
        self.dtype = dtype
        self.np_dtype = NP_DTYPE[dtype]
        self.geom = (m, n, nprow, npcol, mb, nb, lld)
        self.lld = lld
        self.P = nprow * npcol
        self.h = self.L.cmo_create(dtype, m, n, nprow, npcol, mb, nb, lld)
        if not self.h:
            raise ValueError(f"invalid distribution (reference asserts :417-428): {self.geom}")
        if record_wire:
            self.L.cmo_wire_enable(self.h, 1)

    def close(self):
        if self.h:
            self.L.cmo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def roc(cls, m, np_, ip, mb):
        return cls.lib().cmo_roc(m, np_, ip, mb)

    @classmethod
    def owner(cls, i, j, nprow, npcol, mb, nb, lld):
        r, ij = _i(), _l()
        cls.lib().cmo_owner(i, j, nprow, npcol, mb, nb, lld, C.byref(r), C.byref(ij))
        return r.value, ij.value

    def m_local(self, r):
        return self.L.cmo_m_local(self.h, r)

    def n_local(self, r):
        return self.L.cmo_n_local(self.h, r)

    def set_flush_threshold(self, t):
        self.L.cmo_set_flush_threshold(self.h, t)

    def update_slice(self, src, data, m_u, n_u, ld, trans, ix, jx):
        data = np.ascontiguousarray(data, dtype=self.np_dtype)
        ix, jx = _as_long(ix), _as_long(jx)
        self.L.cmo_update_slice(self.h, src, _ptr(data), m_u, n_u, ld, int(trans), _lptr(ix), _lptr(jx))

    def update_sym(self, src, data, n_u, ld, trans, ix):
        data = np.ascontiguousarray(data, dtype=self.np_dtype)
        ix = _as_long(ix)
        self.L.cmo_update_sym(self.h, src, _ptr(data), n_u, ld, int(trans), _lptr(ix))

    def update_elem(self, src, val, i, j):
        v = np.array([val], dtype=self.np_dtype)
        self.L.cmo_update_elem(self.h, src, _ptr(v), i, j)

    def fill_lower(self):
        self.L.cmo_fill_lower(self.h)

    def commit(self):
        self.L.cmo_commit(self.h)

    def reset(self):
        self.L.cmo_reset(self.h)

    def get(self, i, j):
        v = np.zeros(1, dtype=self.np_dtype)
        self.L.cmo_get(self.h, i, j, _ptr(v))
        return v[0]

    def local(self, r):
        """copy of rank r's lld x n_local block as a (n_local, lld) C-array == column-major."""
        n = self.lld * self.n_local(r)
        p = self.L.cmo_local_data(self.h, r)
        buf = (C.c_char * (n * np.dtype(self.np_dtype).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=self.np_dtype).copy()

    def wire(self, src, dst):
        n = self.L.cmo_wire_count(self.h, src, dst)
        inds = np.zeros(n, dtype=np.int64)
        data = np.zeros(n, dtype=self.np_dtype)
        self.L.cmo_wire_copy(self.h, src, dst, _lptr(inds), _ptr(data))
        return inds, data

    def wire_clear(self):
        self.L.cmo_wire_clear(self.h)

    def elements_applied(self):
        return self.L.cmo_elements_applied(self.h)

    def messages_sent(self):
        return self.L.cmo_messages_sent(self.h)


def ref_available():
    return os.path.exists(REF_SO)


class RefShim:
    """The reference's own ConvergentMatrix (P threads-as-ranks) via oracle/ref_driver.cpp."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(REF_SO)
            L.cmref_have_config.argtypes = [_i] + [_l] * 5
            L.cmref_open.restype = _vp
            L.cmref_open.argtypes = [_i] + [_l] * 7 + [_i]
            L.cmref_close.argtypes = [_vp]
            L.cmref_nranks.argtypes = [_vp]
            L.cmref_update_slice.argtypes = [_vp, _i, _vp, _l, _l, _l, _i, _lp, _lp]
            L.cmref_update_sym.argtypes = [_vp, _i, _vp, _l, _l, _i, _lp]
            L.cmref_update_elem.argtypes = [_vp, _i, _vp, _l, _l]
            L.cmref_update_elems.argtypes = [_vp, _i, _vp, _lp, _lp, _l]
            L.cmref_fill_lower.argtypes = [_vp]
            L.cmref_commit.argtypes = [_vp]
            L.cmref_reset.argtypes = [_vp]
            L.cmref_local_data.restype = _vp
            L.cmref_local_data.argtypes = [_vp, _i]
            for f in ("cmref_m_local", "cmref_n_local", "cmref_pgrid_row", "cmref_pgrid_col"):
                getattr(L, f).restype = _l
                getattr(L, f).argtypes = [_vp, _i]
            L.cmref_get.argtypes = [_vp, _i, _l, _l, _vp]
            L.cmref_set_flush_threshold.argtypes = [_vp, _i]
            L.cmref_wire_count.restype = _l
            L.cmref_wire_count.argtypes = [_vp, _i, _i]
            L.cmref_wire_copy.argtypes = [_vp, _i, _i, _lp, _vp]
            L.cmref_wire_clear.argtypes = [_vp]
            L.cmref_bench_slices.argtypes = [_vp, C.POINTER(_vp), C.POINTER(_lp), C.POINTER(_lp),
                                             _l, _l, _l, _l, _i, C.POINTER(C.c_double), _l]
            cls._lib = L
        return cls._lib

    @classmethod
    def have_config(cls, dtype, nprow, npcol, mb, nb, lld):
        return bool(cls.lib().cmref_have_config(dtype, nprow, npcol, mb, nb, lld))

    def __init__(self, dtype, m, n, nprow, npcol, mb, nb, lld, record_wire=False):
        self.L = self.lib()
        self.dtype = dtype
        self.np_dtype = NP_DTYPE[dtype]
        self.lld = lld
        self.P = nprow * npcol
        self.h = self.L.cmref_open(dtype, nprow, npcol, mb, nb, lld, m, n, int(record_wire))
        if not self.h:
            raise RuntimeError("reference config not instantiated in oracle/ref_configs.def "
                               f"(or a session is already open): {(dtype, nprow, npcol, mb, nb, lld)}")

    def close(self):
        if self.h:
            self.L.cmref_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def m_local(self, r):
        return self.L.cmref_m_local(self.h, r)

    def n_local(self, r):
        return self.L.cmref_n_local(self.h, r)

    def pgrid_row(self, r):
        return self.L.cmref_pgrid_row(self.h, r)

    def pgrid_col(self, r):
        return self.L.cmref_pgrid_col(self.h, r)

    def set_flush_threshold(self, t):
        self.L.cmref_set_flush_threshold(self.h, t)

    def update_slice(self, src, data, m_u, n_u, ld, trans, ix, jx):
        data = np.ascontiguousarray(data, dtype=self.np_dtype)
        ix, jx = _as_long(ix), _as_long(jx)
        self.L.cmref_update_slice(self.h, src, _ptr(data), m_u, n_u, ld, int(trans), _lptr(ix), _lptr(jx))

    def update_sym(self, src, data, n_u, ld, trans, ix):
        data = np.ascontiguousarray(data, dtype=self.np_dtype)
        ix = _as_long(ix)
        self.L.cmref_update_sym(self.h, src, _ptr(data), n_u, ld, int(trans), _lptr(ix))

    def update_elem(self, src, val, i, j):
        v = np.array([val], dtype=self.np_dtype)
        self.L.cmref_update_elem(self.h, src, _ptr(v), i, j)

    def update_elems(self, src, vals, is_, js):
        vals = np.ascontiguousarray(vals, dtype=self.np_dtype)
        is_, js = _as_long(is_), _as_long(js)
        self.L.cmref_update_elems(self.h, src, _ptr(vals), _lptr(is_), _lptr(js), len(vals))

    def fill_lower(self):
        self.L.cmref_fill_lower(self.h)

    def commit(self):
        self.L.cmref_commit(self.h)

    def reset(self):
        self.L.cmref_reset(self.h)

    def get(self, rank, i, j):
        v = np.zeros(1, dtype=self.np_dtype)
        self.L.cmref_get(self.h, rank, i, j, _ptr(v))
        return v[0]

    def local(self, r):
        n = self.lld * self.n_local(r)
        p = self.L.cmref_local_data(self.h, r)
        buf = (C.c_char * (n * np.dtype(self.np_dtype).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=self.np_dtype).copy()

    def wire(self, src, dst):
        n = self.L.cmref_wire_count(self.h, src, dst)
        inds = np.zeros(n, dtype=np.int64)
        data = np.zeros(n, dtype=self.np_dtype)
        self.L.cmref_wire_copy(self.h, src, dst, _lptr(inds), _ptr(data))
        return inds, data

    def wire_clear(self):
        self.L.cmref_wire_clear(self.h)

    def bench_slices(self, data, ix, jx, m_u, n_u, commit_every=0, pin=True, data_period=0):
        """data/ix/jx: per-rank lists of contiguous arrays ([nupd, n_u, m_u] payloads stored
        column-major per update, [nupd, m_u], [nupd, n_u]). Returns (total_s, update_s, commit_s).
        data_period > 0: data holds only that many payloads per rank, cycled over the nupd updates."""
        P = self.P
        nupd = ix[0].shape[0]
        dptr = (_vp * P)(*[_ptr(d) for d in data])
        iptr = (_lp * P)(*[_lptr(a) for a in ix])
        jptr = (_lp * P)(*[_lptr(a) for a in jx])
        out = (C.c_double * 3)()
        self.L.cmref_bench_slices(self.h, dptr, iptr, jptr, nupd, m_u, n_u, commit_every, int(pin), out, data_period)
        return out[0], out[1], out[2]
