r"""ctypes view of the C ABI of ``libpyci_b200.so`` (``include/pyci_b200.h``).

This is the binding a maintainer of the reference would write to reach the B200 path without pybind11
(INTEGRATION.md shows the same calls from C++).  ``bench.py`` and the ``-m gpu`` tests use it to drive the
library with device-resident inputs; ``tests/test_abi.py`` uses :data:`PROTOTYPES` to check that every
function the header declares is exported.  Loading the library needs no GPU; every compute entry point
fails with ``PYCI_ERR_CUDA`` when no device is visible (there is no CPU fallback).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBRARY = os.path.join(_HERE, "libpyci_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "pyci_b200.h")

OK, ERR_VALUE, ERR_TYPE, ERR_RUNTIME, ERR_CUDA, ERR_MEMORY, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
DOCI, FULLCI, GENCI = 0, 1, 2

_vp, _i, _l, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_double
_vpp = ctypes.POINTER(ctypes.c_void_p)


class SolveStats(ctypes.Structure):
    _fields_ = [("matvecs", _l), ("iterations", _l), ("restarts", _l), ("residual", _d), ("seconds", _d),
                ("spmv_seconds", _d)]


# name -> (restype, argtypes); one entry per PYCI_API function of include/pyci_b200.h
PROTOTYPES = {
    "pyci_last_error": (ctypes.c_char_p, []),
    "pyci_abi_version": (_i, []),
    "pyci_device_count": (_i, []),
    "pyci_ctx_create": (_i, [_i, _vp, _vpp]),
    "pyci_ctx_destroy": (None, [_vp]),
    "pyci_ctx_synchronize": (_i, [_vp]),
    "pyci_ctx_release_memory": (_i, [_vp]),
    "pyci_host_alloc": (_i, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]),
    "pyci_host_free": (None, [_vp]),
    "pyci_nccl_unique_id": (_i, [_vp]),
    "pyci_ctx_init_comm": (_i, [_vp, _i, _i, _vp]),
    "pyci_ctx_rank": (_i, [_vp]),
    "pyci_ctx_nranks": (_i, [_vp]),
    "pyci_ctx_launch_count": (_l, [_vp]),
    "pyci_ctx_reset_launch_count": (None, [_vp]),
    "pyci_ham_upload": (_i, [_vp, _l, _d, _vp, _vp, _vp, _vp, _vp, _vpp]),
    "pyci_ham_destroy": (None, [_vp]),
    "pyci_wfn_upload": (_i, [_vp, _i, _l, _l, _l, _l, _vp, _vpp]),
    "pyci_wfn_create_all_dets": (_i, [_vp, _i, _l, _l, _l, _vpp]),
    "pyci_wfn_destroy": (None, [_vp]),
    "pyci_wfn_reindex": (_i, [_vp]),
    "pyci_wfn_index_seconds": (_d, [_vp]),
    "pyci_wfn_index_dets": (_i, [_vp, _l, _vp, _vp]),
    "pyci_wfn_ndet": (_l, [_vp]),
    "pyci_wfn_download_dets": (_i, [_vp, _l, _l, _vp]),
    "pyci_wfn_add_hci": (_i, [_vp, _vp, _vp, _vp, _d, ctypes.POINTER(_l)]),
    "pyci_compute_enpt2": (_i, [_vp, _vp, _vp, _vp, _d, _d, ctypes.POINTER(_d), ctypes.POINTER(_l)]),
    "pyci_wfn_ext_seconds": (_d, [_vp]),
    "pyci_op_build": (_i, [_vp, _vp, _vp, _l, _l, _i, _vpp]),
    "pyci_op_build_shard": (_i, [_vp, _vp, _vp, _l, _l, _i, _i, _i, _vpp]),
    "pyci_op_destroy": (None, [_vp]),
    "pyci_op_update": (_i, [_vp, _vp, _vp]),
    "pyci_op_nrow": (_l, [_vp]),
    "pyci_op_ncol": (_l, [_vp]),
    "pyci_op_row_begin": (_l, [_vp]),
    "pyci_op_row_count": (_l, [_vp]),
    "pyci_op_size": (_l, [_vp]),
    "pyci_op_stored_nnz": (_l, [_vp]),
    "pyci_op_ecore": (_d, [_vp]),
    "pyci_op_build_times": (_i, [_vp, _vp]),
    "pyci_op_fill_seconds": (_d, [_vp]),
    "pyci_op_fill_kernel": (ctypes.c_char_p, [_vp]),
    "pyci_op_count_kernel": (ctypes.c_char_p, [_vp]),
    "pyci_op_export_csr": (_i, [_vp, _vp, _vp, _vp]),
    "pyci_op_export_rows": (_i, [_vp, _l, _vp, _l, _vp, _vp, _vp]),
    "pyci_op_matvec": (_i, [_vp, _vp, _vp]),
    "pyci_op_matvec_dev": (_i, [_vp, _vp, _vp]),
    "pyci_op_time_spmv": (_i, [_vp, _i, _i, _l, _vp]),
    "pyci_op_set_spmv_shape": (_i, [_vp, _i, _i]),
    "pyci_op_set_spmv_block": (_i, [_vp, _i, _i]),
    "pyci_op_get_element": (_i, [_vp, _l, _l, _vp]),
    "pyci_op_solve": (_i, [_vp, _l, _vp, _l, _l, _d, _vp, _vp, ctypes.POINTER(SolveStats)]),
    "pyci_compute_rdms": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "pyci_compute_transition_rdms": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pyci_compute_overlap": (_i, [_vp, _vp, _vp, _vp, _vp, ctypes.POINTER(_d)]),
}

_LIB = None


class PyciError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("pyci_b200 status %d: %s" % (status, message))
        self.status = status


def lib():
    """Load ``libpyci_b200.so`` (once) and attach the prototypes.  Raises OSError if it is not built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIBRARY):
            raise OSError("%s is not built: run `make -C %s`" % (LIBRARY, os.path.join(_HERE, "csrc")))
        L = ctypes.CDLL(LIBRARY, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(status):
    if status != OK:
        raise PyciError(status, lib().pyci_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """pyci_ctx: one device + stream (+ NCCL communicator)."""

    def __init__(self, device=0, stream=0):
        self.handle = ctypes.c_void_p()
        check(lib().pyci_ctx_create(device, ctypes.c_void_p(stream or None), ctypes.byref(self.handle)))

    def close(self):
        if self.handle:
            lib().pyci_ctx_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def synchronize(self):
        check(lib().pyci_ctx_synchronize(self.handle))

    def release_memory(self):
        """hand the pool's unused blocks back to the driver (they are invisible to torch's allocator otherwise)"""
        check(lib().pyci_ctx_release_memory(self.handle))

    def init_comm(self, rank, nranks, unique_id):
        buf = ctypes.create_string_buffer(bytes(unique_id), 128) if nranks > 1 else None
        check(lib().pyci_ctx_init_comm(self.handle, rank, nranks, buf))

    @property
    def launches(self):
        return lib().pyci_ctx_launch_count(self.handle)

    def reset_launches(self):
        lib().pyci_ctx_reset_launch_count(self.handle)


def nccl_unique_id():
    buf = ctypes.create_string_buffer(128)
    check(lib().pyci_nccl_unique_id(buf))
    return buf.raw


class Ham:
    """pyci_ham: integrals resident in HBM."""

    def __init__(self, ctx, nbasis, ecore, one_mo=None, two_mo=None, h=None, v=None, w=None):
        self.keep = [_f64(a) for a in (one_mo, two_mo, h, v, w)]
        self.handle = ctypes.c_void_p()
        check(lib().pyci_ham_upload(ctx.handle, nbasis, float(ecore), *[_ptr(a) for a in self.keep],
                                    ctypes.byref(self.handle)))

    def close(self):
        if self.handle:
            lib().pyci_ham_destroy(self.handle)
            self.handle = ctypes.c_void_p()


class Wfn:
    """pyci_wfn: determinant array + hash index resident in HBM."""

    def __init__(self, ctx, kind, nbasis, nocc_up, nocc_dn, dets=None):
        """dets=None: the complete space, generated on the device (Wfn::add_all_dets)."""
        self.ctx = ctx
        self.handle = ctypes.c_void_p()
        if dets is None:
            check(lib().pyci_wfn_create_all_dets(ctx.handle, kind, nbasis, nocc_up, nocc_dn, ctypes.byref(self.handle)))
            self.ndet = int(lib().pyci_wfn_ndet(self.handle))
            self.det_shape = (2, 1) if kind == FULLCI else (1,)
            return
        dets = np.ascontiguousarray(dets, dtype=np.uint64)
        self.ndet = int(dets.shape[0])
        self.det_shape = tuple(dets.shape[1:])
        check(lib().pyci_wfn_upload(ctx.handle, kind, nbasis, nocc_up, nocc_dn, self.ndet, _ptr(dets),
                                    ctypes.byref(self.handle)))

    def reindex(self):
        check(lib().pyci_wfn_reindex(self.handle))
        return lib().pyci_wfn_index_seconds(self.handle)

    def index_dets(self, dets):
        dets = np.ascontiguousarray(dets, dtype=np.uint64)
        out = np.empty(dets.shape[0], dtype=np.int64)
        check(lib().pyci_wfn_index_dets(self.handle, dets.shape[0], _ptr(dets), _ptr(out)))
        return out

    def download_dets(self, start=0, n=None):
        n = self.ndet - start if n is None else n
        out = np.empty((n,) + self.det_shape, dtype=np.uint64)
        check(lib().pyci_wfn_download_dets(self.handle, start, n, _ptr(out)))
        return out

    def add_hci(self, ham, coeffs, eps=1.0e-5):
        """pyci.add_hci on the device wave function; returns the appended determinants."""
        c = _f64(coeffs)
        nnew = _l(0)
        old = self.ndet
        check(lib().pyci_wfn_add_hci(self.ctx.handle, ham.handle, self.handle, _ptr(c), eps, ctypes.byref(nnew)))
        self.ndet = int(lib().pyci_wfn_ndet(self.handle))
        return self.download_dets(old, nnew.value)

    def compute_enpt2(self, ham, coeffs, energy, eps=1.0e-5):
        """pyci.compute_enpt2 (FullCI / GenCI kinds); returns (energy + correction, external determinants)."""
        c = _f64(coeffs)
        out, nt = _d(0.0), _l(0)
        check(lib().pyci_compute_enpt2(self.ctx.handle, ham.handle, self.handle, _ptr(c), energy, eps,
                                       ctypes.byref(out), ctypes.byref(nt)))
        return out.value, nt.value

    def ext_seconds(self):
        return lib().pyci_wfn_ext_seconds(self.handle)

    def close(self):
        if self.handle:
            lib().pyci_wfn_destroy(self.handle)
            self.handle = ctypes.c_void_p()


class Op:
    """pyci_op: CSR row shard resident in HBM."""

    def __init__(self, ctx, ham, wfn, nrow=-1, ncol=-1, symmetric=True, shard=None):
        """shard = (rank, nranks): that row block of the row-sharded operator, built without a communicator."""
        self.handle = ctypes.c_void_p()
        if shard is None:
            check(lib().pyci_op_build(ctx.handle, ham.handle, wfn.handle, nrow, ncol, int(bool(symmetric)),
                                      ctypes.byref(self.handle)))
        else:
            check(lib().pyci_op_build_shard(ctx.handle, ham.handle, wfn.handle, nrow, ncol, int(bool(symmetric)),
                                            int(shard[0]), int(shard[1]), ctypes.byref(self.handle)))
        self._refresh()

    def _refresh(self):
        L = lib()
        self.nrow, self.ncol = L.pyci_op_nrow(self.handle), L.pyci_op_ncol(self.handle)
        self.row_begin, self.row_count = L.pyci_op_row_begin(self.handle), L.pyci_op_row_count(self.handle)
        self.stored_nnz = L.pyci_op_stored_nnz(self.handle)

    @property
    def size(self):
        """SparseOp::size of this rank's rows in the reference's storage (summed on the device at first use)"""
        return lib().pyci_op_size(self.handle)

    def update(self, ham, wfn):
        """SparseOp::update: grow to all determinants now in wfn (incremental; square symmetric, one rank)."""
        check(lib().pyci_op_update(self.handle, ham.handle, wfn.handle))
        self._refresh()

    def close(self):
        if self.handle:
            lib().pyci_op_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def build_times(self):
        t = np.zeros(4)
        check(lib().pyci_op_build_times(self.handle, _ptr(t)))
        return dict(index=t[0], count_scan=t[1], fill_sort=t[2], total=t[3])

    def fill_kernel(self):
        return lib().pyci_op_fill_kernel(self.handle).decode()

    def count_kernel(self):
        return lib().pyci_op_count_kernel(self.handle).decode()

    def fill_seconds(self):
        return lib().pyci_op_fill_seconds(self.handle)

    def export_csr(self):
        indptr = np.empty(self.row_count + 1, dtype=np.int64)
        check(lib().pyci_op_export_csr(self.handle, _ptr(indptr), None, None))
        indices = np.empty(indptr[-1], dtype=np.int64)
        data = np.empty(indptr[-1], dtype=np.float64)
        check(lib().pyci_op_export_csr(self.handle, _ptr(indptr), _ptr(indices), _ptr(data)))
        return indptr, indices, data

    def export_rows(self, rows):
        """(indptr, indices, data) of the listed global rows in the reference's layout (pyci_op_export_rows)."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        indptr = np.zeros(len(rows) + 1, dtype=np.int64)
        check(lib().pyci_op_export_rows(self.handle, len(rows), _ptr(rows), 0, _ptr(indptr), None, None))
        total = int(indptr[-1])
        indices = np.empty(total, dtype=np.int64)
        data = np.empty(total, dtype=np.float64)
        check(lib().pyci_op_export_rows(self.handle, len(rows), _ptr(rows), total, _ptr(indptr), _ptr(indices), _ptr(data)))
        return indptr, indices, data

    def matvec(self, x, out=None):
        x = _f64(x)
        y = np.empty(self.nrow) if out is None else out
        check(lib().pyci_op_matvec(self.handle, _ptr(x), _ptr(y)))
        return y

    def matvec_dev(self, x_dev_ptr, y_dev_ptr):
        check(lib().pyci_op_matvec_dev(self.handle, ctypes.c_void_p(x_dev_ptr), ctypes.c_void_p(y_dev_ptr)))

    def time_spmv(self, warmup=3, reps=10, flush_bytes=0):
        ms = np.zeros(reps)
        check(lib().pyci_op_time_spmv(self.handle, warmup, reps, flush_bytes, _ptr(ms)))
        return ms

    def set_spmv_shape(self, threads_per_row=0, ctas_per_sm=4):
        check(lib().pyci_op_set_spmv_shape(self.handle, threads_per_row, ctas_per_sm))

    def set_spmv_block(self, block_threads=256, depth=2):
        check(lib().pyci_op_set_spmv_block(self.handle, block_threads, depth))

    def get_element(self, i, j):
        v = ctypes.c_double(0.0)
        check(lib().pyci_op_get_element(self.handle, i, j, ctypes.byref(v)))
        return v.value

    def solve(self, n=1, c0=None, ncv=-1, maxiter=-1, tol=1e-12):
        evals = np.empty(n)
        evecs = np.empty((n, self.nrow))
        stats = SolveStats()
        c0 = _f64(c0)
        check(lib().pyci_op_solve(self.handle, n, _ptr(c0), ncv, maxiter, tol, _ptr(evals), _ptr(evecs),
                                  ctypes.byref(stats)))
        return evals, evecs, {f: getattr(stats, f) for f, _ in SolveStats._fields_}


def compute_rdms(ctx, wfn, kind, nbasis, coeffs):
    n = nbasis
    if kind == DOCI:
        r1, r2 = np.empty((n, n)), np.empty((n, n))
    elif kind == FULLCI:
        r1, r2 = np.empty((2, n, n)), np.empty((3, n, n, n, n))
    else:
        r1, r2 = np.empty((n, n)), np.empty((n, n, n, n))
    c = _f64(coeffs)
    check(lib().pyci_compute_rdms(ctx.handle, wfn.handle, _ptr(c), _ptr(r1), _ptr(r2)))
    return r1, r2


def compute_transition_rdms(ctx, wfn1, wfn2, kind, nbasis, coeffs1, coeffs2):
    n = nbasis
    if kind == DOCI:
        r1, r2 = np.empty((n, n)), np.empty((n, n))
    elif kind == FULLCI:
        r1, r2 = np.empty((2, n, n)), np.empty((3, n, n, n, n))
    else:
        r1, r2 = np.empty((n, n)), np.empty((n, n, n, n))
    c1, c2 = _f64(coeffs1), _f64(coeffs2)
    check(lib().pyci_compute_transition_rdms(ctx.handle, wfn1.handle, wfn2.handle, _ptr(c1), _ptr(c2), _ptr(r1), _ptr(r2)))
    return r1, r2


def compute_overlap(ctx, wfn1, wfn2, coeffs1, coeffs2):
    c1, c2 = _f64(coeffs1), _f64(coeffs2)
    out = _d(0.0)
    check(lib().pyci_compute_overlap(ctx.handle, wfn1.handle, wfn2.handle, _ptr(c1), _ptr(c2), ctypes.byref(out)))
    return out.value
