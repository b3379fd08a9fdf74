"""TEST INFRASTRUCTURE ONLY: ctypes front-end of oracle/liboracle.so (plain-C restatement of the
reference hot path, see pyci_oracle.c) plus numpy helpers that restate the reference's input
conventions.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package (pyci_b200) never does.

Reference citations are relative to /root/reference.
"""
import ctypes
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DOCI, FULLCI, GENCI = 0, 1, 2

_c_long_p = ctypes.POINTER(ctypes.c_long)
_c_double_p = ctypes.POINTER(ctypes.c_double)


def build():
    """Compile liboracle.so (gcc) if it is missing or older than its source."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "pyci_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.oracle_sparse_op.restype = ctypes.c_long
        L.oracle_sparse_op.argtypes = [
            ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_long,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_long, ctypes.c_long, ctypes.c_int,
            ctypes.POINTER(_c_long_p), ctypes.POINTER(_c_long_p), ctypes.POINTER(_c_double_p)]
        L.oracle_sparse_op_rows.restype = ctypes.c_long
        L.oracle_sparse_op_rows.argtypes = [
            ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_long,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_long, ctypes.c_void_p, ctypes.c_long, ctypes.c_int,
            ctypes.POINTER(_c_long_p), ctypes.POINTER(_c_long_p), ctypes.POINTER(_c_double_p)]
        L.oracle_free.argtypes = [ctypes.c_void_p]
        L.oracle_matvec.argtypes = [ctypes.c_long] + [ctypes.c_void_p] * 3 + [ctypes.c_int] + \
            [ctypes.c_void_p] * 2
        L.oracle_rdms_doci.argtypes = [ctypes.c_long] * 3 + [ctypes.c_void_p] * 4
        L.oracle_rdms_fullci.argtypes = [ctypes.c_long] * 4 + [ctypes.c_void_p] * 4
        L.oracle_rdms_genci.argtypes = [ctypes.c_long] * 3 + [ctypes.c_void_p] * 4
        L.oracle_all_dets_onespin.argtypes = [ctypes.c_long, ctypes.c_long, ctypes.c_void_p]
        L.oracle_all_dets_twospin.argtypes = [ctypes.c_long] * 3 + [ctypes.c_void_p]
        L.oracle_add_hci.restype = ctypes.c_long
        L.oracle_add_hci.argtypes = [ctypes.c_int] + [ctypes.c_long] * 4 + [ctypes.c_void_p] * 4 + \
            [ctypes.c_double, ctypes.POINTER(ctypes.POINTER(ctypes.c_uint64))]
        L.oracle_enpt2.restype = ctypes.c_int
        L.oracle_enpt2.argtypes = [ctypes.c_int] + [ctypes.c_long] * 4 + [ctypes.c_void_p] * 4 + \
            [ctypes.c_double] * 3 + [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_long)]
        L.oracle_trdms_doci.argtypes = [ctypes.c_long] * 3 + [ctypes.c_void_p, ctypes.c_long] + [ctypes.c_void_p] * 5
        L.oracle_trdms_fullci.argtypes = [ctypes.c_long] * 4 + [ctypes.c_void_p, ctypes.c_long] + [ctypes.c_void_p] * 5
        L.oracle_trdms_genci.argtypes = [ctypes.c_long] * 3 + [ctypes.c_void_p, ctypes.c_long] + [ctypes.c_void_p] * 5
        L.oracle_overlap.argtypes = [ctypes.c_long, ctypes.c_long, ctypes.c_void_p, ctypes.c_long] + \
            [ctypes.c_void_p] * 3 + [ctypes.POINTER(ctypes.c_double)]
        L.oracle_binomial.restype = ctypes.c_long
        L.oracle_binomial.argtypes = [ctypes.c_long, ctypes.c_long]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def nword(nbasis):
    return (nbasis + 63) // 64


# ---------------------------------------------------------------------------------------------
# inputs


def read_fcidump(path):
    """Restates SQuantOp(filename), squantop.cpp:79-161: returns (ecore, one_mo[n,n], two_mo[n,n,n,n])
    with two_mo in physicist order two_mo[i,k,j,l] = (ij|kl), 8-fold symmetric fill."""
    with open(path) as f:
        text = f.read()
    m = re.search(r"&END|^\s*/\s*$", text, flags=re.M)
    header, body = text[:m.start()], text[m.end():]
    n = int(re.search(r"NORB\s*=\s*(\d+)", header).group(1))
    one = np.zeros((n, n))
    two = np.zeros((n, n, n, n))
    ecore = 0.0
    vals = body.split()
    for q in range(0, len(vals) - 4, 5):
        x = float(vals[q].replace("D", "E").replace("d", "e"))
        i, j, k, l = (int(t) for t in vals[q + 1:q + 5])
        if i and j and k and l:
            i, j, k, l = i - 1, j - 1, k - 1, l - 1
            two[i, k, j, l] = x
            two[k, i, l, j] = x
            two[j, k, i, l] = x
            two[i, l, j, k] = x
            two[j, l, i, k] = x
            two[l, j, k, i] = x
            two[k, j, l, i] = x
            two[l, i, k, j] = x
        elif i and j:
            one[i - 1, j - 1] = x
            one[j - 1, i - 1] = x
        else:
            ecore = x
    return ecore, one, two


def senzero_integrals(one_mo, two_mo):
    """h, v, w exactly as squantop.cpp:152-160 / :171-182 derive them."""
    n = one_mo.shape[0]
    h = np.ascontiguousarray(np.diag(one_mo).copy())
    p = np.arange(n)
    v = np.ascontiguousarray(two_mo[p[:, None], p[:, None], p[None, :], p[None, :]])
    w = np.ascontiguousarray(two_mo[p[:, None], p[None, :], p[:, None], p[None, :]] * 2
                             - two_mo[p[:, None], p[None, :], p[None, :], p[:, None]])
    return h, v, w


def synthetic_integrals(n, seed=1234):
    """The synthetic Hamiltonian of SURVEY.md section 8(d) (configs 3-5): random symmetric one_mo with a
    spread diagonal, 0.1*N(0,1) two-electron integrals symmetrised to the 8-fold symmetry of real
    orbitals, returned in physicist order.  Every element is non-zero."""
    rng = np.random.default_rng(seed)
    h = rng.standard_normal((n, n))
    h = (h + h.T) / 2
    h[np.arange(n), np.arange(n)] += np.arange(n)
    g = 0.1 * rng.standard_normal((n, n, n, n))
    g = g + g.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    two_mo = np.ascontiguousarray(g.transpose(0, 2, 1, 3))
    return 0.0, np.ascontiguousarray(h), two_mo


def spin_orbital_integrals(one_mo, two_mo):
    """Spatial (n) -> spin-orbital (2n, alpha block first) integrals such that
    GenCI(2n, N) on the result == FullCI(n, na, nb) on the input."""
    n = one_mo.shape[0]
    h = np.zeros((2 * n, 2 * n))
    h[:n, :n] = one_mo
    h[n:, n:] = one_mo
    g = np.zeros((2 * n,) * 4)
    for s in (0, n):
        for t in (0, n):
            g[s:s + n, t:t + n, s:s + n, t:t + n] = two_mo
    return h, g


def all_dets(kind, nbasis, nocc_up, nocc_dn=0):
    """add_all_dets order: onespinwfn.cpp:173-217 (colex) / twospinwfn.cpp:181-245."""
    L = lib()
    nw = nword(nbasis)
    if kind == FULLCI:
        nd = L.oracle_binomial(nbasis, nocc_up) * L.oracle_binomial(nbasis, nocc_dn)
        dets = np.zeros((nd, 2, nw), dtype=np.uint64)
        L.oracle_all_dets_twospin(nbasis, nocc_up, nocc_dn, _p(dets))
    else:
        nd = L.oracle_binomial(nbasis, nocc_up)
        dets = np.zeros((nd, nw), dtype=np.uint64)
        L.oracle_all_dets_onespin(nbasis, nocc_up, _p(dets))
    return dets


# ---------------------------------------------------------------------------------------------
# the path


def sparse_op(kind, nbasis, nocc_up, nocc_dn, dets, ints, nrow=-1, ncol=-1, symmetric=True, rows=None):
    """Restates pyci.sparse_op(ham, wfn, nrow, ncol, symmetric) (sparseop.cpp:49-71,186-502).
    ints = (one_mo, two_mo) for FullCI/GenCI, (h, v, w) for DOCI.
    rows: optional list of row indices -- row k of the result is row rows[k] of the operator (the reference's
    row loop, sparseop.cpp:196-199, treats rows independently); nrow is then ignored.
    Returns (indptr int64[nrow+1], indices int64[nnz], data float64[nnz])."""
    L = lib()
    dets = np.ascontiguousarray(dets, dtype=np.uint64)
    ndet = dets.shape[0]
    ncol = ndet if ncol < 0 else ncol
    if rows is not None:
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        nrow = len(rows)
    else:
        nrow = ndet if nrow < 0 else nrow
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in ints] + [None]
    ip, ix, dv = _c_long_p(), _c_long_p(), _c_double_p()
    nnz = L.oracle_sparse_op_rows(kind, nbasis, nocc_up, nocc_dn, ndet, _p(dets), _p(arrs[0]), _p(arrs[1]),
                                  _p(arrs[2]), nrow, _p(rows), ncol, int(bool(symmetric)),
                                  ctypes.byref(ip), ctypes.byref(ix), ctypes.byref(dv))
    if nnz < 0:
        raise MemoryError("oracle_sparse_op failed")
    indptr = np.ctypeslib.as_array(ip, shape=(nrow + 1,)).astype(np.int64, copy=True)
    if nnz > 0:
        indices = np.ctypeslib.as_array(ix, shape=(nnz,)).astype(np.int64, copy=True)
        data = np.ctypeslib.as_array(dv, shape=(nnz,)).astype(np.float64, copy=True)
    else:
        indices, data = np.zeros(0, dtype=np.int64), np.zeros(0)
    L.oracle_free(ip)
    L.oracle_free(ix)
    L.oracle_free(dv)
    return indptr, indices, data


def sparse_op_updated(kind, nbasis, nocc_up, nocc_dn, dets, ints, nrow0, sizes, symmetric=True):
    """Restates SparseOp::update (sparseop.cpp:175-201) applied after every growth of the wave function: the operator is
    built with nrow0 rows and sizes[0] columns from the first sizes[0] determinants; update k appends the rows
    [rows so far, sizes[k]) with ncol = sizes[k] and leaves the rows it already has untouched -- so a non-symmetric
    operator keeps, in its old rows, only the columns those rows were built with.  dets is the final list (the wave
    function only appends).  Returns (indptr, indices, data) of the operator after the last update."""
    blocks, done = [(0, nrow0, sizes[0])], nrow0
    for n in sizes[1:]:
        blocks.append((done, n, n))
        done = max(done, n)
    ips, ixs, dvs, off = [np.zeros(1, dtype=np.int64)], [], [], 0
    for lo, hi, ncol in blocks:
        if hi <= lo:
            continue
        ip, ix, dv = sparse_op(kind, nbasis, nocc_up, nocc_dn, dets, ints, ncol=ncol, symmetric=symmetric,
                               rows=np.arange(lo, hi, dtype=np.int64))
        ips.append(ip[1:] + off)
        ixs.append(ix)
        dvs.append(dv)
        off += int(ip[-1])
    return (np.concatenate(ips), np.concatenate(ixs) if ixs else np.zeros(0, dtype=np.int64),
            np.concatenate(dvs) if dvs else np.zeros(0))


def matvec(indptr, indices, data, x, symmetric):
    """Restates SparseOp::perform_op / perform_op_symm (sparseop.cpp:96-112)."""
    nrow = len(indptr) - 1
    y = np.zeros(nrow)
    x = np.ascontiguousarray(x, dtype=np.float64)
    lib().oracle_matvec(nrow, _p(indptr), _p(indices), _p(data), int(bool(symmetric)), _p(x), _p(y))
    return y


def compute_rdms(kind, nbasis, nocc_up, nocc_dn, dets, coeffs):
    """Restates pyci.compute_rdms (rdm.cpp:20-65, 269-530; GenCI: intended semantics, see pyci_oracle.c)."""
    L = lib()
    dets = np.ascontiguousarray(dets, dtype=np.uint64)
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    n = nbasis
    if kind == DOCI:
        d0, d2 = np.zeros((n, n)), np.zeros((n, n))
        rc = L.oracle_rdms_doci(n, nocc_up, dets.shape[0], _p(dets), _p(c), _p(d0), _p(d2))
        out = (d0, d2)
    elif kind == FULLCI:
        r1, r2 = np.zeros((2, n, n)), np.zeros((3, n, n, n, n))
        rc = L.oracle_rdms_fullci(n, nocc_up, nocc_dn, dets.shape[0], _p(dets), _p(c), _p(r1), _p(r2))
        out = (r1, r2)
    else:
        r1, r2 = np.zeros((n, n)), np.zeros((n, n, n, n))
        rc = L.oracle_rdms_genci(n, nocc_up, dets.shape[0], _p(dets), _p(c), _p(r1), _p(r2))
        out = (r1, r2)
    if rc != 0:
        raise MemoryError("oracle rdms failed")
    return out


def add_hci(kind, nbasis, nocc_up, nocc_dn, dets, ints, coeffs, eps=1.0e-5):
    """Restates pyci.add_hci(ham, wfn, coeffs, eps) (hci.cpp:22-279): returns the determinants the reference
    would append, in its order.  ints = (one_mo, two_mo) for FullCI/GenCI, (h, v, w) for DOCI."""
    L = lib()
    dets = np.ascontiguousarray(dets, dtype=np.uint64)
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in ints]
    out = ctypes.POINTER(ctypes.c_uint64)()
    n = L.oracle_add_hci(kind, nbasis, nocc_up, nocc_dn, dets.shape[0], _p(dets), _p(a[0]), _p(a[1]), _p(c),
                         float(eps), ctypes.byref(out))
    if n < 0:
        raise MemoryError("oracle_add_hci failed")
    shape = (n,) + tuple(dets.shape[1:])
    words = int(np.prod(shape))
    new = np.ctypeslib.as_array(out, shape=(max(words, 1),))[:words].astype(np.uint64, copy=True).reshape(shape)
    L.oracle_free(out)
    return new


def compute_enpt2(kind, nbasis, nocc_up, nocc_dn, dets, ints, coeffs, energy, ecore, eps=1.0e-6):
    """Restates pyci.compute_enpt2(ham, wfn, coeffs, energy, eps) (enpt2.cpp:21-400); ints = (one_mo, two_mo)
    also for DOCI, which the reference evaluates on the FullCI image of the wave function (enpt2.cpp:376-380).
    Returns (energy + correction, number of external determinants)."""
    L = lib()
    dets = np.ascontiguousarray(dets, dtype=np.uint64)
    if kind == DOCI:
        d = dets.reshape(dets.shape[0], 1, -1)
        dets = np.ascontiguousarray(np.concatenate([d, d], axis=1))
        kind, nocc_dn = FULLCI, nocc_up
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in ints]
    out, nt = ctypes.c_double(0.0), ctypes.c_long(0)
    rc = L.oracle_enpt2(kind, nbasis, nocc_up, nocc_dn, dets.shape[0], _p(dets), _p(a[0]), _p(a[1]), _p(c),
                        float(energy), float(ecore), float(eps), ctypes.byref(out), ctypes.byref(nt))
    if rc != 0:
        raise MemoryError("oracle_enpt2 failed")
    return out.value, nt.value


def compute_transition_rdms(kind, nbasis, nocc_up, nocc_dn, dets1, dets2, coeffs1, coeffs2):
    """Restates pyci.compute_transition_rdms(wfn1, wfn2, coeffs1, coeffs2) (rdm.cpp:634-1009; GenCI: intended
    semantics, see pyci_oracle.c)."""
    L = lib()
    d1 = np.ascontiguousarray(dets1, dtype=np.uint64)
    d2 = np.ascontiguousarray(dets2, dtype=np.uint64)
    c1 = np.ascontiguousarray(coeffs1, dtype=np.float64)
    c2 = np.ascontiguousarray(coeffs2, dtype=np.float64)
    n = nbasis
    if kind == DOCI:
        r1, r2 = np.zeros((n, n)), np.zeros((n, n))
        rc = L.oracle_trdms_doci(n, nocc_up, d1.shape[0], _p(d1), d2.shape[0], _p(d2), _p(c1), _p(c2), _p(r1), _p(r2))
    elif kind == FULLCI:
        r1, r2 = np.zeros((2, n, n)), np.zeros((3, n, n, n, n))
        rc = L.oracle_trdms_fullci(n, nocc_up, nocc_dn, d1.shape[0], _p(d1), d2.shape[0], _p(d2), _p(c1), _p(c2),
                                   _p(r1), _p(r2))
    else:
        r1, r2 = np.zeros((n, n)), np.zeros((n, n, n, n))
        rc = L.oracle_trdms_genci(n, nocc_up, d1.shape[0], _p(d1), d2.shape[0], _p(d2), _p(c1), _p(c2), _p(r1), _p(r2))
    if rc != 0:
        raise MemoryError("oracle transition rdms failed")
    return r1, r2


def compute_overlap(dets1, dets2, coeffs1, coeffs2):
    """Restates pyci.compute_overlap(wfn1, wfn2, coeffs1, coeffs2) (overlap.cpp:17-58)."""
    d1 = np.ascontiguousarray(dets1, dtype=np.uint64)
    d2 = np.ascontiguousarray(dets2, dtype=np.uint64)
    nw = int(np.prod(d1.shape[1:]))
    c1 = np.ascontiguousarray(coeffs1, dtype=np.float64)
    c2 = np.ascontiguousarray(coeffs2, dtype=np.float64)
    out = ctypes.c_double(0.0)
    if lib().oracle_overlap(nw, d1.shape[0], _p(d1), d2.shape[0], _p(d2), _p(c1), _p(c2), ctypes.byref(out)) != 0:
        raise MemoryError("oracle overlap failed")
    return out.value


def full_symmetric(indptr, indices, data, n):
    """scipy CSR of L + strict_lower(L)^T from a lower-triangular export."""
    import scipy.sparse as sp
    low = sp.csr_matrix((data, indices, indptr), shape=(n, n))
    return (low + sp.tril(low, -1).T).tocsr()


def lowest_eigenpair(indptr, indices, data, n, symmetric=True, tol=1e-12, ncv=30):
    """Eigenvalue oracle: ARPACK on the exported CSR, standing in for the
    reference's un-vendored Spectra call (sparseop.cpp:127-139).  ecore is NOT added."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    A = full_symmetric(indptr, indices, data, n) if symmetric else \
        sp.csr_matrix((data, indices, indptr), shape=(n, n))
    if n <= 64:
        w, v = np.linalg.eigh(A.toarray())
        return w[0], v[:, 0]
    w, v = sla.eigsh(A, k=1, which="SA", tol=tol, ncv=min(ncv, n - 1))
    return w[0], v[:, 0]


def spinize_rdms(d1, d2):
    """Generalised spin-orbital RDMs from DOCI (d0,d2) or FullCI spin blocks; restates
    pyci/utility.py:94-150 so the energy identity of test_routines.py:130-133 can be checked."""
    n = d1.shape[1]
    r1 = np.zeros((2 * n, 2 * n))
    r2 = np.zeros((2 * n,) * 4)
    a, b = slice(0, n), slice(n, 2 * n)
    if d1.ndim == 2:
        p = np.arange(n)
        r1[a, a][p, p] = d1[p, p]
        r1[b, b][p, p] = d1[p, p]
        for s, t in ((a, b), (b, a)):
            blk = r2[s, t, s, t]
            blk[p[:, None], p[:, None], p[None, :], p[None, :]] += d1
            blk[p[:, None], p[None, :], p[:, None], p[None, :]] += d2
        for s in (a, b):
            r2[s, s, s, s][p[:, None], p[None, :], p[:, None], p[None, :]] += d2
        r2 -= r2.transpose(1, 0, 2, 3)
        r2 -= r2.transpose(0, 1, 3, 2)
        r2 *= 0.5
    else:
        r1[a, a] += d1[0]
        r1[b, b] += d1[1]
        r2[a, a, a, a] += d2[0]
        r2[b, b, b, b] += d2[1]
        r2[a, b, a, b] += d2[2]
        r2[b, a, b, a] += d2[2].transpose(1, 0, 3, 2)
        r2[a, b, b, a] -= d2[2].transpose(0, 1, 3, 2)
        r2[b, a, a, b] -= d2[2].transpose(1, 0, 2, 3)
    return r1, r2
