"""Parity of the sm_100a path (through pyci_b200._pyci -> C ABI -> CUDA kernels) with the golden vectors of
the compiled reference and with the CPU oracle.  Bars (BASELINE.json north_star): indptr/indices bit-exact,
matrix elements <= 1e-12 relative (they are in fact bit-identical), eigenvalues <= 1e-10 Eh."""
import numpy as np
import pytest

from conftest import datafile, seeded_vec, sha
from oracle import oracle as O

pytestmark = pytest.mark.gpu

DATA_RTOL = 1e-12   # matrix elements, relative to the largest |element| of the operator
E_ATOL = 1e-10      # eigenvalues, Hartree

SMALL = [("h4_sto3g", "fullci", (2, 2)), ("lih_sto6g", "fullci", (2, 2)), ("BH_sto-3g_eq", "fullci", (3, 3)),
         ("h6_sto_3g", "fullci", (4, 2)), ("be_ccpvdz", "doci", (2, 2)), ("h2_sto3g", "fullci", (1, 1))]
KIND = {"doci": O.DOCI, "fullci": O.FULLCI, "genci": O.GENCI}


@pytest.fixture(scope="module")
def pyci():
    import pyci_b200
    assert pyci_b200.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return pyci_b200


def make(pyci, fn, kind, occ):
    ham = pyci.hamiltonian(datafile(fn))
    wfn = getattr(pyci, kind + "_wfn")(ham.nbasis, *occ)
    wfn.add_all_dets()
    return ham, wfn


def assert_csr(op, indptr, indices, data):
    assert op.indptr().dtype == np.int64 and op.indices().dtype == np.int64 and op.data().dtype == np.float64
    assert np.array_equal(op.indptr(), indptr)
    assert np.array_equal(op.indices(), indices)
    assert op.size == len(indices)
    if len(data):
        assert np.max(np.abs(op.data() - data)) <= DATA_RTOL * np.max(np.abs(data))


@pytest.mark.parametrize("fn,kind,occ", SMALL)
def test_small_csr_matvec_energy(pyci, small, fn, kind, occ):
    ham, wfn = make(pyci, fn, kind, occ)
    tag = f"{fn}.{kind}{occ[0]}{occ[1]}"
    assert np.array_equal(wfn.to_det_array(), small[tag + ".dets"])
    x = seeded_vec(len(wfn), 11)
    for name, kw in (("sym", dict(symmetric=True)), ("nonsym", dict(symmetric=False)),
                     ("rect", dict(nrow=len(wfn) - 10, symmetric=False))):
        if f"{tag}.{name}.indptr" not in small:
            continue
        op = pyci.sparse_op(ham, wfn, **kw)
        assert_csr(op, small[f"{tag}.{name}.indptr"], small[f"{tag}.{name}.indices"], small[f"{tag}.{name}.data"])
        assert np.array_equal(op.data(), small[f"{tag}.{name}.data"])  # bit-identical in practice
        y = op(x)
        ref = small[f"{tag}.{name}.y"]
        assert y.shape == ref.shape
        np.testing.assert_allclose(y, ref, rtol=0, atol=1e-12 * max(1.0, np.abs(ref).max()))
    op = pyci.sparse_op(ham, wfn)
    if len(wfn) > 1:
        es, cs = op.solve(n=1, tol=1e-10)
        assert abs(es[0] - float(small[tag + ".E0"])) <= E_ATOL
        assert cs.shape == (1, len(wfn))
        r = op(cs[0]) - (es[0] - ham.ecore) * cs[0]
        assert np.linalg.norm(r) < 1e-7


@pytest.mark.parametrize("fn,kind,occ,pinned", [
    ("be_ccpvdz", "fullci", (2, 2), -14.617409507),   # BASELINE config 1 (test_routines.py:44)
    ("h2o_ccpvdz", "doci", (5, 5), -75.634588422),    # BASELINE config 2 (test_routines.py:45)
    ("li2_ccpvdz", "doci", (3, 3), -14.878455349),    # test_routines.py:41
])
def test_configs_digest_and_energy(pyci, digests, fn, kind, occ, pinned):
    ham, wfn = make(pyci, fn, kind, occ)
    tag = f"{fn}.{kind}{occ[0]}{occ[1]}"
    g = digests[tag + ".sym"]
    assert len(wfn) == g["ndet"]
    op = pyci.sparse_op(ham, wfn)
    assert op.size == g["nnz"] and op.shape == (g["ndet"], g["ndet"])
    assert (sha(op.indptr()), sha(op.indices())) == (g["indptr"], g["indices"])
    assert sha(op.data()) == g["data"]        # values bit-identical to the reference
    x = seeded_vec(len(wfn), 11)
    assert abs(np.linalg.norm(op(x)) - g["y_norm"]) <= 1e-11 * g["y_norm"]
    # the reference's own test call: solve(n=1, ncv=30, tol=1e-6), atol 1e-9 (test_routines.py:53-54)
    es, cs = op.solve(n=1, ncv=30, tol=1.0e-6)
    assert abs(es[0] - pinned) <= 1e-9
    es, cs = op.solve(n=1, tol=1.0e-10)
    assert abs(es[0] - g["E0"]) <= E_ATOL
    gn = digests[tag + ".nonsym"]
    opn = pyci.sparse_op(ham, wfn, symmetric=False)
    assert (opn.size, sha(opn.indptr()), sha(opn.indices()), sha(opn.data())) == \
        (gn["nnz"], gn["indptr"], gn["indices"], gn["data"])
    assert abs(np.linalg.norm(opn(x)) - gn["y_norm"]) <= 1e-11 * gn["y_norm"]
    # RDMs (atomics: order of accumulation differs, compare numerically)
    c = seeded_vec(len(wfn), 12)
    c /= np.linalg.norm(c)
    r1, r2 = pyci.compute_rdms(wfn, c)
    gr = digests[tag + ".rdm"]
    assert abs(r1.sum() - gr["rdm1_sum"]) < 1e-10 and abs(np.abs(r2).sum() - gr["rdm2_abs_sum"]) < 1e-9
    o1, o2 = O.compute_rdms(KIND[kind], ham.nbasis, occ[0], occ[1], wfn.to_det_array(), c)
    np.testing.assert_allclose(r1, o1, rtol=0, atol=1e-12)
    np.testing.assert_allclose(r2, o2, rtol=0, atol=1e-12)


@pytest.mark.parametrize("fill", ["default", "warp-specialised"])
@pytest.mark.parametrize("n,occ", [(10, (4, 4)), (9, (4, 3))])
def test_synthetic_fullci_digest(pyci, digests, monkeypatch, n, occ, fill):
    """Both forms of the complete-space fill kernel (fill_complete_kernel, and fill_complete_ws_kernel behind
    PYCI_B200_FILL_WS) against the digests of the compiled reference's operator."""
    if fill == "warp-specialised":
        monkeypatch.setenv("PYCI_B200_FILL_WS", "1")
    ecore, one, two = O.synthetic_integrals(n, 1234)
    ham = pyci.hamiltonian(ecore, one, two)
    wfn = pyci.fullci_wfn(n, *occ)
    wfn.add_all_dets()
    tag = f"syn{n}.fullci{occ[0]}{occ[1]}"
    for name, sym in (("sym", True), ("nonsym", False)):
        g = digests[f"{tag}.{name}"]
        op = pyci.sparse_op(ham, wfn, symmetric=sym)
        assert (op.size, sha(op.indptr()), sha(op.indices()), sha(op.data())) == \
            (g["nnz"], g["indptr"], g["indices"], g["data"])
        if sym:
            es, _ = op.solve(n=1, tol=1e-10)
            assert abs(es[0] - g["E0"]) <= E_ATOL


def test_selected_space_and_rdms(pyci, small):
    ecore, one, two = O.synthetic_integrals(8, 1234)
    ham = pyci.hamiltonian(ecore, one, two)
    dets = small["syn8.fullci32.sel.dets"]
    wfn = pyci.fullci_wfn(8, 3, 2, dets)
    for name, sym in (("sym", True), ("nonsym", False)):
        op = pyci.sparse_op(ham, wfn, symmetric=sym)
        assert_csr(op, small[f"syn8.fullci32.sel.{name}.indptr"], small[f"syn8.fullci32.sel.{name}.indices"],
                   small[f"syn8.fullci32.sel.{name}.data"])
    c = seeded_vec(len(wfn), 12)
    c /= np.linalg.norm(c)
    r1, r2 = pyci.compute_rdms(wfn, c)
    np.testing.assert_allclose(r1, small["syn8.fullci32.sel.rdm1"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(r2, small["syn8.fullci32.sel.rdm2"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("fn,kind,occ", SMALL)
def test_small_rdms(pyci, small, fn, kind, occ):
    ham, wfn = make(pyci, fn, kind, occ)
    tag = f"{fn}.{kind}{occ[0]}{occ[1]}"
    c = seeded_vec(len(wfn), 12)
    c /= np.linalg.norm(c)
    r1, r2 = pyci.compute_rdms(wfn, c)
    assert r1.shape == small[tag + ".rdm1"].shape and r2.shape == small[tag + ".rdm2"].shape
    np.testing.assert_allclose(r1, small[tag + ".rdm1"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(r2, small[tag + ".rdm2"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("fn,occ", [("h4_sto3g", (2, 2)), ("BH_sto-3g_eq", (3, 3)), ("h6_sto_3g", (4, 2))])
def test_genci(pyci, genci_golden, fn, occ):
    """GenCI equals (i) the reference compiled with the two loop bounds fixed and (ii) FullCI on the spatial
    integrals; its RDMs equal the spin-expanded FullCI RDMs."""
    ecore, one, two = O.read_fcidump(datafile(fn))
    n = one.shape[0]
    h2, g2 = O.spin_orbital_integrals(one, two)
    tag = f"{fn}.genci{sum(occ)}"
    gd = genci_golden[tag + ".dets"]
    hamg = pyci.hamiltonian(ecore, h2, g2)
    wg = pyci.genci_wfn(2 * n, sum(occ), 0, gd)
    fw = pyci.fullci_wfn(n, *occ)
    fw.add_all_dets()
    assert np.array_equal(pyci.genci_wfn(fw).to_det_array(), gd)
    hamf = pyci.hamiltonian(ecore, one, two)
    for name, sym in (("sym", True), ("nonsym", False)):
        op = pyci.sparse_op(hamg, wg, symmetric=sym)
        assert_csr(op, genci_golden[f"{tag}.{name}.indptr"], genci_golden[f"{tag}.{name}.indices"],
                   genci_golden[f"{tag}.{name}.data"])
        opf = pyci.sparse_op(hamf, fw, symmetric=sym)
        assert np.array_equal(op.indptr(), opf.indptr()) and np.array_equal(op.indices(), opf.indices())
        assert np.array_equal(op.data(), opf.data())
    es, cs = pyci.sparse_op(hamg, wg).solve(n=1, tol=1e-10)
    esf, csf = pyci.sparse_op(hamf, fw).solve(n=1, tol=1e-10)
    assert abs(es[0] - esf[0]) <= E_ATOL
    c = seeded_vec(len(fw), 2)
    c /= np.linalg.norm(c)
    s1, s2 = pyci.spinize_rdms(*pyci.compute_rdms(fw, c))
    g1, g2r = pyci.compute_rdms(wg, c)
    np.testing.assert_allclose(g1, s1, rtol=0, atol=1e-13)
    np.testing.assert_allclose(g2r, s2, rtol=0, atol=1e-13)


def test_rdm_energy_identity(pyci):
    """test_routines.py:83-133 restated: traces, DOCI energy from reduced integrals, spin-orbital identity."""
    for fn, kind, occ in (("be_ccpvdz", "doci", (2, 2)), ("h2o_ccpvdz", "doci", (5, 5)), ("lih_sto6g", "fullci", (2, 2))):
        ham, wfn = make(pyci, fn, kind, occ)
        op = pyci.sparse_op(ham, wfn)
        es, cs = op.solve(n=1, ncv=30, tol=1.0e-6)
        d1, d2 = pyci.compute_rdms(wfn, cs[0])
        if kind == "doci":
            assert abs(np.trace(d1) - wfn.nocc_up) < 1e-9
            assert abs(np.sum(d2) - wfn.nocc_up * (wfn.nocc_up - 1)) < 1e-9
            k0, k2 = pyci.reduce_senzero_integrals(ham.h, ham.v, ham.w, wfn.nocc_up)
            e = ham.ecore + np.einsum("ij,ij", k0, d1) + np.einsum("ij,ij", k2, d2)
            assert abs(e - es[0]) < 1e-9
        rdm1, rdm2 = pyci.spinize_rdms(d1, d2)
        _, one, two = O.read_fcidump(datafile(fn))
        h2, g2 = O.spin_orbital_integrals(one, two)
        anti = g2 - g2.transpose(0, 1, 3, 2)
        e = ham.ecore + np.einsum("ij,ij", h2, rdm1) + 0.25 * np.einsum("ijkl,ijkl", anti, rdm2)
        assert abs(e - es[0]) < 1e-9
        assert np.all(np.abs(rdm1 - rdm1.T) < 1e-9)


def test_rectangular_like_reference(pyci):
    """test_routines.py:57-80: excitation-selected space, nrow = len-10, symmetric=False, op(ones) equals the
    row sums of get_element."""
    ham = pyci.hamiltonian(datafile("be_ccpvdz"))
    wfn = pyci.fullci_wfn(ham.nbasis, 2, 2)
    pyci.add_excitations(wfn, 0, 1, 2)
    nrow = len(wfn) - 10
    op = pyci.sparse_op(ham, wfn, nrow, symmetric=False)
    assert op.shape == (nrow, len(wfn))
    y = op(np.ones(op.shape[1], dtype=pyci.c_double))
    assert y.ndim == 1 and y.shape[0] == nrow
    ip, ix, dv = op.indptr(), op.indices(), op.data()
    z = np.add.reduceat(dv, ip[:-1])
    np.testing.assert_allclose(y, z, rtol=0, atol=1e-11)
    rng = np.random.default_rng(3)
    for i in rng.integers(0, nrow, 5):
        row = dict(zip(ix[ip[i]:ip[i + 1]].tolist(), dv[ip[i]:ip[i + 1]].tolist()))
        for j in list(row)[:4] + rng.integers(0, len(wfn), 4).tolist():
            assert op.get_element(int(i), int(j)) == row.get(int(j), 0.0)
    dets = wfn.to_det_array()
    oi, ox, od = O.sparse_op(O.FULLCI, ham.nbasis, 2, 2, dets, (ham.one_mo, ham.two_mo), nrow=nrow, symmetric=False)
    assert_csr(op, oi, ox, od)


def test_row_pointer_results_own_their_recycled_buffers(pyci, monkeypatch):
    """op.indptr() returns arrays over page-locked buffers that are recycled once numpy drops the array: an array that
    is still alive keeps its contents when later calls run, outlives its operator, is writable like the reference's
    copy, and equals what the pageable path returns."""
    ham = pyci.hamiltonian(datafile("be_ccpvdz"))
    wfn = pyci.fullci_wfn(ham.nbasis, 2, 2)
    pyci.add_excitations(wfn, 0, 1, 2)
    op = pyci.sparse_op(ham, wfn)
    a = op.indptr()
    keep = a.copy()
    assert a.dtype == np.int64 and a.shape == (len(wfn) + 1,) and a.flags.c_contiguous and a.flags.writeable
    b = op.indptr()
    assert np.array_equal(a, b) and a.ctypes.data != b.ctypes.data
    addr = b.ctypes.data
    del b
    small = pyci.sparse_op(ham, wfn, 7, symmetric=False).indptr()  # another size class: does not touch a's buffer
    c = op.indptr()  # the buffer b gave back
    assert c.ctypes.data == addr and np.array_equal(c, keep)
    del op
    assert np.array_equal(a, keep) and small.shape == (8,)
    a[0] = 5  # the caller's array: writing it does not reach the library
    monkeypatch.setenv("PYCI_B200_NO_PINNED_RESULTS", "1")
    op2 = pyci.sparse_op(ham, wfn)
    assert np.array_equal(op2.indptr(), keep)


def test_update_after_adding_determinants(pyci):
    """HCI-style growth (sparseop.cpp:175-178): update(ham, wfn) after appending determinants gives the same
    operator as a fresh build."""
    ham = pyci.hamiltonian(datafile("lih_sto6g"))
    wfn = pyci.fullci_wfn(ham.nbasis, 2, 2)
    pyci.add_excitations(wfn, 0, 1)
    op = pyci.sparse_op(ham, wfn)
    n0 = op.shape[0]
    pyci.add_excitations(wfn, 2)
    assert len(wfn) > n0
    op.update(ham, wfn)
    fresh = pyci.sparse_op(ham, wfn)
    assert op.shape == fresh.shape == (len(wfn), len(wfn))
    assert np.array_equal(op.indptr(), fresh.indptr()) and np.array_equal(op.indices(), fresh.indices())
    assert np.array_equal(op.data(), fresh.data())
    oi, ox, od = O.sparse_op(O.FULLCI, ham.nbasis, 2, 2, wfn.to_det_array(), (ham.one_mo, ham.two_mo))
    assert_csr(op, oi, ox, od)


def test_solve_guards_and_edge_cases(pyci):
    ham = pyci.hamiltonian(datafile("h2_sto3g"))
    wfn = pyci.fullci_wfn(ham.nbasis, 1, 1)
    wfn.add_all_dets()
    op = pyci.sparse_op(ham, wfn)
    with pytest.raises(ValueError):          # sparseop.cpp:116-117
        op.solve(n=len(wfn))
    rect = pyci.sparse_op(ham, wfn, len(wfn) - 1, symmetric=False)
    with pytest.raises(TypeError):           # sparseop.cpp:118-119
        rect.solve(n=1)
    es, cs = op.solve(n=len(wfn) - 1, tol=1e-10)     # several roots (test_odometer.py style)
    ip, ix, dv = O.sparse_op(O.FULLCI, ham.nbasis, 1, 1, wfn.to_det_array(), (ham.one_mo, ham.two_mo))
    w = np.linalg.eigvalsh(O.full_symmetric(ip, ix, dv, len(wfn)).toarray()) + ham.ecore
    np.testing.assert_allclose(es, w[:len(es)][::-1], rtol=0, atol=1e-9)   # largest of the n lowest first
    # single determinant: E = H00 + ecore, c = [1]  (sparseop.cpp:120-124)
    one = pyci.fullci_wfn(ham.nbasis, 1, 1)
    one.add_hartreefock_det()
    op1 = pyci.sparse_op(ham, one)
    es, cs = op1.solve()
    assert es.shape == (1,) and cs.shape == (1, 1) and cs[0, 0] == 1.0
    assert es[0] == op1.get_element(0, 0) + ham.ecore
    # start vector
    ham = pyci.hamiltonian(datafile("lih_sto6g"))
    wfn = pyci.fullci_wfn(ham.nbasis, 2, 2)
    wfn.add_all_dets()
    op = pyci.sparse_op(ham, wfn)
    e_ref, _ = op.solve(n=1, tol=1e-10)
    c0 = np.zeros(len(wfn))
    c0[0] = 1.0
    e_c0, _ = op.solve(n=1, c0=c0, tol=1e-10)
    assert abs(e_ref[0] - e_c0[0]) <= E_ATOL
    e3, c3 = op.solve(n=3, tol=1e-9)
    ip, ix, dv = O.sparse_op(O.FULLCI, ham.nbasis, 2, 2, wfn.to_det_array(), (ham.one_mo, ham.two_mo))
    w = np.linalg.eigvalsh(O.full_symmetric(ip, ix, dv, len(wfn)).toarray()) + ham.ecore
    # the three lowest, largest first: Spectra's default sorting of the selected pairs (sparseop.cpp:134), which
    # pyci/test/test_odometer.py relies on
    np.testing.assert_allclose(e3, w[:3][::-1], rtol=0, atol=1e-9)
    for k in range(3):  # every vector sits beside its own value
        assert np.linalg.norm(op(c3[k]) - (e3[k] - ham.ecore) * c3[k]) < 1e-6


def test_unsupported_and_invalid_inputs_fail_loudly(pyci):
    ecore, one, two = O.synthetic_integrals(66, 7)
    ham = pyci.hamiltonian(ecore, one, two)
    wfn = pyci.doci_wfn(66, 2, 2)
    wfn.add_all_dets()
    with pytest.raises(RuntimeError):        # nbasis > 64: construction / SpMV / solve only (multiword.cu)
        pyci.compute_rdms(wfn, np.ones(len(wfn)) / np.sqrt(len(wfn)))
    with pytest.raises(RuntimeError):
        pyci.add_hci(pyci.hamiltonian(*O.synthetic_integrals(66, 7)), pyci.genci_wfn(66, 3, 0, np.array([[7, 0]], dtype=np.uint64)),
                     np.ones(1), eps=1e-3)
    ham = pyci.hamiltonian(datafile("h4_sto3g"))
    dets = np.array([[3, 3], [3, 3]], dtype=np.uint64)  # duplicate determinant
    with pytest.raises(ValueError):
        pyci.sparse_op(ham, pyci.fullci_wfn(4, 2, 2, dets))


def test_key_modes(pyci):
    """64-bit keys (FullCI 16 < nbasis <= 32) and 128-bit keys (nbasis > 32); GenCI 64-bit (nbasis > 32)."""
    for n, occ in ((20, (2, 1)), (40, (1, 1))):
        ecore, one, two = O.synthetic_integrals(n, 3)
        ham = pyci.hamiltonian(ecore, one, two)
        wfn = pyci.fullci_wfn(n, *occ)
        wfn.add_all_dets()
        op = pyci.sparse_op(ham, wfn)
        oi, ox, od = O.sparse_op(O.FULLCI, n, occ[0], occ[1], wfn.to_det_array(), (one, two))
        assert_csr(op, oi, ox, od)
        assert np.array_equal(op.data(), od)
    n = 36
    ecore, one, two = O.synthetic_integrals(n, 4)
    ham = pyci.hamiltonian(ecore, one, two)
    wfn = pyci.genci_wfn(n, 2, 0)
    pyci.add_excitations(wfn, 0, 1, 2)
    op = pyci.sparse_op(ham, wfn, symmetric=False)
    oi, ox, od = O.sparse_op(O.GENCI, n, 2, 0, wfn.to_det_array(), (one, two), symmetric=False)
    assert_csr(op, oi, ox, od)
    wfn = pyci.doci_wfn(40, 2, 2)
    wfn.add_all_dets()
    op = pyci.sparse_op(pyci.hamiltonian(*O.synthetic_integrals(40, 5)), wfn)
    _, one, two = O.synthetic_integrals(40, 5)
    oi, ox, od = O.sparse_op(O.DOCI, 40, 2, 2, wfn.to_det_array(), O.senzero_integrals(one, two))
    assert_csr(op, oi, ox, od)


@pytest.mark.parametrize("n,occ", [(9, (3, 3)), (8, (4, 2)), (10, (3, 2)), (7, (1, 1)), (6, (5, 5))])
def test_fill_paths_sorted_incomplete_unsorted_rectangular(pyci, n, occ):
    """The three fill strategies give the same CSR as the oracle: sorted complete space (direct slots),
    sorted incomplete / rectangular (slots + hit compaction), shuffled determinant order (radix sort)."""
    ecore, one, two = O.synthetic_integrals(n, 99)
    ham = pyci.hamiltonian(ecore, one, two)
    full = pyci.fullci_wfn(n, *occ)
    full.add_all_dets()
    dets = full.to_det_array()
    rng = np.random.default_rng(5)
    keep = np.sort(rng.choice(len(dets), size=max(2, (2 * len(dets)) // 3), replace=False))
    cases = {
        "sorted-complete": (dets, {}),
        "sorted-complete-rect": (dets, dict(nrow=len(dets) - 3, ncol=len(dets) - 5, symmetric=False)),
        "sorted-incomplete": (dets[keep], {}),
        "sorted-incomplete-nonsym": (dets[keep], dict(symmetric=False)),
        "shuffled": (dets[rng.permutation(len(dets))], {}),
        "shuffled-incomplete": (dets[rng.permutation(keep)], dict(symmetric=False)),
    }
    for name, (d, kw) in cases.items():
        wfn = pyci.fullci_wfn(n, occ[0], occ[1], np.ascontiguousarray(d))
        op = pyci.sparse_op(ham, wfn, **kw)
        oi, ox, od = O.sparse_op(O.FULLCI, n, occ[0], occ[1], d, (one, two), **kw)
        assert np.array_equal(op.indptr(), oi), name
        assert np.array_equal(op.indices(), ox), name
        assert np.array_equal(op.data(), od), name
        x = seeded_vec(op.shape[1], 3)
        y = op(x)
        yo = O.matvec(oi, ox, od, x, kw.get("symmetric", True))
        np.testing.assert_allclose(y, yo, rtol=0, atol=1e-12 * max(1.0, np.abs(yo).max()), err_msg=name)


def test_config5_style_selected_genci(pyci):
    """Config 5 in miniature: a seniority-zero selection inside a GenCI spin-orbital space (most excitation
    candidates miss the index), CSR against the oracle, E0 against ARPACK, RDM energy identity."""
    from pyci_b200.synthetic import seniority_zero_genci_dets, spin_orbital_integrals, synthetic_integrals
    K, P = 9, 3
    _, one, two = synthetic_integrals(K, 1234)
    h2, g2 = spin_orbital_integrals(one, two)
    ham = pyci.hamiltonian(0.0, h2, g2)
    for ndet in (None, 60):
        dets = seniority_zero_genci_dets(K, P, ndet)
        wfn = pyci.genci_wfn(2 * K, 2 * P, 0, dets)
        assert np.array_equal(wfn.to_det_array(), dets)
        op = pyci.sparse_op(ham, wfn)
        oi, ox, od = O.sparse_op(O.GENCI, 2 * K, 2 * P, 0, dets, (h2, g2))
        assert_csr(op, oi, ox, od)
        assert np.array_equal(op.data(), od)
        es, cs = op.solve(n=1, tol=1e-10)
        e0, _ = O.lowest_eigenpair(oi, ox, od, len(wfn))
        assert abs(es[0] - e0) <= E_ATOL
        r1, r2 = pyci.compute_rdms(wfn, cs[0])
        o1, o2 = O.compute_rdms(O.GENCI, 2 * K, 2 * P, 0, dets, cs[0])
        np.testing.assert_allclose(r1, o1, rtol=0, atol=1e-12)
        np.testing.assert_allclose(r2, o2, rtol=0, atol=1e-12)
        anti = g2 - g2.transpose(0, 1, 3, 2)
        e = np.einsum("ij,ij", h2, r1) + 0.25 * np.einsum("ijkl,ijkl", anti, r2)
        assert abs(e - es[0]) < 1e-9


# ---- add_hci / compute_enpt2 (hci.cpp, enpt2.cpp) ----------------------------------------------------
from conftest import HCI_CASES, sorted_rows  # noqa: E402

PT2_RTOL = 1e-12  # ENPT2 energy: fp64 atomics sum the external terms in a different order than the hash map


@pytest.mark.parametrize("tag,fn,kind,occ,steps", HCI_CASES)
def test_add_hci_and_enpt2_against_reference_golden(pyci, hci_golden, tag, fn, kind, occ, steps):
    """Same inputs as the compiled reference's trajectory: the appended SET equals the reference's, the order is
    the oracle's first-encounter order, the ENPT2 energies match."""
    ham = pyci.hamiltonian(datafile(fn))
    ecore, one, two = O.read_fcidump(datafile(fn))
    n = one.shape[0]
    ints = O.senzero_integrals(one, two) if kind == "doci" else (one, two)
    for it in range(steps):
        g = {k: hci_golden[f"{tag}.{it}.{k}"] for k in ("dets", "coeffs", "energy", "eps", "enpt2", "enpt2_tight", "new_sorted")}
        eps, e, c = float(g["eps"]), float(g["energy"]), g["coeffs"]
        wfn = getattr(pyci, kind + "_wfn")(n, occ[0], occ[1], g["dets"])
        for key, ee in (("enpt2", eps), ("enpt2_tight", eps * 1e-2)):
            pt = pyci.compute_enpt2(ham, wfn, c, e, ee)
            assert abs(pt - float(g[key])) <= PT2_RTOL * abs(float(g[key]))
        before = len(wfn)
        nadd = pyci.add_hci(ham, wfn, c, eps=eps)
        new = wfn.to_det_array()[before:]
        assert nadd == len(new) == len(g["new_sorted"]) and len(wfn) == before + nadd
        assert np.array_equal(sorted_rows(new), g["new_sorted"])
        assert np.array_equal(new, O.add_hci(KIND[kind], n, occ[0], occ[1], g["dets"], ints, c, eps))
        # the grown wave function is usable: every new determinant is indexed where it was appended
        for j in (0, nadd // 2, nadd - 1):
            if nadd:
                assert wfn.index_det(new[j]) == before + j
        assert pyci.add_hci(ham, wfn, np.zeros(len(wfn)), eps=eps) == 0  # c_i = 0: eps / |c_i| = inf


def test_hci_loop_like_reference_test(pyci):
    """test_routines.py:435-460 restated (run_hci): HF -> [sparse_op.solve -> add_hci -> op.update]* reaches the
    FullCI energy of Be cc-pVDZ; the ENPT2 estimate of an intermediate space lies between E_var and E_FCI-ish."""
    ham = pyci.hamiltonian(datafile("be_ccpvdz"))
    wfn = pyci.fullci_wfn(ham.nbasis, 2, 2)
    wfn.add_hartreefock_det()
    op = pyci.sparse_op(ham, wfn)
    es, cs = op.solve(n=1, tol=1e-9)
    sizes = [len(wfn)]
    pt_mid = None
    while True:
        nadd = pyci.add_hci(ham, wfn, cs[0], eps=1.0e-5)
        if nadd == 0:
            break
        op.update(ham, wfn)
        es, cs = op.solve(n=1, tol=1e-9)
        sizes.append(len(wfn))
        if pt_mid is None:
            pt_mid = (es[0], pyci.compute_enpt2(ham, wfn, cs[0], es[0], 1.0e-6))
    assert sizes == sorted(sizes) and sizes[-1] > 1000
    assert abs(es[0] - (-14.617409507)) < 1e-7   # test_routines.py:44 with the eps = 1e-5 truncation
    assert pt_mid[1] < pt_mid[0]                 # second-order correction is negative
    assert op.shape == (len(wfn), len(wfn))


@pytest.mark.parametrize("n,occ,frac", [(8, (3, 2), 5), (10, (3, 3), 9), (34, (2, 1), 3)])
def test_hci_enpt2_synthetic_selected_spaces(pyci, n, occ, frac):
    """Synthetic integrals (every element non-zero), a thinned FullCI space: exact order vs the oracle, ENPT2 to 1e-12,
    GenCI on spin-orbital integrals gives the image of the FullCI selection.  n = 34 exercises 128-bit two-spin keys."""
    ecore, one, two = O.synthetic_integrals(n, 4321)
    ham = pyci.hamiltonian(ecore, one, two)
    full = pyci.fullci_wfn(n, *occ)
    full.add_all_dets()
    fd = np.ascontiguousarray(full.to_det_array()[::frac][:4000])
    wfn = pyci.fullci_wfn(n, occ[0], occ[1], fd)
    c = seeded_vec(len(fd), 3)
    c /= np.linalg.norm(c)
    eps = 0.02
    pt = pyci.compute_enpt2(ham, wfn, c, -1.0, eps)
    pto, nt = O.compute_enpt2(O.FULLCI, n, occ[0], occ[1], fd, (one, two), c, -1.0, ecore, eps)
    assert nt > 0 and abs(pt - pto) <= PT2_RTOL * abs(pto)
    ref_new = O.add_hci(O.FULLCI, n, occ[0], occ[1], fd, (one, two), c, eps)
    nadd = pyci.add_hci(ham, wfn, c, eps=eps)
    assert nadd == len(ref_new) > 0
    assert np.array_equal(wfn.to_det_array()[len(fd):], ref_new)
    if 2 * n <= 64:
        h2, g2 = O.spin_orbital_integrals(one, two)
        gd = np.ascontiguousarray((fd[:, 0] | (fd[:, 1] << np.uint64(n))).reshape(-1, 1))
        wg = pyci.genci_wfn(2 * n, sum(occ), 0, gd)
        hamg = pyci.hamiltonian(ecore, h2, g2)
        ptg = pyci.compute_enpt2(hamg, wg, c, -1.0, eps)
        assert abs(ptg - pto) <= PT2_RTOL * abs(pto)
        assert pyci.add_hci(hamg, wg, c, eps=eps) == nadd
        newg = wg.to_det_array()[len(gd):]
        assert np.array_equal(newg, O.add_hci(O.GENCI, 2 * n, sum(occ), 0, gd, (h2, g2), c, eps))
        imgs = (ref_new[:, 0] | (ref_new[:, 1] << np.uint64(n))).reshape(-1, 1)
        assert np.array_equal(sorted_rows(newg), sorted_rows(imgs))


def test_hci_table_growth_and_c_abi(pyci):
    """A walk that overflows the first external table (2^16 slots) is re-run with a larger one; through the C ABI
    (ctypes) with device-resident handles, as a reference-side binding would call it."""
    from pyci_b200 import cabi
    n, occ = 12, (3, 3)
    ecore, one, two = O.synthetic_integrals(n, 99)
    ctx = cabi.Context(0)
    dets = O.all_dets(O.FULLCI, n, *occ)[::40]
    c = seeded_vec(len(dets), 8)
    ham = cabi.Ham(ctx, n, ecore, one, two)
    wfn = cabi.Wfn(ctx, cabi.FULLCI, n, occ[0], occ[1], dets)
    ref_new = O.add_hci(O.FULLCI, n, occ[0], occ[1], dets, (one, two), c, 1e-4)
    assert len(ref_new) > 0.6 * 65536  # more than the first table admits
    pt, nt = wfn.compute_enpt2(ham, c, -5.0, 1e-4)
    pto, nto = O.compute_enpt2(O.FULLCI, n, occ[0], occ[1], dets, (one, two), c, -5.0, ecore, 1e-4)
    assert nt == nto == len(ref_new) and abs(pt - pto) <= PT2_RTOL * abs(pto)
    new = wfn.add_hci(ham, c, 1e-4)
    assert np.array_equal(new, ref_new)
    assert wfn.ndet == len(dets) + len(ref_new) and wfn.ext_seconds() > 0
    assert np.array_equal(wfn.index_dets(new[:100]), len(dets) + np.arange(100))
    with pytest.raises(cabi.PyciError):
        cabi.Wfn(ctx, cabi.DOCI, n, 3, 3, np.unique(dets[:, 0]).reshape(-1, 1)).compute_enpt2(ham, c, -5.0, 1e-4)  # DOCI: FullCI image only
    wfn.close()
    ham.close()
    ctx.close()


def test_hci_enpt2_row_chunks_merge(pyci, monkeypatch):
    """PYCI_B200_EXT_SPLIT walks the rows in chunks with one external table each and merges the chunk lists with the
    kernel that also merges the per-rank lists of the row-sharded path: same result as the single walk."""
    n, occ = 10, (3, 3)
    ecore, one, two = O.synthetic_integrals(n, 4321)
    ham = pyci.hamiltonian(ecore, one, two)
    full = pyci.fullci_wfn(n, *occ)
    full.add_all_dets()
    fd = np.ascontiguousarray(full.to_det_array()[::9])
    c = seeded_vec(len(fd), 3)
    c /= np.linalg.norm(c)
    pto, nt = O.compute_enpt2(O.FULLCI, n, occ[0], occ[1], fd, (one, two), c, -1.0, ecore, 0.02)
    ref_new = O.add_hci(O.FULLCI, n, occ[0], occ[1], fd, (one, two), c, 0.02)
    gd = np.ascontiguousarray((fd[:, 0] | (fd[:, 1] << np.uint64(n))).reshape(-1, 1))
    h2, g2 = O.spin_orbital_integrals(one, two)
    hamg = pyci.hamiltonian(ecore, h2, g2)
    ref_newg = O.add_hci(O.GENCI, 2 * n, sum(occ), 0, gd, (h2, g2), c, 0.02)
    for split in ("2", "3", "7"):
        monkeypatch.setenv("PYCI_B200_EXT_SPLIT", split)
        wfn = pyci.fullci_wfn(n, occ[0], occ[1], fd)
        pt = pyci.compute_enpt2(ham, wfn, c, -1.0, 0.02)
        assert abs(pt - pto) <= PT2_RTOL * abs(pto), (split, pt, pto)
        assert pyci.add_hci(ham, wfn, c, eps=0.02) == len(ref_new)
        assert np.array_equal(wfn.to_det_array()[len(fd):], ref_new)
        wg = pyci.genci_wfn(2 * n, sum(occ), 0, gd)
        assert abs(pyci.compute_enpt2(hamg, wg, c, -1.0, 0.02) - pto) <= PT2_RTOL * abs(pto)
        assert pyci.add_hci(hamg, wg, c, eps=0.02) == len(ref_newg)
        assert np.array_equal(wg.to_det_array()[len(gd):], ref_newg)


@pytest.mark.parametrize("kind,n,occ", [("fullci", 10, (3, 3)), ("genci", 16, (5, 0)), ("doci", 20, (4, 4))])
def test_incremental_update_equals_fresh_build(pyci, kind, n, occ):
    """pyci_op_update (SparseOp::update, sparseop.cpp:175-201) through the C ABI: grow an operator twice and compare
    with fresh builds -- exported CSR bit-identical, full-row SpMV and E0 equal, repeated and empty updates."""
    from pyci_b200 import cabi
    ecore, one, two = O.synthetic_integrals(n, 77)
    okind = KIND[kind]
    alld = O.all_dets(okind, n, *occ)
    rng = np.random.default_rng(5)
    alld = np.ascontiguousarray(alld[rng.permutation(len(alld))[:2500]])  # a selected, unsorted space
    ints = O.senzero_integrals(one, two) if kind == "doci" else (one, two)
    ckind = {"doci": cabi.DOCI, "fullci": cabi.FULLCI, "genci": cabi.GENCI}[kind]
    ctx = cabi.Context(0)
    h, v, w = O.senzero_integrals(one, two)
    ham = cabi.Ham(ctx, n, ecore, one, two, h, v, w)
    cuts = (700, 1600, 1600, 2500)
    wfn = cabi.Wfn(ctx, ckind, n, occ[0], occ[1], alld[:cuts[0]])
    op = cabi.Op(ctx, ham, wfn)
    wfn.close()
    x = seeded_vec(2500, 9)
    for cut in cuts[1:]:
        wfn = cabi.Wfn(ctx, ckind, n, occ[0], occ[1], alld[:cut])
        op.update(ham, wfn)
        fresh = cabi.Op(ctx, ham, wfn)
        assert (op.nrow, op.ncol, op.size, op.stored_nnz) == (cut, cut, fresh.size, fresh.stored_nnz)
        a, b = op.export_csr(), fresh.export_csr()
        assert all(np.array_equal(p, q) for p, q in zip(a, b))
        oi, ox, od = O.sparse_op(okind, n, occ[0], occ[1], alld[:cut], ints)
        assert np.array_equal(a[0], oi) and np.array_equal(a[1], ox) and np.array_equal(a[2], od)
        y, yf = op.matvec(x[:cut]), fresh.matvec(x[:cut])
        assert np.max(np.abs(y - yf)) <= 1e-12 * np.max(np.abs(yf))
        e, _, _ = op.solve(n=1, tol=1e-9)
        ef, _, _ = fresh.solve(n=1, tol=1e-9)
        assert abs(e[0] - ef[0]) <= E_ATOL
        for i, j in ((cut - 1, 0), (cut // 2, cut // 2), (cut - 1, cut - 2)):
            assert op.get_element(i, j) == fresh.get_element(i, j)
        fresh.close()
        wfn.close()
    # operators the incremental path does not cover are refused (the host class rebuilds them)
    wfn = cabi.Wfn(ctx, ckind, n, occ[0], occ[1], alld[:900])
    rect = cabi.Op(ctx, ham, wfn, nrow=800, ncol=900, symmetric=True)
    with pytest.raises(cabi.PyciError) as ei:
        rect.update(ham, wfn)
    assert ei.value.status == cabi.ERR_UNSUPPORTED
    rect.close()
    # a non-symmetric operator grows like the reference's: rows appended, old rows keep the columns they were built with
    rect = cabi.Op(ctx, ham, wfn, nrow=800, ncol=900, symmetric=False)
    rect.update(ham, wfn)  # same wave function: rows 800..899 appear, all 900 columns
    wfn.close()
    wfn = cabi.Wfn(ctx, ckind, n, occ[0], occ[1], alld[:1500])
    rect.update(ham, wfn)
    rect.update(ham, wfn)  # nothing to add
    oi, ox, od = O.sparse_op_updated(okind, n, occ[0], occ[1], alld[:1500], ints, 800, [900, 900, 1500], symmetric=False)
    a = rect.export_csr()
    assert (rect.nrow, rect.ncol, rect.size) == (1500, 1500, len(ox))
    assert np.array_equal(a[0], oi) and np.array_equal(a[1], ox) and np.array_equal(a[2], od)
    y = rect.matvec(x[:1500])
    np.testing.assert_allclose(y, O.matvec(oi, ox, od, x[:1500], False), rtol=0, atol=1e-11)
    assert rect.get_element(10, 1400) == 0.0  # an old row never gained the new columns
    rect.close()
    wfn.close()
    op.close()
    ham.close()
    ctx.close()


from conftest import UPDATE_CASES, UPDATE_MODES  # noqa: E402


@pytest.mark.parametrize("tag,fn,kind,occ", UPDATE_CASES)
@pytest.mark.parametrize("mode,symm", UPDATE_MODES)
def test_update_like_the_compiled_reference(pyci, update_golden, tag, fn, kind, occ, mode, symm):
    """op.update(ham, wfn) through the host module, the calls of tests/golden/make_golden_update.py: the operator after
    two growth steps equals the compiled reference's -- symmetric (rows appended to the lower triangle), non-symmetric
    and rectangular non-symmetric (old rows keep their columns)."""
    g, key = update_golden, "%s.%s" % (tag, mode)
    ham = pyci.hamiltonian(datafile(fn))
    wfn = getattr(pyci, kind + "_wfn")(ham.nbasis, *occ)
    levels = {"fullci": [(0, 1), (2,), (3,)], "doci": [(0, 1), (2,)]}[kind]
    pyci.add_excitations(wfn, *levels[0])
    sizes, nrow0 = g[key + ".sizes"].tolist(), int(g[key + ".nrow0"])
    assert len(wfn) == sizes[0]
    op = pyci.sparse_op(ham, wfn, nrow0, len(wfn), symmetric=symm)
    for k, lv in enumerate(levels[1:], 1):
        pyci.add_excitations(wfn, *lv)
        assert len(wfn) == sizes[k]
        op.update(ham, wfn)
    assert np.array_equal(wfn.to_det_array(), g[key + ".dets"])
    assert op.shape == tuple(g[key + ".shape"]) and op.size == int(g[key + ".nnz"])
    assert np.array_equal(op.indptr(), g[key + ".indptr"])
    assert sha(op.indices()) == str(g[key + ".indices.sha256"]) and sha(op.data()) == str(g[key + ".data.sha256"])
    if symm:
        es, _ = op.solve(n=1, tol=1e-9)
        fresh = pyci.sparse_op(ham, wfn)
        ef, _ = fresh.solve(n=1, tol=1e-9)
        assert abs(es[0] - ef[0]) <= E_ATOL
    else:
        xx = seeded_vec(len(wfn), 4)
        ip, ix, dv = op.indptr(), op.indices(), op.data()
        np.testing.assert_allclose(op(xx), O.matvec(ip, ix, dv, xx, False), rtol=0, atol=1e-11)


# ---- compute_transition_rdms / compute_overlap (rdm.cpp:634-1009, overlap.cpp) ---------------------------
from conftest import TRDM_CASES  # noqa: E402


def close(a, b):
    """fp64 atomics sum in a different order from run to run: 1e-13 relative to the largest element"""
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-13 * max(1.0, float(np.max(np.abs(b)))))


@pytest.mark.parametrize("tag,kind,n,occ", TRDM_CASES)
def test_transition_rdms_and_overlap(pyci, trdm_golden, tag, kind, n, occ):
    g = {k: trdm_golden[f"{tag}.{k}"] for k in ("dets1", "dets2", "c1", "c2", "rdm1", "rdm2", "overlap")}
    cls = getattr(pyci, kind + "_wfn")
    w1, w2 = cls(n, occ[0], occ[1], g["dets1"]), cls(n, occ[0], occ[1], g["dets2"])
    r1, r2 = pyci.compute_transition_rdms(w1, w2, g["c1"], g["c2"])
    assert r1.shape == g["rdm1"].shape and r2.shape == g["rdm2"].shape
    close(r1, g["rdm1"])
    close(r2, g["rdm2"])
    otol = 1e-13 * max(1.0, abs(float(g["overlap"])))
    assert abs(pyci.compute_overlap(w1, w2, g["c1"], g["c2"]) - float(g["overlap"])) <= otol
    assert abs(pyci.compute_overlap(w2, w1, g["c2"], g["c1"]) - float(g["overlap"])) <= otol
    s1, s2 = pyci.compute_transition_rdms(w1, w1, g["c1"], g["c1"])
    q1, q2 = pyci.compute_rdms(w1, g["c1"])
    close(s1, q1)
    close(s2, q2)
    with pytest.raises(ValueError):
        pyci.compute_transition_rdms(w1, w2, g["c1"][:-1], g["c2"])


def test_transition_rdms_genci_equals_spinized_fullci(pyci):
    """GenCI (intended semantics) on the spin-orbital images equals the spin-expanded FullCI transition RDMs; a
    mismatched pair of spaces is refused."""
    n, occ = 6, (3, 2)
    full = pyci.fullci_wfn(n, *occ)
    full.add_all_dets()
    d = full.to_det_array()
    d1, d2 = np.ascontiguousarray(d[::2]), np.ascontiguousarray(d[::3])
    c1, c2 = seeded_vec(len(d1), 4), seeded_vec(len(d2), 6)
    f1, f2 = pyci.compute_transition_rdms(pyci.fullci_wfn(n, occ[0], occ[1], d1), pyci.fullci_wfn(n, occ[0], occ[1], d2), c1, c2)
    s1, s2 = pyci.spinize_rdms(f1, f2)
    img = lambda x: np.ascontiguousarray((x[:, 0] | (x[:, 1] << np.uint64(n))).reshape(-1, 1))  # noqa: E731
    g1, g2 = pyci.compute_transition_rdms(pyci.genci_wfn(2 * n, sum(occ), 0, img(d1)), pyci.genci_wfn(2 * n, sum(occ), 0, img(d2)), c1, c2)
    o1, o2 = O.compute_transition_rdms(O.GENCI, 2 * n, sum(occ), 0, img(d1), img(d2), c1, c2)
    close(g1, o1)
    close(g2, o2)
    close(g1, s1)
    close(g2[:n, :n, :n, :n], s2[:n, :n, :n, :n])
    close(g2[n:, n:, n:, n:], s2[n:, n:, n:, n:])
    with pytest.raises(ValueError):
        pyci.compute_transition_rdms(pyci.fullci_wfn(n, 3, 2, d1), pyci.fullci_wfn(n, 2, 2), c1, c2)


@pytest.mark.parametrize("kind,n,occ", [("fullci", 10, (3, 3)), ("genci", 16, (5, 0)), ("doci", 20, (4, 4)), ("fullci", 34, (2, 1))])
def test_bloom_filter_in_front_of_the_index(pyci, monkeypatch, kind, n, occ):
    """PYCI_B200_BLOOM=1 puts the blocked Bloom filter (built automatically once the slot table outgrows L2) in front
    of every index probe: operator, RDMs, add_hci and ENPT2 are unchanged; no determinant of the space is lost."""
    ecore, one, two = O.synthetic_integrals(n, 31)
    okind = KIND[kind]
    alld = O.all_dets(okind, n, *occ)
    dets = np.ascontiguousarray(alld[np.random.default_rng(2).permutation(len(alld))[:3000]])
    ham = pyci.hamiltonian(ecore, one, two)
    c = seeded_vec(len(dets), 12)
    c /= np.linalg.norm(c)
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("PYCI_B200_BLOOM", flag)
        wfn = getattr(pyci, kind + "_wfn")(n, occ[0], occ[1], dets)
        op = pyci.sparse_op(ham, wfn)
        r1, r2 = pyci.compute_rdms(wfn, c)
        pt = pyci.compute_enpt2(ham, wfn, c, -3.0, 2.0e-3)
        nadd = pyci.add_hci(ham, wfn, c, eps=2.0e-3)
        res[flag] = (op.indptr(), op.indices(), op.data(), r1, r2, pt, nadd, wfn.to_det_array()[len(dets):])
    a, b = res["0"], res["1"]
    assert all(np.array_equal(p, q) for p, q in zip(a[:3], b[:3]))
    close(b[3], a[3])
    close(b[4], a[4])
    assert abs(a[5] - b[5]) <= PT2_RTOL * abs(a[5]) and a[6] == b[6] and np.array_equal(a[7], b[7])
    ints = O.senzero_integrals(one, two) if kind == "doci" else (one, two)
    oi, ox, od = O.sparse_op(okind, n, occ[0], occ[1], dets, ints)
    assert np.array_equal(b[0], oi) and np.array_equal(b[1], ox) and np.array_equal(b[2], od)


def test_config3_full_size_properties(pyci):
    """BASELINE config 3 at its full size (FullCI 14 orbitals 4a4b: 1 002 001 determinants, 2.2e9 stored entries) through
    size-independent properties: entry counts from the combinatorial formulas, the first rows
    bit-exact against the oracle, symmetry and linearity of the product, the eigen-residual of the solve, and the
    traces / energy identity of the RDMs of that state."""
    from math import comb

    from pyci_b200 import cabi
    n, occ = 14, (4, 4)
    ecore, one, two = O.synthetic_integrals(n, 1234)
    ham = pyci.hamiltonian(ecore, one, two)
    wfn = pyci.fullci_wfn(n, *occ)
    wfn.add_all_dets()
    dets = wfn.to_det_array()
    ndet = len(wfn)
    assert ndet == comb(n, 4) ** 2 == 1002001
    a, va = 4, n - 4
    off = 2 * a * va + 2 * comb(a, 2) * comb(va, 2) + (a * va) ** 2  # connected determinants of every row
    op = pyci.sparse_op(ham, wfn)
    st = op.stats()
    assert st["stored_nnz"] == ndet * (off + 1) == 2225444221 and st["fill_kernel"] == "fill_complete_kernel"
    assert op.size == ndet * off // 2 + ndet == 1113223111  # lower triangle + diagonal
    ip = op.indptr()
    assert ip[0] == 0 and ip[-1] == op.size and np.all(np.diff(ip) >= 1) and ip.dtype == np.int64
    # the first rows, every column, bit for bit (rectangular slice through the C ABI)
    ctx = cabi.Context(0)
    dham = cabi.Ham(ctx, n, ecore, one, two)
    dwfn = cabi.Wfn(ctx, cabi.FULLCI, n, occ[0], occ[1], dets)
    k = 48
    head = cabi.Op(ctx, dham, dwfn, nrow=k, ncol=ndet, symmetric=False)
    hi, hx, hd = head.export_csr()
    oi, ox, od = O.sparse_op(O.FULLCI, n, occ[0], occ[1], dets, (one, two), nrow=k, ncol=ndet, symmetric=False)
    assert np.array_equal(hi, oi) and np.array_equal(hx, ox) and np.array_equal(hd, od)
    for i in (0, 7, k - 1):  # the symmetric operator holds the same lower-triangle elements
        for j in hx[hi[i]:hi[i + 1]]:
            if j <= i:
                assert op.get_element(i, int(j)) == hd[hi[i] + int(np.searchsorted(hx[hi[i]:hi[i + 1]], j))]
    head.close()
    dwfn.close()
    dham.close()
    ctx.close()
    # symmetry and linearity of y = A x
    x, y = seeded_vec(ndet, 21), seeded_vec(ndet, 22)
    ax, ay = op(x), op(y)
    assert abs(y @ ax - x @ ay) <= 1e-11 * abs(y @ ax)
    az = op(0.5 * x - 2.0 * y)
    assert np.max(np.abs(az - (0.5 * ax - 2.0 * ay))) <= 1e-11 * np.max(np.abs(az))
    # eigen-residual of the solve and the Rayleigh quotient
    es, cs = op.solve(n=1, tol=1e-9)
    c = cs[0]
    hc = op(c)
    e0 = es[0] - ecore
    assert abs(c @ c - 1.0) <= 1e-10 and abs(c @ hc - e0 * (c @ c)) <= 1e-9
    assert np.linalg.norm(hc - e0 * c) <= 1e-7 * abs(e0)
    assert e0 < float(np.min([op.get_element(i, i) for i in range(0, 40)]))  # variational: below the lowest diagonals
    # RDMs of that state: traces and the energy identity (test_routines.py:115-133)
    d1, d2 = pyci.compute_rdms(wfn, c)
    assert abs(np.trace(d1[0]) - occ[0]) <= 1e-10 and abs(np.trace(d1[1]) - occ[1]) <= 1e-10
    h2, g2 = O.spin_orbital_integrals(one, two)
    r1, r2 = pyci.spinize_rdms(d1, d2)
    energy = ecore + np.einsum("ij,ij", h2, r1) + 0.25 * (np.einsum("ijkl,ijkl", g2, r2) - np.einsum("ijlk,ijkl", g2, r2))
    assert abs(energy - es[0]) <= 1e-9


@pytest.mark.parametrize("kind,n,occ,count", [("doci", 66, (2, 2), None), ("doci", 130, (3, 3), 4000),
                                              ("genci", 70, (3, 0), 5000), ("fullci", 65, (2, 1), 6000),
                                              ("fullci", 129, (2, 2), 3000)])
def test_multiword_determinants(pyci, kind, n, occ, count):
    """nbasis > 64 (nword = 2 or 3; the reference builds these, common.cpp:280-282, and its wave-function tests use 65
    and 129 orbitals, pyci/test/test_wavefunction.py:45): the generic multi-word construction against the oracle --
    complete and selected spaces, symmetric / non-symmetric / rectangular, product and lowest eigenvalue."""
    okind = KIND[kind]
    ecore, one, two = O.synthetic_integrals(n, 13)
    ints = O.senzero_integrals(one, two) if kind == "doci" else (one, two)
    ham = pyci.hamiltonian(ecore, one, two)
    if count is None:
        wfn = getattr(pyci, kind + "_wfn")(n, *occ)
        wfn.add_all_dets()
        dets = wfn.to_det_array()
        assert np.array_equal(dets, O.all_dets(okind, n, *occ))
    else:
        # a selected space: seeded random occupations (distinct), through the occupation-array constructor
        rng = np.random.default_rng(3)
        seen, occs = set(), []
        while len(occs) < count:
            a = tuple(sorted(rng.choice(n, occ[0], replace=False)))
            b = tuple(sorted(rng.choice(n, occ[1], replace=False))) if kind == "fullci" else ()
            if (a, b) not in seen:
                seen.add((a, b))
                if kind == "fullci":
                    row = np.zeros((2, occ[0]), dtype=np.int64)
                    row[0] = a
                    row[1, :occ[1]] = b
                    occs.append(row)
                else:
                    occs.append(np.array(a, dtype=np.int64))
        wfn = getattr(pyci, kind + "_wfn")(n, occ[0], occ[1], np.ascontiguousarray(np.array(occs)))
        dets = wfn.to_det_array()
    assert dets.shape[-1] == (n + 63) // 64
    for kw in (dict(), dict(symmetric=False), dict(nrow=len(dets) - 3, ncol=len(dets) - 5, symmetric=False)):
        op = pyci.sparse_op(ham, wfn, **kw)
        oi, ox, od = O.sparse_op(okind, n, occ[0], occ[1], dets, ints, **kw)
        assert np.array_equal(op.indptr(), oi) and np.array_equal(op.indices(), ox), kw
        assert np.array_equal(op.data(), od), kw
        x = seeded_vec(op.shape[1], 5)
        yo = O.matvec(oi, ox, od, x, kw.get("symmetric", True))
        np.testing.assert_allclose(op(x), yo, rtol=0, atol=1e-12 * max(1.0, np.abs(yo).max()))
    op = pyci.sparse_op(ham, wfn)
    es, cs = op.solve(n=1, tol=1e-10)
    oi, ox, od = O.sparse_op(okind, n, occ[0], occ[1], dets, ints)
    e0, _ = O.lowest_eigenpair(oi, ox, od, len(dets))
    assert abs(es[0] - (e0 + ecore)) <= E_ATOL
