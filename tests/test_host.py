"""Host side of the boundary (pyci_b200._pyci, C++/pybind11) against the golden vectors made from the
compiled reference and against the reference's own wave-function / Hamiltonian tests
(pyci/test/test_wavefunction.py:28-184, test_hamiltonian.py:26-39).  No GPU needed."""
import numpy as np
import pytest

import pyci_b200 as pyci
from conftest import datafile
from oracle import oracle as O

SMALL = [("h4_sto3g", "fullci", (2, 2)), ("lih_sto6g", "fullci", (2, 2)), ("BH_sto-3g_eq", "fullci", (3, 3)),
         ("h6_sto_3g", "fullci", (4, 2)), ("be_ccpvdz", "doci", (2, 2)), ("h2_sto3g", "fullci", (1, 1))]


@pytest.mark.parametrize("fn,kind,occ", SMALL)
def test_add_all_dets_order_matches_reference(small, fn, kind, occ):
    ham = pyci.hamiltonian(datafile(fn))
    wfn = getattr(pyci, kind + "_wfn")(ham.nbasis, *occ)
    wfn.add_all_dets()
    ref = small[f"{fn}.{kind}{occ[0]}{occ[1]}.dets"]
    got = wfn.to_det_array()
    assert got.dtype == np.uint64 and got.shape == ref.shape
    assert np.array_equal(got, ref)
    # index_det is the inverse of the ordering
    for i in (0, len(wfn) // 2, len(wfn) - 1):
        assert wfn.index_det(wfn[i]) == i


@pytest.mark.parametrize("fn", ["be_ccpvdz", "h2o_ccpvdz", "li2_ccpvdz", "h4_sto3g"])
def test_fcidump_reader_matches_oracle_reader(fn):
    ham = pyci.hamiltonian(datafile(fn))
    ecore, one, two = O.read_fcidump(datafile(fn))
    assert ham.ecore == ecore
    assert np.array_equal(ham.one_mo, one) and np.array_equal(ham.two_mo, two)
    h, v, w = O.senzero_integrals(one, two)
    assert np.array_equal(ham.h, h) and np.array_equal(ham.v, v) and np.array_equal(ham.w, w)


def test_fcidump_round_trip(tmp_path):
    ham = pyci.hamiltonian(datafile("be_ccpvdz"))
    out = str(tmp_path / "x.fcidump")
    ham.to_file(out, nelec=4, ms2=0)
    back = pyci.hamiltonian(out)
    assert back.nbasis == ham.nbasis
    np.testing.assert_allclose(back.ecore, ham.ecore, rtol=0, atol=1e-12)
    np.testing.assert_allclose(back.one_mo, ham.one_mo, rtol=0, atol=1e-12)
    np.testing.assert_allclose(back.two_mo, ham.two_mo, rtol=0, atol=1e-12)


@pytest.mark.parametrize("cls,args", [(pyci.doci_wfn, (10, 11, 11)), (pyci.doci_wfn, (10, 5, 4)),
                                      (pyci.fullci_wfn, (10, 11, 2)), (pyci.fullci_wfn, (10, 2, 3)),
                                      (pyci.genci_wfn, (10, 11, 0)), (pyci.genci_wfn, (10, 4, 1))])
def test_bad_occupations_raise_value_error(cls, args):
    with pytest.raises(ValueError):
        cls(*args)


@pytest.mark.parametrize("nbasis", [16, 64, 65, 129])
def test_wfn_file_and_array_round_trips(tmp_path, nbasis):
    for cls, occ in ((pyci.doci_wfn, (3, 3)), (pyci.fullci_wfn, (2, 1)), (pyci.genci_wfn, (3, 0))):
        wfn = cls(nbasis, *occ)
        wfn.add_hartreefock_det()
        wfn.add_excited_dets(1)
        n = len(wfn)
        assert n > 1
        path = str(tmp_path / ("w%d.bin" % nbasis))
        wfn.to_file(path)
        back = cls(path)
        assert len(back) == n and np.array_equal(back.to_det_array(), wfn.to_det_array())
        again = cls(nbasis, occ[0], occ[1], wfn.to_det_array())
        assert np.array_equal(again.to_det_array(), wfn.to_det_array())
        occs = cls(nbasis, occ[0], occ[1], wfn.to_occ_array())
        assert np.array_equal(occs.to_det_array(), wfn.to_det_array())
        assert wfn.index_det(wfn[n - 1]) == n - 1
        assert wfn.add_det(wfn[0]) == -1  # already present


def test_add_all_dets_counts():
    from math import comb
    w = pyci.doci_wfn(12, 4, 4)
    w.add_all_dets()
    assert len(w) == comb(12, 4)
    w = pyci.fullci_wfn(8, 3, 2)
    w.add_all_dets()
    assert len(w) == comb(8, 3) * comb(8, 2)
    w = pyci.genci_wfn(9, 3, 0)
    w.add_all_dets()
    assert len(w) == comb(9, 3)


def test_excitation_levels_and_hartree_fock():
    w = pyci.fullci_wfn(6, 2, 2)
    w.add_excited_dets(0)
    assert len(w) == 1 and w.to_occ_array().tolist() == [[[0, 1], [0, 1]]]
    w.add_excited_dets(1)
    assert len(w) == 1 + 2 * 2 * 4
    pyci.add_excitations(w, 2)
    assert len(w) == 1 + 16 + (1 * 6 * 2 + 8 * 8)


def test_bit_helpers_and_threads():
    assert pyci.popcnt(np.array([0b1011, 0b1], dtype=np.uint64)) == 4
    assert pyci.ctz(np.array([0b1000], dtype=np.uint64)) == 3
    old = pyci.get_num_threads()
    pyci.set_num_threads(2)
    assert pyci.get_num_threads() == 2
    pyci.set_num_threads(old)


def test_synthetic_integrals_match_oracle_copy():
    from pyci_b200.synthetic import spin_orbital_integrals, synthetic_integrals
    for n in (4, 9):
        a, b = synthetic_integrals(n, 1234), O.synthetic_integrals(n, 1234)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        ha, ga = spin_orbital_integrals(a[1], a[2])
        hb, gb = O.spin_orbital_integrals(b[1], b[2])
        assert np.array_equal(ha, hb) and np.array_equal(ga, gb)
        # 8-fold symmetry of real orbitals in physicist order
        g = a[2]
        assert np.array_equal(g, g.transpose(1, 0, 3, 2)) and np.array_equal(g, g.transpose(2, 1, 0, 3))


def test_utility_helpers_match_oracle_restatement():
    rng = np.random.default_rng(3)
    n = 5
    d1, d2 = rng.standard_normal((2, n, n)), rng.standard_normal((3, n, n, n, n))
    a, b = pyci.spinize_rdms(d1, d2), O.spinize_rdms(d1, d2)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    d0, dd = rng.standard_normal((n, n)), rng.standard_normal((n, n))
    a, b = pyci.spinize_rdms(d0, dd), O.spinize_rdms(d0, dd)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    one, two = rng.standard_normal((n, n)), rng.standard_normal((n, n, n, n))
    for x, y in zip(pyci.make_senzero_integrals(one, two), O.senzero_integrals(one, two)):
        assert np.array_equal(x, y)
    # spin_free_rdms (utility.py:406-421): spin blocks summed; checked equal to the reference's function when written
    s1, s2 = pyci.spin_free_rdms(d1, d2)
    g1, g2 = O.spinize_rdms(d1, d2)
    u, v = slice(0, n), slice(n, 2 * n)
    assert np.array_equal(s1, g1[u, u] + g1[v, v])
    assert np.array_equal(s2, g2[u, u, u, u] + g2[u, v, u, v] + g2[v, u, v, u] + g2[v, v, v, v])
    assert abs(np.einsum("pqpq", s2) - sum(np.einsum("pqpq", g2[x, y, x, y]) for x in (u, v) for y in (u, v))) < 1e-12
    with pytest.raises(NotImplementedError):
        pyci.spin_free_rdms(d0, dd)


def test_bulk_append_of_new_determinants():
    """Wfn::append_new_dets, the host side of add_hci: determinants selected on the device (distinct, absent) are
    appended without per-determinant look-ups; the dictionary then finds every determinant at its position and
    still rejects duplicates."""
    import numpy as np

    import pyci_b200 as pyci
    for cls, args, nw in ((pyci.fullci_wfn, (10, 3, 3), 2), (pyci.doci_wfn, (16, 4, 4), 1), (pyci.genci_wfn, (12, 5, 0), 1)):
        full = cls(*args)
        full.add_all_dets()
        d = full.to_det_array()
        d = np.ascontiguousarray(d[np.random.default_rng(3).permutation(len(d))])
        for cut in (0, 1, len(d) // 3):
            w = cls(*args, d[:cut]) if cut else cls(*args)
            w._append_new_dets(d[cut:])
            assert len(w) == len(d) and np.array_equal(w.to_det_array(), d)
            for i in (0, cut, len(d) // 2, len(d) - 1):
                assert w.index_det(d[i]) == i
            assert w.add_det(d[len(d) // 2]) == -1 and len(w) == len(d)
            w._append_new_dets(d[:0])  # empty append is a no-op
            assert len(w) == len(d)
        if nw > 1:  # a ragged array (not a whole number of determinants) is refused
            with pytest.raises(ValueError):
                cls(*args)._append_new_dets(np.zeros(2 * nw + 1, dtype=np.uint64))


def _selector_cases():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "make_golden_selectors", os.path.join(os.path.dirname(__file__), "golden", "make_golden_selectors.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_selectors_match_the_reference_python_layer():
    """add_seniorities / add_cost (odometers) / add_gkci against tests/golden/selectors.npz: determinant arrays made by the
    reference's own Python functions over its compiled classes, insertion order included."""
    import os
    G = _selector_cases()
    with np.load(os.path.join(os.path.dirname(__file__), "golden", "selectors.npz")) as f:
        gold = {k: f[k] for k in f.files}

    def same(w, tag):
        return len(w) == len(gold[tag]) and (len(w) == 0 or np.array_equal(w.to_det_array(), gold[tag]))

    for tag, shape, sens in G.SENIORITY:
        w = pyci.fullci_wfn(*shape)
        pyci.add_seniorities(w, *sens)
        assert same(w, tag), tag
    for tag, cls, shape, t, qmax in G.ODOMETER:
        w = getattr(pyci, cls)(*shape)
        pyci.add_cost(w, G.costs(shape[0], 7), qmax, t)
        assert same(w, tag), tag
        w2 = getattr(pyci, cls)(*shape)
        odo = pyci.odometer_two_spin if cls == "fullci_wfn" else pyci.utility.odometer_one_spin
        odo(w2, G.costs(shape[0], 7), t, qmax)
        assert same(w2, tag), tag
    for tag, cls, shape, kw in G.GKCI:
        w = getattr(pyci, cls)(*shape)
        kw = dict(kw)
        if kw.get("mode") == "interval":
            kw["energies"] = G.costs(shape[0] + 1, 11)
        if kw.get("mode") == "nodes":
            kw["mode"] = G.costs(shape[0] + 1, 13)
        pyci.add_gkci(w, **kw)
        assert same(w, tag), tag


def test_selector_argument_errors():
    """seniority_ci.py:43-52, gkci.py:63-67: wrong wave-function type -> TypeError, impossible seniority or unknown node
    model -> ValueError."""
    with pytest.raises(TypeError):
        pyci.add_seniorities(pyci.doci_wfn(6, 2, 2), 0)
    w = pyci.fullci_wfn(6, 3, 1)
    for bad in (0, 3, 6):          # below nocc_up - nocc_dn, wrong parity, above the bound
        with pytest.raises(ValueError):
            pyci.add_seniorities(w, bad)
    assert len(w) == 0
    with pytest.raises(ValueError):
        pyci.add_gkci(pyci.doci_wfn(6, 2, 2), mode="nope")
    # seniority sectors partition the full space
    full = pyci.fullci_wfn(6, 3, 1)
    full.add_all_dets()
    pyci.add_seniorities(w, 2, 4)
    assert len(w) == len(full)
