"""Pin the CPU oracle (oracle/pyci_oracle.c) against golden vectors produced by the reference's own
compiled sources (tests/golden/make_golden.py) and against the energies hard-coded in the reference's
tests (pyci/test/test_routines.py:41-45).  CPU only."""
import numpy as np
import pytest

from conftest import datafile, seeded_vec, sha
from oracle import oracle as O

SMALL = [("h4_sto3g", "fullci", (2, 2)), ("lih_sto6g", "fullci", (2, 2)), ("BH_sto-3g_eq", "fullci", (3, 3)),
         ("h6_sto_3g", "fullci", (4, 2)), ("be_ccpvdz", "doci", (2, 2)), ("h2_sto3g", "fullci", (1, 1))]
KIND = {"doci": O.DOCI, "fullci": O.FULLCI, "genci": O.GENCI}


def load(fn, kind):
    ecore, one, two = O.read_fcidump(datafile(fn))
    ints = O.senzero_integrals(one, two) if kind == "doci" else (one, two)
    return ecore, one.shape[0], ints


@pytest.mark.parametrize("fn,kind,occ", SMALL)
def test_oracle_small_csr_bit_exact(small, fn, kind, occ):
    ecore, n, ints = load(fn, kind)
    tag = f"{fn}.{kind}{occ[0]}{occ[1]}"
    dets = O.all_dets(KIND[kind], n, *occ)
    assert np.array_equal(dets, small[tag + ".dets"])
    x = seeded_vec(len(dets), 11)
    for name, kw in (("sym", dict(symmetric=True)), ("nonsym", dict(symmetric=False)),
                     ("rect", dict(nrow=len(dets) - 10, symmetric=False))):
        if tag + f".{name}.indptr" not in small:
            continue
        ip, ix, dv = O.sparse_op(KIND[kind], n, occ[0], occ[1], dets, ints, **kw)
        assert np.array_equal(ip, small[f"{tag}.{name}.indptr"])
        assert np.array_equal(ix, small[f"{tag}.{name}.indices"])
        assert np.array_equal(dv, small[f"{tag}.{name}.data"])  # bit-exact values
        y = O.matvec(ip, ix, dv, x, name == "sym")
        np.testing.assert_allclose(y, small[f"{tag}.{name}.y"], rtol=0, atol=1e-12)
    ip, ix, dv = O.sparse_op(KIND[kind], n, occ[0], occ[1], dets, ints)
    e0, _ = O.lowest_eigenpair(ip, ix, dv, len(dets))
    assert abs(e0 + ecore - float(small[tag + ".E0"])) < 1e-10


@pytest.mark.parametrize("fn,kind,occ", SMALL[:4] + SMALL[4:5])
def test_oracle_row_list_equals_reference_rows(small, fn, kind, occ):
    """oracle_sparse_op_rows (arbitrary rows: the reference's startrow loop, sparseop.cpp:186-201) reproduces
    the rows of the compiled reference's operators -- the entry the full-size sampled-row parity gates use."""
    _, n, ints = load(fn, kind)
    tag = f"{fn}.{kind}{occ[0]}{occ[1]}"
    dets = small[tag + ".dets"]
    nd = len(dets)
    rows = np.unique(np.concatenate([[0, nd - 1, nd // 2], np.random.default_rng(5).integers(0, nd, 40)]))[::-1]
    for name, sym in (("sym", True), ("nonsym", False)):
        if tag + f".{name}.indptr" not in small:
            continue
        rp, rx, rv = small[f"{tag}.{name}.indptr"], small[f"{tag}.{name}.indices"], small[f"{tag}.{name}.data"]
        ip, ix, dv = O.sparse_op(KIND[kind], n, occ[0], occ[1], dets, ints, symmetric=sym, rows=rows)
        assert len(ip) == len(rows) + 1
        for k, r in enumerate(rows):
            assert np.array_equal(ix[ip[k]:ip[k + 1]], rx[rp[r]:rp[r + 1]])
            assert np.array_equal(dv[ip[k]:ip[k + 1]], rv[rp[r]:rp[r + 1]])
    # restricted column range (rectangular operators) and an empty list
    ip, ix, dv = O.sparse_op(KIND[kind], n, occ[0], occ[1], dets, ints, symmetric=False, ncol=nd // 2, rows=rows)
    fp, fx, fv = O.sparse_op(KIND[kind], n, occ[0], occ[1], dets, ints, symmetric=False, ncol=nd // 2)
    for k, r in enumerate(rows):
        assert np.array_equal(ix[ip[k]:ip[k + 1]], fx[fp[r]:fp[r + 1]])
    ip, ix, dv = O.sparse_op(KIND[kind], n, occ[0], occ[1], dets, ints, rows=np.zeros(0, dtype=np.int64))
    assert ip.tolist() == [0] and len(ix) == 0


@pytest.mark.parametrize("fn,kind,occ", SMALL)
def test_oracle_small_rdms_bit_exact(small, fn, kind, occ):
    _, n, _ = load(fn, kind)
    tag = f"{fn}.{kind}{occ[0]}{occ[1]}"
    dets = small[tag + ".dets"]
    c = seeded_vec(len(dets), 12)
    c /= np.linalg.norm(c)
    r1, r2 = O.compute_rdms(KIND[kind], n, occ[0], occ[1], dets, c)
    assert np.array_equal(r1, small[tag + ".rdm1"])
    assert np.array_equal(r2, small[tag + ".rdm2"])


def test_oracle_selected_space(small):
    _, one, two = O.synthetic_integrals(8, 1234)
    dets = small["syn8.fullci32.sel.dets"]
    for name, sym in (("sym", True), ("nonsym", False)):
        ip, ix, dv = O.sparse_op(O.FULLCI, 8, 3, 2, dets, (one, two), symmetric=sym)
        assert np.array_equal(ip, small[f"syn8.fullci32.sel.{name}.indptr"])
        assert np.array_equal(ix, small[f"syn8.fullci32.sel.{name}.indices"])
        assert np.array_equal(dv, small[f"syn8.fullci32.sel.{name}.data"])
    c = seeded_vec(len(dets), 12)
    c /= np.linalg.norm(c)
    r1, r2 = O.compute_rdms(O.FULLCI, 8, 3, 2, dets, c)
    assert np.array_equal(r1, small["syn8.fullci32.sel.rdm1"])
    assert np.array_equal(r2, small["syn8.fullci32.sel.rdm2"])


@pytest.mark.parametrize("fn,kind,occ,pinned", [
    ("be_ccpvdz", "fullci", (2, 2), -14.617409507),   # BASELINE config 1, test_routines.py:44
    ("h2o_ccpvdz", "doci", (5, 5), -75.634588422),    # BASELINE config 2, test_routines.py:45
    ("li2_ccpvdz", "doci", (3, 3), -14.878455349),    # test_routines.py:41
])
def test_oracle_configs_digest_and_energy(digests, fn, kind, occ, pinned):
    ecore, n, ints = load(fn, kind)
    tag = f"{fn}.{kind}{occ[0]}{occ[1]}"
    dets = O.all_dets(KIND[kind], n, *occ)
    g = digests[tag + ".sym"]
    assert len(dets) == g["ndet"]
    ip, ix, dv = O.sparse_op(KIND[kind], n, occ[0], occ[1], dets, ints)
    assert (len(ix), sha(ip), sha(ix), sha(dv)) == (g["nnz"], g["indptr"], g["indices"], g["data"])
    e0, c = O.lowest_eigenpair(ip, ix, dv, len(dets))
    assert abs(e0 + ecore - g["E0"]) < 1e-10
    assert abs(e0 + ecore - pinned) < 1e-9       # the reference's own tolerance
    y = O.matvec(ip, ix, dv, seeded_vec(len(dets), 11), True)
    assert abs(np.linalg.norm(y) - g["y_norm"]) < 1e-9 * g["y_norm"]
    gn = digests[tag + ".nonsym"]
    ip, ix, dv = O.sparse_op(KIND[kind], n, occ[0], occ[1], dets, ints, symmetric=False)
    assert (len(ix), sha(ip), sha(ix), sha(dv)) == (gn["nnz"], gn["indptr"], gn["indices"], gn["data"])
    cc = seeded_vec(len(dets), 12)
    cc /= np.linalg.norm(cc)
    r1, r2 = O.compute_rdms(KIND[kind], n, occ[0], occ[1], dets, cc)
    assert (sha(r1), sha(r2)) == (digests[tag + ".rdm"]["rdm1"], digests[tag + ".rdm"]["rdm2"])


def test_oracle_multiword_and_unequal_spin(digests):
    _, one, two = O.synthetic_integrals(66, 7)
    dets = O.all_dets(O.DOCI, 66, 2)
    ip, ix, dv = O.sparse_op(O.DOCI, 66, 2, 2, dets, O.senzero_integrals(one, two))
    g = digests["syn66.doci22.sym"]
    assert (len(ix), sha(ip), sha(ix), sha(dv)) == (g["nnz"], g["indptr"], g["indices"], g["data"])
    _, one, two = O.synthetic_integrals(9, 1234)
    dets = O.all_dets(O.FULLCI, 9, 4, 3)
    ip, ix, dv = O.sparse_op(O.FULLCI, 9, 4, 3, dets, (one, two))
    g = digests["syn9.fullci43.sym"]
    assert (len(ix), sha(ip), sha(ix), sha(dv)) == (g["nnz"], g["indptr"], g["indices"], g["data"])


@pytest.mark.parametrize("fn,occ", [("h4_sto3g", (2, 2)), ("BH_sto-3g_eq", (3, 3)), ("h6_sto_3g", (4, 2))])
def test_oracle_genci(genci_golden, fn, occ):
    """GenCI: (i) equals the reference compiled with the two loop bounds fixed, (ii) equals FullCI on the
    spatial integrals bit-for-bit, (iii) RDMs equal the spin-expanded FullCI RDMs."""
    ecore, one, two = O.read_fcidump(datafile(fn))
    n = one.shape[0]
    fd = O.all_dets(O.FULLCI, n, *occ)
    gd = (fd[:, 0, :] | (fd[:, 1, :] << np.uint64(n))).astype(np.uint64)
    tag = f"{fn}.genci{sum(occ)}"
    assert np.array_equal(gd, genci_golden[tag + ".dets"])
    h2, g2 = O.spin_orbital_integrals(one, two)
    for name, sym in (("sym", True), ("nonsym", False)):
        got = O.sparse_op(O.GENCI, 2 * n, sum(occ), 0, gd, (h2, g2), symmetric=sym)
        for a, key in zip(got, ("indptr", "indices", "data")):
            assert np.array_equal(a, genci_golden[f"{tag}.{name}.{key}"])
        full = O.sparse_op(O.FULLCI, n, occ[0], occ[1], fd, (one, two), symmetric=sym)
        for a, b in zip(got, full):
            assert np.array_equal(a, b)
    c = seeded_vec(len(fd), 2)
    c /= np.linalg.norm(c)
    s1, s2 = O.spinize_rdms(*O.compute_rdms(O.FULLCI, n, occ[0], occ[1], fd, c))
    g1, g2r = O.compute_rdms(O.GENCI, 2 * n, sum(occ), 0, gd, c)
    np.testing.assert_allclose(g1, s1, rtol=0, atol=1e-14)
    np.testing.assert_allclose(g2r, s2, rtol=0, atol=1e-14)


def test_oracle_rdm_energy_identity():
    """E = ecore + sum h g1 + 1/4 sum <pq||rs> G2 (test_routines.py:130-133) on Be/cc-pVDZ DOCI and H4 FullCI."""
    for fn, kind, occ in (("be_ccpvdz", "doci", (2, 2)), ("h4_sto3g", "fullci", (2, 2))):
        ecore, n, ints = load(fn, kind)
        _, one, two = O.read_fcidump(datafile(fn))
        dets = O.all_dets(KIND[kind], n, *occ)
        ip, ix, dv = O.sparse_op(KIND[kind], n, occ[0], occ[1], dets, ints)
        e0, c = O.lowest_eigenpair(ip, ix, dv, len(dets))
        r1, r2 = O.spinize_rdms(*O.compute_rdms(KIND[kind], n, occ[0], occ[1], dets, c))
        h2, g2 = O.spin_orbital_integrals(one, two)
        anti = g2 - g2.transpose(0, 1, 3, 2)
        e = ecore + np.einsum("ij,ij", h2, r1) + 0.25 * np.einsum("ijkl,ijkl", anti, r2)
        assert abs(e - (e0 + ecore)) < 1e-9


# ---- add_hci / compute_enpt2 (hci.cpp, enpt2.cpp) ----------------------------------------------------
from conftest import HCI_CASES, sorted_rows  # noqa: E402


@pytest.mark.parametrize("tag,fn,kind,occ,steps", HCI_CASES)
def test_oracle_hci_enpt2_against_reference(hci_golden, tag, fn, kind, occ, steps):
    """The C restatement selects the same determinant SET as the compiled reference's add_hci (its append order
    is hash-map iteration order, unspecified) and reproduces its ENPT2 energies."""
    ecore, one, two = O.read_fcidump(datafile(fn))
    n = one.shape[0]
    ints = O.senzero_integrals(one, two) if kind == "doci" else (one, two)
    for it in range(steps):
        g = {k: hci_golden[f"{tag}.{it}.{k}"] for k in ("dets", "coeffs", "energy", "eps", "enpt2", "enpt2_tight", "new_sorted")}
        eps, e = float(g["eps"]), float(g["energy"])
        new = O.add_hci(KIND[kind], n, occ[0], occ[1], g["dets"], ints, g["coeffs"], eps)
        assert np.array_equal(sorted_rows(new), g["new_sorted"])
        assert len(new) == 0 or len(np.unique(new.reshape(len(new), -1), axis=0)) == len(new)
        for key, ee in (("enpt2", eps), ("enpt2_tight", eps * 1e-2)):
            pt, _ = O.compute_enpt2(KIND[kind], n, occ[0], occ[1], g["dets"], (one, two), g["coeffs"], e, ecore, ee)
            assert abs(pt - float(g[key])) <= 1e-13 * abs(float(g[key]))


def test_oracle_hci_genci_equals_fullci():
    """GenCI (loops bounded by nvir_up) on spin-orbital integrals selects the images of the FullCI selection and
    gives the same ENPT2 energy."""
    ecore, one, two = O.read_fcidump(datafile("h6_sto_3g"))
    n, occ = one.shape[0], (3, 3)
    h2, g2 = O.spin_orbital_integrals(one, two)
    fd = O.all_dets(O.FULLCI, n, *occ)[::7]
    c = seeded_vec(len(fd), 5)
    c /= np.linalg.norm(c)
    gd = (fd[:, 0] | (fd[:, 1] << np.uint64(n))).reshape(-1, 1)
    newf = O.add_hci(O.FULLCI, n, occ[0], occ[1], fd, (one, two), c, 1.0e-3)
    newg = O.add_hci(O.GENCI, 2 * n, sum(occ), 0, gd, (h2, g2), c, 1.0e-3)
    assert len(newf) > 0
    assert np.array_equal(sorted_rows((newf[:, 0] | (newf[:, 1] << np.uint64(n))).reshape(-1, 1)), sorted_rows(newg))
    ptf, ntf = O.compute_enpt2(O.FULLCI, n, occ[0], occ[1], fd, (one, two), c, -3.0, ecore, 1.0e-4)
    ptg, ntg = O.compute_enpt2(O.GENCI, 2 * n, sum(occ), 0, gd, (h2, g2), c, -3.0, ecore, 1.0e-4)
    assert ntf == ntg and abs(ptf - ptg) <= 1e-12 * abs(ptf)


# ---- compute_transition_rdms / compute_overlap (rdm.cpp:634-1009, overlap.cpp) ---------------------------
from conftest import TRDM_CASES  # noqa: E402


@pytest.mark.parametrize("tag,kind,n,occ", TRDM_CASES)
def test_oracle_transition_rdms_bit_exact(trdm_golden, tag, kind, n, occ):
    g = {k: trdm_golden[f"{tag}.{k}"] for k in ("dets1", "dets2", "c1", "c2", "rdm1", "rdm2", "overlap")}
    r1, r2 = O.compute_transition_rdms(KIND[kind], n, occ[0], occ[1], g["dets1"], g["dets2"], g["c1"], g["c2"])
    assert np.array_equal(r1, g["rdm1"]) and np.array_equal(r2, g["rdm2"])
    assert O.compute_overlap(g["dets1"], g["dets2"], g["c1"], g["c2"]) == float(g["overlap"])
    # T(wfn, wfn, c, c) is the symmetric routine
    s1, s2 = O.compute_transition_rdms(KIND[kind], n, occ[0], occ[1], g["dets1"], g["dets1"], g["c1"], g["c1"])
    q1, q2 = O.compute_rdms(KIND[kind], n, occ[0], occ[1], g["dets1"], g["c1"])
    np.testing.assert_allclose(s1, q1, rtol=0, atol=1e-13)
    np.testing.assert_allclose(s2, q2, rtol=0, atol=1e-13)


def test_oracle_hci_enpt2_limits():
    """Limits that hold for any implementation: a huge eps selects nothing, eps = 0 selects every connected external
    determinant, a complete space has no external space (ENPT2 = E), zero coefficients select nothing."""
    ecore, one, two = O.read_fcidump(datafile("h6_sto_3g"))
    n, occ = one.shape[0], (3, 3)
    full = O.all_dets(O.FULLCI, n, *occ)
    c_full = seeded_vec(len(full), 1)
    assert len(O.add_hci(O.FULLCI, n, occ[0], occ[1], full, (one, two), c_full, 0.0)) == 0
    pt, nt = O.compute_enpt2(O.FULLCI, n, occ[0], occ[1], full, (one, two), c_full, -3.25, ecore, 0.0)
    assert nt == 0 and pt == -3.25
    sel = full[::5]
    c = seeded_vec(len(sel), 2)
    assert len(O.add_hci(O.FULLCI, n, occ[0], occ[1], sel, (one, two), c, 1.0e9)) == 0
    assert len(O.add_hci(O.FULLCI, n, occ[0], occ[1], sel, (one, two), np.zeros(len(sel)), 1.0e-9)) == 0
    # eps = 0 keeps every excitation with a non-zero element: a subset of the complement that contains every
    # determinant the build would connect to the selection
    everything = O.add_hci(O.FULLCI, n, occ[0], occ[1], sel, (one, two), c, 0.0)
    have = {tuple(r) for r in sel.reshape(len(sel), -1).tolist()}
    new = {tuple(r) for r in everything.reshape(len(everything), -1).tolist()}
    assert not (new & have) and len(new) == len(everything)
    ip, ix, dv = O.sparse_op(O.FULLCI, n, occ[0], occ[1], full, (one, two), symmetric=False)
    pos = {tuple(r): i for i, r in enumerate(full.reshape(len(full), -1).tolist())}
    rows = [pos[t] for t in have]
    connected = set()
    for r in rows:
        for j, v in zip(ix[ip[r]:ip[r + 1]], dv[ip[r]:ip[r + 1]]):
            if v != 0.0 and tuple(full[j].reshape(-1).tolist()) not in have:
                connected.add(tuple(full[j].reshape(-1).tolist()))
    assert new == connected


from conftest import UPDATE_CASES, UPDATE_MODES  # noqa: E402


@pytest.mark.parametrize("tag,fn,kind,occ", UPDATE_CASES)
@pytest.mark.parametrize("mode,symm", UPDATE_MODES)
def test_oracle_update_equals_reference(update_golden, tag, fn, kind, occ, mode, symm):
    """SparseOp::update (sparseop.cpp:175-201) as the compiled reference applies it, twice in a row: appended rows see
    the grown wave function, rows the operator already has are left alone -- for a non-symmetric operator they keep
    the columns they were built with (fewer entries than a fresh build has)."""
    g, key = update_golden, "%s.%s" % (tag, mode)
    ecore, one, two = O.read_fcidump(datafile(fn))
    ints = O.senzero_integrals(one, two) if kind == "doci" else (one, two)
    okind = {"doci": O.DOCI, "fullci": O.FULLCI}[kind]
    dets, sizes, nrow0 = g[key + ".dets"], g[key + ".sizes"].tolist(), int(g[key + ".nrow0"])
    ip, ix, dv = O.sparse_op_updated(okind, one.shape[0], occ[0], occ[1], dets, ints, nrow0, sizes, symm)
    assert tuple(g[key + ".shape"]) == (sizes[-1], sizes[-1]) and len(ix) == int(g[key + ".nnz"])
    assert np.array_equal(ip, g[key + ".indptr"])
    assert sha(ix) == str(g[key + ".indices.sha256"]) and sha(dv) == str(g[key + ".data.sha256"])
    if not symm:  # the point of the non-symmetric cases: not what a fresh build gives
        fresh = O.sparse_op(okind, one.shape[0], occ[0], occ[1], dets, ints, symmetric=False)
        assert len(fresh[1]) > len(ix)



def test_direct_ci_generator_of_the_cfg4_energy():
    """tests/golden/make_golden_e0_direct.py (the string-driven product behind the cfg4 entry of e0_syn.json): its product
    equals the oracle's FullCI matrix on a random vector, and its E0 of syn8 equals the matrix-based golden."""
    import importlib.util
    import json
    import os
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location("make_golden_e0_direct", os.path.join(GOLDEN, "make_golden_e0_direct.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert m.self_test(6) < 1e-13
    with open(os.path.join(GOLDEN, "e0_syn.json")) as f:
        gold = json.load(f)
    r = m.lowest(8)
    assert abs(r["E0"] - gold["syn8"]["E0"]) < 1e-11 and r["ndet"] == gold["syn8"]["ndet"]
    assert abs(gold["syn14"]["direct_ci_E0"] - gold["syn14"]["E0"]) < 1e-11  # recorded when the cfg4 entry was made
    assert gold["syn16"]["operator"].startswith("string-driven") and gold["syn16"]["ndet"] == 3312400


@pytest.mark.parametrize("K,P,nd", [(8, 3, 56), (10, 3, 100), (16, 5, 3000)])
def test_config5_style_genci_operator_equals_the_reference_doci_operator(K, P, nd):
    """Config 5's spaces are seniority-zero selections inside a GenCI spin-orbital space.  The reference's GenCI kernels
    are defective, its DOCI kernels are not: between closed-shell determinants only pair
    excitations survive, with element <kk|ll> = v[k,l], and the diagonal is the DOCI diagonal.  So the oracle's GenCI
    operator over spin-orbital integrals must equal the COMPILED REFERENCE's DOCI operator over the pair strings: same
    structure bit for bit, data to rounding (the sums run in a different order) -- a pin of the config-5 parity chain
    (device == oracle bit for bit) on the reference's own code."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref"))
    R = pytest.importorskip("pyci_ref")
    from pyci_b200.synthetic import seniority_zero_genci_dets
    ecore, one, two = O.synthetic_integrals(K, 77)
    hs, gs = O.spin_orbital_integrals(one, two)
    dets = seniority_zero_genci_dets(K, P, nd)
    pair = (dets[:, 0] & np.uint64((1 << K) - 1)).reshape(-1, 1).copy()
    assert np.array_equal(dets[:, 0], pair[:, 0] | (pair[:, 0] << np.uint64(K)))
    for sym in (True, False):
        ip, ix, dv = O.sparse_op(O.GENCI, 2 * K, 2 * P, 0, dets, (hs, gs), symmetric=sym)
        op = R.sparse_op(R.secondquant_op(ecore, one, two), R.doci_wfn(K, P, P, pair), symmetric=sym)
        assert np.array_equal(ip, op.indptr()) and np.array_equal(ix, op.indices())
        ref = op.data()
        assert np.max(np.abs(dv - ref)) <= 1e-13 * np.max(np.abs(ref))
