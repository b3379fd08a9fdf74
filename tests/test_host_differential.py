"""Differential tests of the host boundary: random call sequences on pyci_b200's wave-function classes and on the
COMPILED REFERENCE's (oracle/_ref/pyci_ref), same results and same exception types; random selector calls against the
reference's own Python functions (read from /root/reference where it exists, build container only).  CPU only.

Left out on purpose, because the reference itself misbehaves there ("host classes"):
`wfn[i]` out of range (the reference reads past its array; here IndexError), conversions to genci_wfn (SURVEY section 0
fact 9), `to_det_array(low > 0, high)` of two-spin wave functions (twospinwfn.cpp:112 offsets by `low * nword` instead of
`low * nword2`), the unused down-spin tail of `to_occ_array` rows when nocc_dn < nocc_up (uninitialised in the reference,
zero here), and the ORDER in which `add_dets_from_wfn` appends (the reference walks its hash map, onespinwfn.cpp:246-249;
the set is compared)."""
import os
import sys

import numpy as np
import pytest

import pyci_b200 as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
try:
    import pyci_ref as R
except ImportError:  # the compiled reference is built where /root/reference exists and travels with the snapshot
    R = None
needs_ref = pytest.mark.skipif(R is None, reason="oracle/_ref/pyci_ref not built")


class Pair:
    """runs the same call on both implementations and compares results / exception types"""

    def __init__(self):
        self.n = 0

    def __call__(self, tag, fa, fb):
        self.n += 1
        ra = rb = ea = eb = None
        try:
            ra = fa()
        except Exception as e:  # noqa: BLE001 - the type is the thing compared
            ea = type(e).__name__
        try:
            rb = fb()
        except Exception as e:  # noqa: BLE001
            eb = type(e).__name__
        assert ea == eb, (tag, ea, eb)
        if ea is None:
            if isinstance(ra, np.ndarray):
                assert ra.shape == rb.shape and ra.dtype == rb.dtype and np.array_equal(ra, rb), tag
            else:
                assert ra == rb, (tag, ra, rb)


def rows_as_set(d):
    return sorted(map(bytes, d.reshape(len(d), -1)))


@needs_ref
def test_wavefunction_classes_against_the_compiled_reference():
    rng = np.random.default_rng(5)
    cmp = Pair()
    for _ in range(120):
        nb = int(rng.choice([3, 5, 8, 20, 63, 64, 65, 100, 129]))
        up = int(rng.integers(1, min(nb, 4) + 1))
        dn = int(rng.integers(0, up + 1))
        kind = str(rng.choice(["doci", "genci", "fullci"]))
        args = (nb, up, up) if kind == "doci" else (nb, up, 0) if kind == "genci" else (nb, up, dn)
        a, b = getattr(R, kind + "_wfn")(*args), getattr(M, kind + "_wfn")(*args)
        for at in ("nbasis", "nocc", "nocc_up", "nocc_dn", "nvir", "nvir_up", "nvir_dn"):
            cmp(at, lambda: getattr(a, at), lambda: getattr(b, at))
        for _k in range(int(rng.integers(0, 12))):
            if kind == "fullci":
                occ = np.zeros((2, up), dtype=np.int64)
                occ[0] = np.sort(rng.choice(nb, up, replace=False))
                if dn:
                    occ[1, :dn] = np.sort(rng.choice(nb, dn, replace=False))
            else:
                occ = np.sort(rng.choice(nb, up, replace=False)).astype(np.int64)
            cmp("add_occs", lambda: a.add_occs(occ), lambda: b.add_occs(occ))
        cmp("hf", lambda: a.add_hartreefock_det(), lambda: b.add_hartreefock_det())
        if nb <= 20:
            e = int(rng.integers(0, 3))
            cmp("exc", lambda: a.add_excited_dets(e), lambda: b.add_excited_dets(e))
            if len(a) > 2:
                r = a[len(a) // 2]
                cmp("exc_ref", lambda: a.add_excited_dets(1, r), lambda: b.add_excited_dets(1, r))
        cmp("len", lambda: len(a), lambda: len(b))
        cmp("dets", lambda: a.to_det_array(), lambda: b.to_det_array())
        used = (lambda x: x) if kind != "fullci" else (lambda x: np.concatenate([x[:, 0].ravel(), x[:, 1, :dn].ravel()]))
        cmp("occs", lambda: used(a.to_occ_array()), lambda: used(b.to_occ_array()))
        n = len(a)
        lo, hi = int(rng.integers(-1, n + 1)), int(rng.integers(-1, n + 1))
        if kind == "fullci" and lo > 0:
            lo = 0
        if not lo > hi >= 0:
            cmp("dets_slice", lambda: a.to_det_array(lo, hi), lambda: b.to_det_array(lo, hi))
            cmp("occs_slice", lambda: used(a.to_occ_array(lo, hi)), lambda: used(b.to_occ_array(lo, hi)))
        if n:
            i = int(rng.integers(0, n))
            d = a[i]
            cmp("getitem", lambda: a[i], lambda: b[i])
            cmp("index_det", lambda: a.index_det(d), lambda: b.index_det(d))
            cmp("rank_det", lambda: a.rank_det(d), lambda: b.rank_det(d))
            cmp("index_det_from_rank", lambda: a.index_det_from_rank(a.rank_det(d)), lambda: b.index_det_from_rank(b.rank_det(d)))
            cmp("add_det twice", lambda: a.add_det(d), lambda: b.add_det(d))
        for tgt in ("doci", "fullci"):
            def conv(P, w, tgt=tgt):
                c = getattr(P, tgt + "_wfn")(w)
                return np.concatenate([[len(c), c.nbasis, c.nocc_up, c.nocc_dn], c.to_det_array().ravel().astype(np.int64)])
            cmp("convert %s -> %s" % (kind, tgt), lambda: conv(R, a), lambda: conv(M, b))
        a2, b2 = getattr(R, kind + "_wfn")(*args), getattr(M, kind + "_wfn")(*args)
        a2.add_hartreefock_det()
        b2.add_hartreefock_det()
        a2.add_dets_from_wfn(a)
        b2.add_dets_from_wfn(b)
        assert len(a2) == len(b2) and rows_as_set(a2.to_det_array()) == rows_as_set(b2.to_det_array())
    assert cmp.n > 2000


def reference_selectors():
    """the reference's Python selector modules bound to its compiled classes, or None outside the build container"""
    if R is None or not os.path.exists("/root/reference/pyci/gkci.py"):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_selectors", os.path.join(ROOT, "tests", "golden", "make_golden_selectors.py"))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    saved = {k: sys.modules.get(k) for k in ("pyci", "pyci._pyci", "pyci.utility", "pyci.seniority_ci", "pyci.cost_ci", "pyci.gkci")}
    try:
        return G.reference_python_layer()[1]
    finally:  # the stand-in `pyci` package must not leak into other tests
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.skipif(reference_selectors() is None, reason="needs /root/reference (build container)")
def test_selectors_against_the_reference_functions():
    m = reference_selectors()
    rng = np.random.default_rng(99)

    def same(a, b):
        return len(a) == len(b) and (len(a) == 0 or np.array_equal(a.to_det_array(), b.to_det_array()))

    def both(fa, fb):
        ea = eb = None
        try:
            fa()
        except Exception as e:  # noqa: BLE001
            ea = type(e).__name__
        try:
            fb()
        except Exception as e:  # noqa: BLE001
            eb = type(e).__name__
        assert ea == eb, (ea, eb)

    for _ in range(150):
        nb = int(rng.integers(2, 9))
        up = int(rng.integers(1, nb + 1))
        dn = int(rng.integers(0, up + 1))
        kind = str(rng.choice(["doci", "genci", "fullci"]))
        cost = rng.uniform(-1, 3, nb + 1)
        if rng.random() < 0.7:
            cost = np.sort(cost)
        t, q = float(rng.choice([-0.5, 0.0, 0.25, 1.0])), float(rng.uniform(-1, 8))
        args = (nb, up, up) if kind == "doci" else (nb, up, 0) if kind == "genci" else (nb, up, dn)
        a, b = getattr(R, kind + "_wfn")(*args), getattr(M, kind + "_wfn")(*args)
        fa = m["utility"].odometer_two_spin if kind == "fullci" else m["utility"].odometer_one_spin
        fb = M.odometer_two_spin if kind == "fullci" else M.odometer_one_spin
        both(lambda: fa(a, cost, t, q), lambda: fb(b, cost, t, q))
        assert same(a, b), ("odometer", kind, args, t, q)
        a, b = getattr(R, kind + "_wfn")(*args), getattr(M, kind + "_wfn")(*args)
        mode = str(rng.choice(["cntsp", "gamma", "interval", "nodes"]))
        kw = dict(t=t, p=float(rng.uniform(0.5, 2.5)))
        if mode == "interval":
            kw.update(mode="interval", energies=np.sort(rng.uniform(-2, 2, nb + 1)), width=float(rng.uniform(0.1, 1.5)))
        elif mode == "nodes":
            kw.update(mode=np.sort(rng.uniform(0, 4, nb + 1)))
        elif mode == "gamma":
            kw.update(mode="gamma", dim=int(rng.integers(1, 5)))
        both(lambda: m["gkci"].add_gkci(a, **kw), lambda: M.add_gkci(b, **kw))
        assert same(a, b), ("gkci", kind, args, mode)
        if kind == "fullci":
            a, b = R.fullci_wfn(*args), M.fullci_wfn(*args)
            sens = tuple(int(x) for x in rng.integers(0, nb + 2, int(rng.integers(1, 4))))
            both(lambda: m["seniority_ci"].add_seniorities(a, *sens), lambda: M.add_seniorities(b, *sens))
            assert same(a, b), ("seniority", args, sens)


@needs_ref
def test_hamiltonian_class_against_the_compiled_reference(tmp_path):
    """secondquant_op from arrays (other dtypes, Fortran order, negative strides: pyci.h:161-165 forcecast), its derived
    seniority-zero integrals, and FCIDUMP files written by either side: byte-identical files, identical objects read back by
    the other side, same exception types.  (Arrays of the wrong shape are refused here with ValueError; the reference takes
    them and reads past their end.)"""
    rng = np.random.default_rng(3)
    cmp = Pair()
    for it in range(24):
        n = int(rng.integers(1, 7))
        one, two = rng.standard_normal((n, n)), rng.standard_normal((n, n, n, n))
        if it % 4 == 1:
            one = one.astype(np.float32)
        elif it % 4 == 2:
            two = np.asfortranarray(two)
        elif it % 4 == 3:
            one = one[:, ::-1]
        ec = float(rng.standard_normal())
        A, B = R.secondquant_op(ec, one, two), M.secondquant_op(ec, one, two)
        for at in ("nbasis", "ecore", "one_mo", "two_mo", "h", "v", "w"):
            cmp(at, lambda: getattr(A, at), lambda: getattr(B, at))
        for k, kw in enumerate((dict(), dict(nelec=2, ms2=0), dict(nelec=3, ms2=1, tol=0.5), dict(tol=1e-3))):
            fa, fb = str(tmp_path / ("a%d_%d.fcidump" % (it, k))), str(tmp_path / ("b%d_%d.fcidump" % (it, k)))
            A.to_file(fa, **kw)
            B.to_file(fb, **kw)
            assert open(fa).read() == open(fb).read(), kw
            CA, CB = R.secondquant_op(fb), M.secondquant_op(fa)
            for at in ("nbasis", "ecore", "one_mo", "two_mo", "h", "v", "w"):
                cmp("read back " + at, lambda: getattr(CA, at), lambda: getattr(CB, at))
    cmp("missing file", lambda: R.secondquant_op("/nonexistent.fcidump"), lambda: M.secondquant_op("/nonexistent.fcidump"))
    cmp("bad arguments", lambda: R.secondquant_op(1.0), lambda: M.secondquant_op(1.0))
    with pytest.raises(ValueError):
        M.secondquant_op(0.0, np.zeros((3, 4)), np.zeros((3, 3, 3, 3)))


@needs_ref
def test_wavefunction_files_are_interchangeable_with_the_compiled_reference(tmp_path):
    """to_file / file constructors (onespinwfn.cpp:60-104, twospinwfn.cpp:60-107): byte-identical files, each side reads
    the other's, reading a file with another class behaves alike, a missing file raises the same error."""
    for k, (kind, args) in enumerate((("doci", (10, 3, 3)), ("doci", (70, 2, 2)), ("fullci", (8, 3, 2)), ("fullci", (130, 2, 1)),
                                      ("genci", (12, 4, 0)), ("genci", (65, 3, 0)), ("fullci", (6, 2, 0)))):
        a, b = getattr(R, kind + "_wfn")(*args), getattr(M, kind + "_wfn")(*args)
        for w in (a, b):
            w.add_hartreefock_det()
            w.add_excited_dets(1)
            w.add_excited_dets(2)
        fa, fb = str(tmp_path / ("a%d.bin" % k)), str(tmp_path / ("b%d.bin" % k))
        a.to_file(fa)
        b.to_file(fb)
        assert open(fa, "rb").read() == open(fb, "rb").read()
        ca, cb = getattr(R, kind + "_wfn")(fb), getattr(M, kind + "_wfn")(fa)
        assert len(ca) == len(cb) == len(a) and np.array_equal(ca.to_det_array(), cb.to_det_array())
        cmp = Pair()
        for other in ("doci", "fullci", "genci"):
            def load(P, other=other):
                x = getattr(P, other + "_wfn")(fa)
                return (len(x), x.nbasis, x.nocc_up, x.nocc_dn)
            cmp("%s file as %s" % (kind, other), lambda: load(R), lambda: load(M))
    for P in (R, M):
        with pytest.raises(RuntimeError, match="iostream error"):
            P.doci_wfn("/nonexistent")
