"""Full-size and sharded parity of the sm_100a path, on ONE GPU, through the C ABI:

* sampled-row parity (bit-exact indptr / indices / data against the CPU oracle's row-list entry, which is pinned
  against the compiled reference in tests/test_oracle.py) of the operators bench.py times: BASELINE config 3, and
  config 4 both whole and as the row shards the 2/4/8-GPU runs build (`row0 != 0` branch of the complete-space fill);
* the lowest eigenvalue of config 3 against tests/golden/e0_syn.json (oracle-built operator + ARPACK);
* determinants unranked on the device (Wfn::add_all_dets), the row-list export, shard builds of every fill path;
* the segment-pair join of selected spaces (join.cuh) against the oracle and against enumerate-and-probe.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, seeded_vec
from oracle import oracle as O

pytestmark = pytest.mark.gpu

E_ATOL = 1e-10


@pytest.fixture(scope="module")
def cabi():
    from pyci_b200 import cabi as C
    assert C.lib().pyci_device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return C


@pytest.fixture(scope="module")
def ctx(cabi):
    c = cabi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def e0_golden():
    with open(os.path.join(GOLDEN, "e0_syn.json")) as f:
        return json.load(f)


def sample_rows(row0, nloc, nb, rng, nrandom, extra=()):
    """first / last rows of the shard, both sides of every alpha-string boundary near the ends and of a sample of the
    others, both sides of a sample of uniform CTA range boundaries (148 and 4 x 148 ranges), random rest"""
    rows = {row0, row0 + 1, row0 + nloc - 1, row0 + nloc - 2}
    bounds = np.arange(-(-row0 // nb) * nb, row0 + nloc, nb)
    if len(bounds) > 40:
        bounds = np.concatenate([bounds[:10], bounds[-10:], rng.choice(bounds[10:-10], 20, replace=False)])
    for parts in (148, 592):
        per = -(-nloc // parts)
        cb = row0 + per * np.arange(1, parts)
        bounds = np.concatenate([bounds, rng.choice(cb[cb < row0 + nloc], min(24, len(cb)), replace=False)])
    for b in bounds:
        rows.update((int(b) - 1, int(b), int(b) + 1))
    rows.update(int(r) for r in rng.integers(row0, row0 + nloc, nrandom))
    rows.update(int(r) for r in extra)
    rows = np.array(sorted(r for r in rows if row0 <= r < row0 + nloc), dtype=np.int64)
    return rows


def assert_rows_equal(op, rows, kind, n, occ, dets, ints, symmetric=True, ncol=-1):
    gi, gx, gd = op.export_rows(rows)
    oi, ox, od = O.sparse_op(kind, n, occ[0], occ[1], dets, ints, symmetric=symmetric, ncol=ncol, rows=rows)
    assert np.array_equal(gi, oi), "row lengths differ"
    assert np.array_equal(gx, ox), "column indices differ"
    assert np.array_equal(gd, od), "matrix elements differ (bitwise)"
    return len(ox)


@pytest.mark.parametrize("kind,n,occ", [("fullci", 9, (4, 3)), ("fullci", 6, (3, 3)), ("fullci", 5, (5, 2)),
                                        ("doci", 12, (4, 4)), ("genci", 10, (4, 0)), ("fullci", 34, (2, 1))])
def test_all_dets_generated_on_device(cabi, ctx, kind, n, occ):
    """pyci_wfn_create_all_dets unranks the reference's add_all_dets order (onespinwfn.cpp:173-217,
    twospinwfn.cpp:181-245) bit for bit; operators built from it equal those built from uploaded determinants."""
    K = {"doci": (cabi.DOCI, O.DOCI), "fullci": (cabi.FULLCI, O.FULLCI), "genci": (cabi.GENCI, O.GENCI)}[kind]
    ref = O.all_dets(K[1], n, *occ)
    w = cabi.Wfn(ctx, K[0], n, occ[0], occ[1])
    assert w.ndet == len(ref)
    got = w.download_dets()
    assert np.array_equal(got.reshape(ref.shape), ref)
    assert np.array_equal(w.index_dets(ref[::7]), np.arange(len(ref))[::7])  # the deferred index is built on demand
    ecore, one, two = O.synthetic_integrals(n, 3)
    ints = O.senzero_integrals(one, two) if kind == "doci" else (one, two)
    ham = cabi.Ham(ctx, n, ecore, one, two, *(ints if kind == "doci" else (None, None, None)))
    w2 = cabi.Wfn(ctx, K[0], n, occ[0], occ[1], ref)
    a, b = cabi.Op(ctx, ham, w), cabi.Op(ctx, ham, w2)
    ea, eb = a.export_csr(), b.export_csr()
    oi, ox, od = O.sparse_op(K[1], n, occ[0], occ[1], ref, ints)
    for x, y, z in zip(ea, eb, (oi, ox, od)):
        assert np.array_equal(x, y) and np.array_equal(x, z)
    assert a.size == b.size == len(ox)
    for h in (a, b, w, w2, ham):
        h.close()


def test_host_module_uses_device_generation(monkeypatch):
    """pyci_b200.sparse_op on a wave function filled by add_all_dets sends no determinants (full_space flag);
    anything added afterwards, or PYCI_B200_UPLOAD_DETS=1, takes the upload path: same operator either way."""
    import pyci_b200 as pyci
    n, occ = 8, (3, 2)
    ecore, one, two = O.synthetic_integrals(n, 5)
    ham = pyci.hamiltonian(ecore, one, two)
    wfn = pyci.fullci_wfn(n, *occ)
    wfn.add_all_dets()
    op = pyci.sparse_op(ham, wfn)
    monkeypatch.setenv("PYCI_B200_UPLOAD_DETS", "1")
    op2 = pyci.sparse_op(ham, wfn)
    monkeypatch.delenv("PYCI_B200_UPLOAD_DETS")
    for f in ("indptr", "indices", "data"):
        assert np.array_equal(getattr(op, f)(), getattr(op2, f)())
    oi, ox, od = O.sparse_op(O.FULLCI, n, occ[0], occ[1], wfn.to_det_array(), (one, two))
    assert np.array_equal(op.indptr(), oi) and np.array_equal(op.indices(), ox) and np.array_equal(op.data(), od)
    assert op.size == len(ox)
    # a selected space built from an array is not "full": upload path
    sel = pyci.fullci_wfn(n, occ[0], occ[1], np.ascontiguousarray(wfn.to_det_array()[::2]))
    ops = pyci.sparse_op(ham, sel)
    si, sx, sd = O.sparse_op(O.FULLCI, n, occ[0], occ[1], sel.to_det_array(), (one, two))
    assert np.array_equal(ops.indices(), sx) and np.array_equal(ops.data(), sd)


@pytest.mark.parametrize("symmetric", [True, False])
def test_export_rows_equals_export_csr(cabi, ctx, symmetric):
    n, occ = 9, (3, 3)
    ecore, one, two = O.synthetic_integrals(n, 17)
    dets = O.all_dets(O.FULLCI, n, *occ)
    ham = cabi.Ham(ctx, n, ecore, one, two)
    for d in (dets, np.ascontiguousarray(dets[np.random.default_rng(1).permutation(len(dets))[:4000]])):
        w = cabi.Wfn(ctx, cabi.FULLCI, n, occ[0], occ[1], d)
        op = cabi.Op(ctx, ham, w, symmetric=symmetric)
        ip, ix, dv = op.export_csr()
        rows = np.array([len(d) - 1, 0, 5, 5, 1234, 77], dtype=np.int64)
        gi, gx, gd = op.export_rows(rows)
        for k, r in enumerate(rows):
            assert np.array_equal(gx[gi[k]:gi[k + 1]], ix[ip[r]:ip[r + 1]])
            assert np.array_equal(gd[gi[k]:gi[k + 1]], dv[ip[r]:ip[r + 1]])
        gi, gx, gd = op.export_rows(np.zeros(0, dtype=np.int64))
        assert gi.tolist() == [0] and len(gx) == 0
        with pytest.raises(cabi.PyciError):
            op.export_rows([len(d)])
        op.close()
        w.close()
    ham.close()


@pytest.mark.parametrize("case", ["complete", "selected-unsorted", "sorted-incomplete", "genci", "doci"])
def test_row_shards_on_one_gpu(cabi, ctx, case):
    """pyci_op_build_shard(rank, nranks): every shard of a row-sharded operator (what the 2/4/8-GPU runs build,
    `row0 != 0`) equals the rows of the oracle; concatenated shards equal the single-rank export."""
    rng = np.random.default_rng(9)
    if case == "genci":
        n, occ, kind, ck = 14, (5, 0), O.GENCI, cabi.GENCI
        dets = O.all_dets(kind, n, 5)
        dets = np.ascontiguousarray(dets[rng.permutation(len(dets))[:1500]])
    elif case == "doci":
        n, occ, kind, ck = 16, (4, 4), O.DOCI, cabi.DOCI
        dets = O.all_dets(kind, n, 4)
    else:
        n, occ, kind, ck = 9, (3, 3), O.FULLCI, cabi.FULLCI
        dets = O.all_dets(kind, n, *occ)
        if case == "selected-unsorted":
            dets = np.ascontiguousarray(dets[rng.permutation(len(dets))[:5000]])
        elif case == "sorted-incomplete":
            dets = np.ascontiguousarray(dets[np.sort(rng.choice(len(dets), 5000, replace=False))])
    ecore, one, two = O.synthetic_integrals(n, 23)
    ints = O.senzero_integrals(one, two) if kind == O.DOCI else (one, two)
    ham = cabi.Ham(ctx, n, ecore, one, two, *(ints if kind == O.DOCI else (None, None, None)))
    w = cabi.Wfn(ctx, ck, n, occ[0], occ[1], dets)
    oi, ox, od = O.sparse_op(kind, n, occ[0], occ[1], dets, ints)
    whole = cabi.Op(ctx, ham, w)
    if case == "complete":
        assert whole.fill_kernel() == "fill_complete_kernel"
    x = seeded_vec(len(dets), 4)
    yfull = whole.matvec(x)
    for nranks in (2, 3, 8):
        from pyci_b200.distributed import concat_csr, row_partition
        shards = []
        for r, (lo, cnt) in enumerate(row_partition(len(dets), len(dets), nranks)):
            op = cabi.Op(ctx, ham, w, shard=(r, nranks))
            assert (op.row_begin, op.row_count) == (lo, cnt)
            shards.append(op.export_csr())
            if cnt:
                rows = np.unique(np.concatenate([[lo, lo + cnt - 1], rng.integers(lo, lo + cnt, 5)]))
                assert_rows_equal(op, rows, kind, n, occ, dets, ints)
            with pytest.raises(cabi.PyciError):
                op.matvec(x)  # host matvec needs the context's own layout
            op.close()
        ci, cx, cd = concat_csr(shards)
        assert np.array_equal(ci, oi) and np.array_equal(cx, ox) and np.array_equal(cd, od), (case, nranks)
    np.testing.assert_allclose(yfull, O.matvec(oi, ox, od, x, True), rtol=0, atol=1e-11 * np.abs(yfull).max())
    whole.close()
    w.close()
    ham.close()


def test_config3_sampled_rows_and_e0(cabi, ctx, e0_golden):
    """BASELINE config 3 (the operator bench.py times at N=1): >= 1000 sampled rows of the timed-size operator --
    first / last rows, alpha-string boundaries, CTA range boundaries, random rest -- bit-exact against the oracle, and
    E0 within 1e-10 Eh of the oracle-built operator's ARPACK value (tests/golden/e0_syn.json)."""
    n, occ = 14, (4, 4)
    ecore, one, two = O.synthetic_integrals(n, 1234)
    dets = O.all_dets(O.FULLCI, n, *occ)
    ham = cabi.Ham(ctx, n, ecore, one, two)
    w = cabi.Wfn(ctx, cabi.FULLCI, n, occ[0], occ[1])  # generated on the device
    op = cabi.Op(ctx, ham, w)
    assert op.fill_kernel() == "fill_complete_kernel" and op.stored_nnz == 2225444221 and op.size == 1113223111
    rows = sample_rows(0, len(dets), 1001, np.random.default_rng(3), 900)
    assert len(rows) >= 1000
    nent = assert_rows_equal(op, rows, O.FULLCI, n, occ, dets, (one, two))
    assert nent > 500 * len(rows) * 0.5
    es, _, st = op.solve(n=1, tol=1e-9)
    g = e0_golden["syn14"]
    assert g["ndet"] == len(dets) and abs(es[0] - (g["E0"] + ecore)) <= E_ATOL, (es[0], g["E0"])
    op.close()
    w.close()
    ham.close()


def test_config4_one_gpu_and_its_shards(cabi, ctx):
    """BASELINE config 4 (FullCI 16 orbitals 4a4b, 3 312 400 determinants, 1.06e10 stored entries = 127 GB) on one
    GPU: sampled rows of the whole operator and of shards 1 of 2, 3 of 4 and 5 / 7 of 8 (the row blocks the multi-GPU
    bench builds) bit-exact against the oracle; the SpMV of a shard equals the same rows of the whole product."""
    import torch
    n, occ = 16, (4, 4)
    _, total = torch.cuda.mem_get_info(0)
    if total < 170e9:
        pytest.skip("needs a 180 GB device")
    ecore, one, two = O.synthetic_integrals(n, 1234)
    dets = O.all_dets(O.FULLCI, n, *occ)
    ndet, nb = len(dets), 1820
    assert ndet == 3312400
    ham = cabi.Ham(ctx, n, ecore, one, two)
    w = cabi.Wfn(ctx, cabi.FULLCI, n, occ[0], occ[1])
    rng = np.random.default_rng(4)
    x = torch.from_numpy(seeded_vec(ndet, 8)).cuda()
    op = cabi.Op(ctx, ham, w)
    assert op.fill_kernel() == "fill_complete_kernel" and op.stored_nnz == ndet * 3193
    rows = sample_rows(0, ndet, nb, rng, 900)
    assert len(rows) >= 1000
    assert_rows_equal(op, rows, O.FULLCI, n, occ, dets, (one, two))
    y = torch.empty(ndet, dtype=torch.float64, device="cuda")
    op.matvec_dev(x.data_ptr(), y.data_ptr())
    ctx.synchronize()
    ywhole = y.cpu().numpy()
    op.close()
    for r, nranks in ((1, 2), (3, 4), (5, 8), (7, 8)):
        sh = cabi.Op(ctx, ham, w, shard=(r, nranks))
        lo, cnt = sh.row_begin, sh.row_count
        assert lo > 0 and sh.fill_kernel() == "fill_complete_kernel" and sh.stored_nnz == cnt * 3193
        rows = sample_rows(lo, cnt, nb, rng, 150)
        assert_rows_equal(sh, rows, O.FULLCI, n, occ, dets, (one, two))
        ys = torch.empty(cnt, dtype=torch.float64, device="cuda")
        sh.matvec_dev(x.data_ptr(), ys.data_ptr())
        ctx.synchronize()
        assert np.max(np.abs(ys.cpu().numpy() - ywhole[lo:lo + cnt])) <= 1e-12 * np.abs(ywhole).max()
        sh.close()
    w.close()
    ham.close()


def _selected(kind, n, occ, count, seed):
    alld = O.all_dets(kind, n, *occ)
    return np.ascontiguousarray(alld[np.random.default_rng(seed).permutation(len(alld))[:count]])


@pytest.mark.parametrize("kind,n,occ,count", [("genci", 16, (5, 0), 3000), ("fullci", 10, (3, 3), 6000),
                                              ("fullci", 34, (2, 1), 3000), ("genci", 6, (3, 0), 17),
                                              ("fullci", 5, (5, 2), 7), ("genci", 40, (3, 0), 2500)])
def test_segment_pair_join_equals_enumeration(monkeypatch, kind, n, occ, count):
    """PYCI_B200_FORCE_JOIN=1: the stored entries of a selected space come from the segment-pair join (join.cuh)
    instead of one index probe per candidate excitation -- operator (symmetric, non-symmetric, rectangular) and RDMs
    are those of the oracle bit for bit / to rounding, and equal the enumeration's (PYCI_B200_NO_JOIN=1)."""
    import pyci_b200 as pyci
    okind = {"genci": O.GENCI, "fullci": O.FULLCI}[kind]
    dets = _selected(okind, n, occ, count, 6)
    ecore, one, two = O.synthetic_integrals(n, 41)
    ham = pyci.hamiltonian(ecore, one, two)
    c = seeded_vec(len(dets), 2)
    c /= np.linalg.norm(c)
    res = {}
    for mode in ("PYCI_B200_FORCE_JOIN", "PYCI_B200_NO_JOIN", "PYCI_B200_FORCE_JOIN+overflow"):
        if mode.endswith("+overflow"):
            # a recorded-hit list of 6 entries per row: most rows overflow it, and a second join pass stages their
            # hits in the CSR arrays (the path rows of > 1024 entries take in a heat-bath space)
            monkeypatch.setenv("PYCI_B200_HITCAP", "6")
            mode = "PYCI_B200_FORCE_JOIN"
        monkeypatch.setenv(mode, "1")
        monkeypatch.setenv("PYCI_B200_NO_SORTED_PATH", "1")
        wfn = getattr(pyci, kind + "_wfn")(n, occ[0], occ[1], dets)
        out = []
        for kw in (dict(), dict(symmetric=False), dict(nrow=len(dets) - 3, ncol=len(dets) - 5, symmetric=False)):
            op = pyci.sparse_op(ham, wfn, **kw)
            oi, ox, od = O.sparse_op(okind, n, occ[0], occ[1], dets, (one, two), **kw)
            assert np.array_equal(op.indptr(), oi), (mode, kw)
            assert np.array_equal(op.indices(), ox), (mode, kw)
            assert np.array_equal(op.data(), od), (mode, kw)
            out.append(op.data())
        r1, r2 = pyci.compute_rdms(wfn, c)
        o1, o2 = O.compute_rdms(okind, n, occ[0], occ[1], dets, c)
        np.testing.assert_allclose(r1, o1, rtol=0, atol=1e-13)
        np.testing.assert_allclose(r2, o2, rtol=0, atol=1e-13)
        res[mode] = (out, r1, r2)
        monkeypatch.delenv(mode)
        monkeypatch.delenv("PYCI_B200_HITCAP", raising=False)


def test_join_is_chosen_for_config5_style_space(cabi, ctx):
    """A config-5-style space big enough for the automatic choice (GenCI, 32 spin-orbitals / 10 electrons, 10 615
    candidates per row, seniority-zero selection): the join path runs without any switch; operator rows, E0 and RDMs
    against the oracle."""
    from pyci_b200.synthetic import seniority_zero_genci_dets, spin_orbital_integrals, synthetic_integrals
    K, P = 16, 5
    _, one, two = synthetic_integrals(K, 1234)
    h2, g2 = spin_orbital_integrals(one, two)
    dets = seniority_zero_genci_dets(K, P, 4000)
    ham = cabi.Ham(ctx, 2 * K, 0.0, h2, g2)
    w = cabi.Wfn(ctx, cabi.GENCI, 2 * K, 2 * P, 0, dets)
    op = cabi.Op(ctx, ham, w)
    assert op.count_kernel() == "join_rows_kernel"
    rows = np.unique(np.concatenate([[0, len(dets) - 1], np.random.default_rng(2).integers(0, len(dets), 120)]))
    assert_rows_equal(op, rows, O.GENCI, 2 * K, (2 * P, 0), dets, (h2, g2))
    ip, ix, dv = op.export_csr()
    es, cs, _ = op.solve(n=1, tol=1e-10)
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    L = sp.csr_matrix((dv, ix, ip), shape=(len(dets),) * 2)
    A = L + sp.tril(L, -1).T
    e0 = spla.eigsh(A, k=1, which="SA", tol=1e-12)[0][0]
    assert abs(es[0] - e0) <= E_ATOL
    r1, r2 = cabi.compute_rdms(ctx, w, cabi.GENCI, 2 * K, cs[0])
    anti = g2 - g2.transpose(0, 1, 3, 2)
    e = np.einsum("ij,ij", h2, r1) + 0.25 * np.einsum("ijkl,ijkl", anti, r2)
    assert abs(e - es[0]) < 1e-9 and abs(np.trace(r1) - 2 * P) < 1e-10
    op.close()
    w.close()
    ham.close()


@pytest.mark.parametrize("kind,n,occ,count", [("genci", 14, (5, 0), 1800), ("fullci", 9, (3, 3), 5000)])
def test_rows_too_long_for_shared_memory_go_through_hbm(monkeypatch, kind, n, occ, count):
    """Rows longer than the shared-memory row buffer (dominant determinants of a heat-bath space reach 10^4 entries)
    are built by a second launch with the row buffers in HBM.  PYCI_B200_SHORT_CAP forces that launch for every row of
    more than 40 entries: same operator, bit for bit, with and without the segment-pair join."""
    import pyci_b200 as pyci
    okind = {"genci": O.GENCI, "fullci": O.FULLCI}[kind]
    dets = _selected(okind, n, occ, count, 8)
    ecore, one, two = O.synthetic_integrals(n, 19)
    ham = pyci.hamiltonian(ecore, one, two)
    oi, ox, od = O.sparse_op(okind, n, occ[0], occ[1], dets, (one, two))
    assert np.max(np.diff(oi)) > 40
    monkeypatch.setenv("PYCI_B200_SHORT_CAP", "40")
    monkeypatch.setenv("PYCI_B200_NO_SORTED_PATH", "1")
    for join in ("PYCI_B200_FORCE_JOIN", "PYCI_B200_NO_JOIN"):
        monkeypatch.setenv(join, "1")
        wfn = getattr(pyci, kind + "_wfn")(n, occ[0], occ[1], dets)
        op = pyci.sparse_op(ham, wfn)
        assert np.array_equal(op.indptr(), oi) and np.array_equal(op.indices(), ox) and np.array_equal(op.data(), od), join
        monkeypatch.delenv(join)


@pytest.mark.parametrize("sym", ["none", "kl-only", "8-fold", "8-fold+noslice"])
def test_complete_fill_with_and_without_integral_symmetry(monkeypatch, sym):
    """The reference indexes two_mo with no symmetry assumption (sparseop.cpp:287,307,332,352).  The complete-space fill
    keeps only k <= l of its shared-memory slices two_mo[i, :, a, :] when <ik|al> == <il|ak> holds bit for bit (checked
    at upload, pyci_ham::kl_sym) and the full n^2 otherwise: arbitrary integrals, integrals with just that symmetry and
    8-fold symmetric ones give the oracle's operator bit for bit."""
    import pyci_b200 as pyci
    n, occ = 9, (3, 3)
    rng = np.random.default_rng(77)
    if sym.startswith("8-fold"):
        ecore, one, two = O.synthetic_integrals(n, 5)
    else:
        ecore = 0.5
        one = rng.standard_normal((n, n))
        two = rng.standard_normal((n, n, n, n))
        if sym == "kl-only":
            two = np.ascontiguousarray(two + two.transpose(0, 3, 2, 1))  # two[i,k,a,l] == two[i,l,a,k] exactly
            assert np.array_equal(two, two.transpose(0, 3, 2, 1))
    if sym.endswith("noslice"):
        monkeypatch.setenv("PYCI_B200_NO_SLICE", "1")
    ham = pyci.hamiltonian(ecore, one, two)
    wfn = pyci.fullci_wfn(n, *occ)
    wfn.add_all_dets()
    for kw in (dict(), dict(symmetric=False)):
        op = pyci.sparse_op(ham, wfn, **kw)
        assert op.stats()["fill_kernel"] == "fill_complete_kernel"
        oi, ox, od = O.sparse_op(O.FULLCI, n, occ[0], occ[1], wfn.to_det_array(), (one, two), **kw)
        assert np.array_equal(op.indptr(), oi) and np.array_equal(op.indices(), ox)
        assert np.array_equal(op.data(), od), sym
    monkeypatch.setenv("PYCI_B200_NO_PACKED_SLICE", "1")
    op = pyci.sparse_op(ham, wfn)
    oi, ox, od = O.sparse_op(O.FULLCI, n, occ[0], occ[1], wfn.to_det_array(), (one, two))
    assert np.array_equal(op.data(), od)
