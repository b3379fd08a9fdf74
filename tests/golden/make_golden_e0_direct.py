#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Lowest eigenvalue of a synthetic FullCI workload (bench.py `synNN`; cfg4 = syn16) by a
STRING-DRIVEN direct-CI product on the CPU -- a different algorithm from both the product and the row-by-row oracle, and
the only one here that fits FullCI(16, 4a4b) (5.3e9 stored entries = 63 GB as a matrix) in this container's memory:

    H = H_a (x) 1 + 1 (x) H_b + sum_{pq,rs} (pq|rs) E^a_pq E^b_rs

* H_a = H_b: the one-spin Hamiltonian (one-body + same-spin two-body) over the C(n, 4) alpha strings = the oracle's
  FullCI(n, 4, 0) operator (oracle_sparse_op, pinned against the compiled reference in tests/test_oracle.py);
* the alpha-beta term (three quarters of the matrix) is NOT taken from the oracle: single-excitation operators E_pq of
  the strings as index/sign lists, D_rs = C E_rs^T, G = (pq|rs) D as one dense product, sigma += E_pq G_pq -- with the
  CI vector as the matrix C[alpha string, beta string] (determinant i = i_alpha * nstrings + i_beta, add_all_dets order).

ARPACK as in make_golden_e0.py (eigsh k=1, which='SA', tol=1e-12, ncv=30).  The script first reproduces the matrix-based
goldens already in e0_syn.json (syn8/10/12/14) -- that is its own check -- and then adds the sizes asked for:

    python tests/golden/make_golden_e0_direct.py --check 8 10 12      # |E0 - golden| of the sizes that have one
    python tests/golden/make_golden_e0_direct.py 16                   # ~14 GB of RAM, about half an hour on 8 cores
    python tests/golden/make_golden_e0_direct.py --spmv 14 16         # -> tests/golden/spmv_direct.npz

--spmv: y = H x of a seeded x (conftest.seeded_vec(ndet, 8)) by the same product, kept as 4096 sampled entries plus
|y| and x.y -- the fixture the full-size SpMV of the device is compared with (tests/test_zz_independent_product.py).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "e0_syn.json")
SEED = 1234
OCC = (4, 4)


def excitation_lists(strings, n):
    """for every (p, q): (rows I, columns J, signs) with <I| a+_p a_q |J> = sign, over the given bit strings"""
    index = {int(s): i for i, s in enumerate(strings)}
    out = {}
    for p in range(n):
        for q in range(n):
            I, J, S = [], [], []
            for j, s in enumerate(strings):
                s = int(s)
                if not (s >> q) & 1:
                    continue
                if p == q:
                    I.append(j), J.append(j), S.append(1.0)
                    continue
                if (s >> p) & 1:
                    continue
                lo, hi = (p, q) if p < q else (q, p)
                between = bin(s & (((1 << hi) - 1) ^ ((1 << (lo + 1)) - 1))).count("1")
                I.append(index[s ^ (1 << q) ^ (1 << p)]), J.append(j), S.append(-1.0 if between & 1 else 1.0)
            out[p, q] = (np.array(I, dtype=np.int64), np.array(J, dtype=np.int64), np.array(S))
    return out


class DirectCI:
    def __init__(self, n):
        from oracle import oracle as O
        assert OCC[0] == OCC[1]
        _, one, two = O.synthetic_integrals(n, SEED)
        sdets = O.all_dets(O.FULLCI, n, OCC[0], 0)
        self.strings = sdets[:, 0, 0].copy()
        self.ns = len(self.strings)
        ip, ix, dv = O.sparse_op(O.FULLCI, n, OCC[0], 0, sdets, (one, two), symmetric=True)
        self.h1 = O.full_symmetric(ip, ix, dv, self.ns).tocsr()  # one-spin Hamiltonian over the strings
        self.n = n
        self.exc = excitation_lists(self.strings, n)
        # (pq|rs) = <pr|qs> = two_mo[p, r, q, s]  (two_mo in physicist order, squantop.cpp:134-141)
        self.g = np.ascontiguousarray(two.transpose(0, 2, 1, 3).reshape(n * n, n * n))
        self.count = 0
        # sanity: the determinant order this factorisation assumes
        full = O.all_dets(O.FULLCI, n, *OCC)
        assert np.array_equal(full[:, 0, 0], np.repeat(self.strings, self.ns))
        assert np.array_equal(full[:, 1, 0], np.tile(self.strings, self.ns))

    def matvec(self, c):
        ns, n = self.ns, self.n
        C = np.asarray(c, dtype=np.float64).reshape(ns, ns)
        sigma = self.h1 @ C + (self.h1 @ C.T).T
        D = np.zeros((n * n, ns, ns))
        for (r, s), (I, J, S) in self.exc.items():  # D_rs = C E_rs^T: beta index
            D[r * n + s][:, I] = C[:, J] * S
        G = (self.g @ D.reshape(n * n, ns * ns)).reshape(n * n, ns, ns)
        del D
        for (p, q), (I, J, S) in self.exc.items():  # sigma += E_pq G_pq: alpha index
            sigma[I, :] += G[p * n + q][J, :] * S[:, None]
        self.count += 1
        return sigma.reshape(-1)


def lowest(n):
    import scipy.sparse.linalg as spla
    t0 = time.perf_counter()
    H = DirectCI(n)
    nd = H.ns * H.ns
    t1 = time.perf_counter()
    op = spla.LinearOperator((nd, nd), matvec=H.matvec, dtype=np.float64)
    v0 = np.random.default_rng(SEED).standard_normal(nd)
    w, _ = spla.eigsh(op, k=1, which="SA", tol=1e-12, ncv=30, v0=v0)
    t2 = time.perf_counter()
    return dict(E0=float(w[0]), arpack_matvecs=H.count, n=n, ndet=nd, occ=list(OCC), seed=SEED,
                operator="string-driven direct CI (tests/golden/make_golden_e0_direct.py): oracle one-spin Hamiltonian + "
                         "alpha-beta term from single-excitation lists",
                solver="scipy eigsh k=1 which=SA tol=1e-12 ncv=30", cpu_setup_seconds=t1 - t0, cpu_eigsh_seconds=t2 - t1)


def self_test(n=6):
    """the factorised product against the oracle's full FullCI(n, 4, 4) matrix on a random vector"""
    from oracle import oracle as O
    _, one, two = O.synthetic_integrals(n, SEED)
    dets = O.all_dets(O.FULLCI, n, *OCC)
    ip, ix, dv = O.sparse_op(O.FULLCI, n, OCC[0], OCC[1], dets, (one, two), symmetric=True)
    A = O.full_symmetric(ip, ix, dv, len(dets))
    x = np.random.default_rng(1).standard_normal(len(dets))
    y, z = A @ x, DirectCI(n).matvec(x)
    err = np.max(np.abs(y - z)) / np.max(np.abs(y))
    assert err < 1e-13, err
    return err


def spmv_fixture(sizes):
    out = {}
    for n in sizes:
        H = DirectCI(n)
        nd = H.ns * H.ns
        x = np.random.default_rng(8).standard_normal(nd)  # = tests/conftest.py seeded_vec(nd, 8)
        t0 = time.perf_counter()
        y = H.matvec(x)
        idx = np.sort(np.random.default_rng(n).choice(nd, 4096, replace=False))
        idx[0], idx[-1] = 0, nd - 1
        key = "syn%d" % n
        out[key + ".idx"], out[key + ".y"] = idx.astype(np.int64), y[idx]
        out[key + ".norm"], out[key + ".xdoty"], out[key + ".xnorm"] = np.linalg.norm(y), float(x @ y), np.linalg.norm(x)
        print("%s: %d determinants, |y| %.15e, x.y %.15e, %.1f s" % (key, nd, out[key + ".norm"], out[key + ".xdoty"],
                                                                     time.perf_counter() - t0), flush=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "spmv_direct.npz"), **out)


def main(argv):
    if "--spmv" in argv:
        print("self test (n = 6): rel err %.1e" % self_test(), flush=True)
        return spmv_fixture([int(a) for a in argv if a.isdigit()])
    check = "--check" in argv
    sizes = [int(a) for a in argv if a.isdigit()]
    print("self test (n = 6, product against the oracle's matrix): rel err %.1e" % self_test(), flush=True)
    with open(OUT) as f:
        gold = json.load(f)
    for n in sizes:
        r = lowest(n)
        key = "syn%d" % n
        if key in gold:
            print("%s: E0 %.14f, golden %.14f (%s), |diff| %.2e, %d matvecs, %.1f s" % (
                key, r["E0"], gold[key]["E0"], gold[key]["operator"][:24], abs(r["E0"] - gold[key]["E0"]), r["arpack_matvecs"],
                r["cpu_eigsh_seconds"]), flush=True)
            if not check:
                gold[key]["direct_ci_E0"] = r["E0"]
        else:
            print("%s: E0 %.14f, %d matvecs, setup %.1f s, eigsh %.1f s" % (key, r["E0"], r["arpack_matvecs"],
                                                                          r["cpu_setup_seconds"], r["cpu_eigsh_seconds"]), flush=True)
            if not check:
                gold[key] = r
        if not check:
            with open(OUT, "w") as f:
                json.dump(gold, f, indent=1, sort_keys=True)
                f.write("\n")


if __name__ == "__main__":
    main(sys.argv[1:])
