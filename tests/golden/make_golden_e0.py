#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Generates tests/golden/e0_syn.json: lowest eigenvalues of the synthetic FullCI
workloads (bench.py `synNN`, cfg3 = syn14) computed WITHOUT any product code -- the operator is built row
block by row block with the CPU oracle's row-list entry (oracle_sparse_op_rows, pinned against the compiled
reference in tests/test_oracle.py) in a process pool, kept in the reference's lower-triangular CSR, and
diagonalised with ARPACK (scipy.sparse.linalg.eigsh, k=1, which='SA', tol=1e-12, ncv=30: BASELINE.md 4.3,
the stand-in for the reference's Spectra solver, sparseop.cpp:114-146).

    python tests/golden/make_golden_e0.py 10 11 12 14      # orbitals; 14 needs ~30 GB of RAM and ~15 min

The JSON also records the ARPACK matvec count and the CPU seconds (this container), which DESIGN.md quotes.
"""
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "e0_syn.json")
SEED = 1234
OCC = (4, 4)


def _chunk(args):
    n, a, b, tmp = args
    from oracle import oracle as O
    _, one, two = O.synthetic_integrals(n, SEED)
    dets = O.all_dets(O.FULLCI, n, *OCC)
    ip, ix, dv = O.sparse_op(O.FULLCI, n, OCC[0], OCC[1], dets, (one, two), symmetric=True, rows=np.arange(a, b))
    fn = os.path.join(tmp, "c_%d.npz" % a)
    np.savez(fn, ip=ip, ix=ix.astype(np.int32), dv=dv)
    return a, b, fn, int(ip[-1])


def run(n, procs):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from oracle import oracle as O
    nd = O.all_dets(O.FULLCI, n, *OCC).shape[0]
    t0 = time.perf_counter()
    step = max(256, nd // (procs * 24))
    # lower-triangular rows grow with the row index: interleave so that the pool stays busy
    jobs = [(n, a, min(nd, a + step)) for a in range(0, nd, step)]
    with tempfile.TemporaryDirectory(prefix="pyci_e0_") as tmp:
        with mp.Pool(procs) as pool:
            res = sorted(pool.map(_chunk, [j + (tmp,) for j in jobs[::-1]], chunksize=1))
        nnz = sum(r[3] for r in res)
        indptr = np.zeros(nd + 1, dtype=np.int64)
        indices = np.empty(nnz, dtype=np.int32)
        data = np.empty(nnz)
        pos = 0
        for a, b, fn, m in res:
            with np.load(fn) as f:
                indptr[a + 1:b + 1] = pos + f["ip"][1:]
                indices[pos:pos + m] = f["ix"]
                data[pos:pos + m] = f["dv"]
            pos += m
            os.remove(fn)
    t_build = time.perf_counter() - t0
    L = sp.csr_matrix((data, indices, indptr), shape=(nd, nd))
    diag = L.diagonal()
    LT = L.T.tocsr() if nnz < 4e8 else None  # transpose product through csc for the big case (no second copy)
    count = [0]

    def mv(x):
        count[0] += 1
        x = np.asarray(x).ravel()
        y = L @ x
        y += (LT @ x) if LT is not None else (L.T @ x)
        y -= diag * x
        return y

    t1 = time.perf_counter()
    w, _ = spla.eigsh(spla.LinearOperator((nd, nd), matvec=mv, dtype=np.float64), k=1, which="SA", tol=1e-12, ncv=30)
    t_eig = time.perf_counter() - t1
    return {"n": n, "occ": list(OCC), "seed": SEED, "ndet": int(nd), "nnz_lower": int(nnz), "E0": float(w[0]),
            "arpack_matvecs": count[0], "cpu_build_seconds": t_build, "cpu_build_processes": procs,
            "cpu_eigsh_seconds": t_eig, "solver": "scipy eigsh k=1 which=SA tol=1e-12 ncv=30",
            "operator": "oracle_sparse_op_rows, lower-triangular CSR"}


def main():
    ns = [int(a) for a in sys.argv[1:]] or [10, 11, 12]
    procs = int(os.environ.get("E0_PROCS", str(os.cpu_count() or 1)))
    out = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            out = json.load(f)
    for n in ns:
        r = run(n, procs)
        out["syn%d" % n] = r
        print(json.dumps(r), flush=True)
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
