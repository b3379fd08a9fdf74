"""Developer timing probe (not the graded bench): build + SpMV + solve timings for a few problem sizes."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyci_b200 as pyci  # noqa: E402
from oracle import oracle as O  # noqa: E402
from conftest import datafile  # noqa: E402


def run(tag, ham, wfn, solve=True, flush=0):
    t0 = time.time()
    op = pyci.sparse_op(ham, wfn)
    wall = time.time() - t0
    cold = op.stats()["fill_seconds"]
    fills = []
    for _ in range(int(os.environ.get("QB_REBUILDS", "3"))):  # warm rebuilds: module loaded, pool memory reused
        del op
        op = pyci.sparse_op(ham, wfn)
        fills.append(op.stats()["fill_seconds"])
    st = op.stats()
    if fills:
        st["fill_seconds"] = min(fills)
        st["build_seconds"] = st["build_seconds"] - fills[-1] + min(fills)
    ms = op.time_matvec(3, 10, flush)
    nnz = st["stored_nnz"]
    byt = nnz * 12 + (op.shape[0] + 1) * 8 + op.shape[0] * 8 + op.shape[1] * 8
    out = dict(tag=tag, ndet=len(wfn), size=int(op.size), stored_nnz=int(nnz), wall_build_s=round(wall, 4),
               hash_s=st["hash_seconds"], count_s=st["count_seconds"], fill_s=st["fill_seconds"], fill_cold_s=cold,
               nnz_per_s=op.size / max(st["build_seconds"], 1e-9), spmv_ms=float(np.median(ms)),
               spmv_gbs=byt / (np.median(ms) * 1e-3) / 1e9)
    if solve:
        t0 = time.time()
        es, cs = op.solve(n=1, tol=1e-9)
        out.update(E0=float(es[0]), solve_wall_s=round(time.time() - t0, 4), **{k: op.stats()[k] for k in
                   ("matvecs", "iterations", "solve_seconds", "spmv_seconds", "residual")})
    print(json.dumps(out), flush=True)
    return op


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg2", "cfg1", "syn12"]
    for w in which:
        if w == "cfg1":
            ham = pyci.hamiltonian(datafile("be_ccpvdz"))
            wfn = pyci.fullci_wfn(ham.nbasis, 2, 2)
        elif w == "cfg2":
            ham = pyci.hamiltonian(datafile("h2o_ccpvdz"))
            wfn = pyci.doci_wfn(ham.nbasis, 5, 5)
        else:
            n = int(w[3:])
            ham = pyci.hamiltonian(*O.synthetic_integrals(n, 1234))
            wfn = pyci.fullci_wfn(n, 4, 4)
        wfn.add_all_dets()
        run(w, ham, wfn, flush=(256 << 20) if w in ("cfg1", "cfg2") else 0)
