"""Config-5-style space grown by the heat-bath loop ON THE DEVICE (the reference's test_run_hci loop,
pyci/test/test_routines.py:431-470: add_hci -> op.update / rebuild -> solve), then construction, SpMV, E0 and RDMs on
the grown space with a sampled-row parity check against the CPU oracle.  GenCI over 2K spin-orbitals, synthetic
integrals (bench.py's config 5 Hamiltonian); unlike bench.py's seniority-zero selection this space has no structure.

    python tools/hci_grow.py [K=32] [npair=10] [target_ndet=2000000] [eps0=0.4]

Prints one JSON object (profiles/r2_hci_grown.json is a copy of one run).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyci_b200 import cabi  # noqa: E402
from pyci_b200.synthetic import spin_orbital_integrals, synthetic_integrals  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 32
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 10
TARGET = int(sys.argv[3]) if len(sys.argv) > 3 else 2000000
EPS0 = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4

_, one, two = synthetic_integrals(K, 1234)
h2, g2 = spin_orbital_integrals(one, two)
ctx = cabi.Context(0)
ham = cabi.Ham(ctx, 2 * K, 0.0, h2, g2)
hf = np.array([[((1 << NP) - 1) | (((1 << NP) - 1) << K)]], dtype=np.uint64)
wfn = cabi.Wfn(ctx, cabi.GENCI, 2 * K, 2 * NP, 0, hf)
c = np.ones(1)
eps = EPS0
steps = []
t_all = time.time()
while wfn.ndet < TARGET and eps > 1e-8:
    t0 = time.time()
    new = wfn.add_hci(ham, c, eps)
    t_hci = time.time() - t0
    if len(new) == 0:
        eps *= 0.5
        continue
    op = cabi.Op(ctx, ham, wfn)
    es, cs, st = op.solve(n=1, tol=1e-6)
    steps.append(dict(eps=eps, ndet=wfn.ndet, added=int(len(new)), add_hci_device_s=wfn.ext_seconds(), add_hci_wall_s=t_hci,
                      build_s=op.build_times()["total"], count_kernel=op.count_kernel(), E0=float(es[0]),
                      matvecs=st["matvecs"], solve_s=st["seconds"]))
    sys.stderr.write("[hci_grow] eps %.3g -> %d determinants, E0 %.10f\n" % (eps, wfn.ndet, es[0]))
    c = cs[0]
    op.close()
    # aim the next step at roughly x30 growth at most: halve eps while the space is small, gentler near the target
    eps *= 0.5 if wfn.ndet * 30 < TARGET else 0.8
out = dict(K=K, electrons=2 * NP, ndet=wfn.ndet, growth=steps, grow_wall_s=time.time() - t_all)

# the grown space: construction (timed twice: second is warm), SpMV, E0, RDMs
for rep in range(2):
    wfn.reindex()
    op = cabi.Op(ctx, ham, wfn)
    bt = op.build_times()
out["build"] = dict(index_s=float(bt["index"]), count_scan_s=float(bt["count_scan"]), fill_sort_s=float(bt["fill_sort"]),
                    total_s=float(bt["total"]), count_kernel=op.count_kernel(), fill_kernel=op.fill_kernel(),
                    stored_nnz=int(op.stored_nnz), size_reference=int(op.size))
ms = op.time_spmv(3, 10, 0)
nbytes = op.stored_nnz * 12 + (op.row_count + 1) * 8 + op.row_count * 8 + op.ncol * 8
out["spmv"] = dict(ms=float(np.mean(ms)), gbs=nbytes / (float(np.mean(ms)) * 1e-3) / 1e9)
es, cs, st = op.solve(n=1, tol=1e-9)
out["solve"] = dict(E0=float(es[0]), matvecs=st["matvecs"], seconds=st["seconds"], residual=st["residual"])
t0 = time.time()
r1, r2 = cabi.compute_rdms(ctx, wfn, cabi.GENCI, 2 * K, cs[0])
out["rdm_wall_s"] = time.time() - t0
anti = g2 - g2.transpose(0, 1, 3, 2)
out["rdm_energy_identity_abs_error"] = abs(float(np.einsum("ij,ij", h2, r1) + 0.25 * np.einsum("ijkl,ijkl", anti, r2)) - float(es[0]))

# parity of sampled rows against the CPU oracle (the checker)
from oracle import oracle as O  # noqa: E402
dets = wfn.download_dets()
rows = np.unique(np.concatenate([[0, wfn.ndet - 1], np.random.default_rng(7).integers(0, wfn.ndet, 150)]))
gi, gx, gd = op.export_rows(rows)
oi, ox, od = O.sparse_op(O.GENCI, 2 * K, 2 * NP, 0, dets, (h2, g2), rows=rows)
out["parity"] = dict(rows=int(len(rows)), entries=int(len(ox)),
                     structure_equal=bool(np.array_equal(gi, oi) and np.array_equal(gx, ox)),
                     data_bit_identical=bool(np.array_equal(gd, od)))
print(json.dumps(out))
