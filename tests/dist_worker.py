"""Worker of tests/test_gpu_multi.py: one process per GPU (torch.distributed.run), row-sharded operator.
Checks the sharded path against the CPU oracle on every rank: CSR shards (bit-exact structure), SpMV, lowest
eigenvalue (NCCL all-gather of the trial vector), RDMs (all-reduce), add_hci / compute_enpt2 (all-gather +
merge of the per-rank external lists)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch.distributed as dist  # noqa: E402

import pyci_b200 as pyci  # noqa: E402
from conftest import datafile, seeded_vec  # noqa: E402
from oracle import oracle as O  # noqa: E402
from pyci_b200.distributed import concat_csr, init_from_env, row_partition  # noqa: E402

rank, world = init_from_env("nccl")

# ---- Be cc-pVDZ FullCI(2,2): config 1, row-sharded
path = datafile("be_ccpvdz")
ham = pyci.hamiltonian(path)
ecore, one, two = O.read_fcidump(path)
wfn = pyci.fullci_wfn(ham.nbasis, 2, 2)
wfn.add_all_dets()
ndet = len(wfn)
op = pyci.sparse_op(ham, wfn)
ip, ix, dv = O.sparse_op(O.FULLCI, ham.nbasis, 2, 2, wfn.to_det_array(), (one, two))
lo, cnt = row_partition(ndet, ndet, world)[rank]
st = op.stats()
assert (st["row_begin"], st["row_count"]) == (lo, cnt), (st["row_begin"], st["row_count"], lo, cnt)
mine = (op.indptr(), op.indices(), op.data())
assert np.array_equal(mine[0], ip[lo:lo + cnt + 1] - ip[lo])
assert np.array_equal(mine[1], ix[ip[lo]:ip[lo + cnt]])
assert np.array_equal(mine[2], dv[ip[lo]:ip[lo + cnt]])
shards = [None] * world
dist.all_gather_object(shards, mine)
whole = concat_csr(shards)
assert np.array_equal(whole[0], ip) and np.array_equal(whole[1], ix) and np.array_equal(whole[2], dv)
x = seeded_vec(ndet, 3)
y = op(x)
yo = O.matvec(ip, ix, dv, x, True)
assert y.shape == (ndet,) and np.max(np.abs(y - yo)) <= 1e-11 * np.max(np.abs(yo))
es, cs = op.solve(n=1, tol=1e-9)
assert abs(es[0] - (-14.617409507)) < 1e-8, es  # pyci/test/test_routines.py:44
d1, d2 = pyci.compute_rdms(wfn, cs[0])
o1, o2 = O.compute_rdms(O.FULLCI, ham.nbasis, 2, 2, wfn.to_det_array(), cs[0])
assert np.max(np.abs(d1 - o1)) < 1e-12 and np.max(np.abs(d2 - o2)) < 1e-12

# ---- selected space: add_hci / compute_enpt2 with the external lists merged over ranks
n, occ = 10, (3, 3)
_, s1, s2 = O.synthetic_integrals(n, 4321)
hams = pyci.hamiltonian(0.0, s1, s2)
full = pyci.fullci_wfn(n, *occ)
full.add_all_dets()
fd = np.ascontiguousarray(full.to_det_array()[::9])
sel = pyci.fullci_wfn(n, occ[0], occ[1], fd)
c = seeded_vec(len(fd), 3)
c /= np.linalg.norm(c)
pt = pyci.compute_enpt2(hams, sel, c, -1.0, 0.02)
pto, nt = O.compute_enpt2(O.FULLCI, n, occ[0], occ[1], fd, (s1, s2), c, -1.0, 0.0, 0.02)
assert nt > 0 and abs(pt - pto) <= 1e-12 * abs(pto), (pt, pto)
new = O.add_hci(O.FULLCI, n, occ[0], occ[1], fd, (s1, s2), c, 0.02)
nadd = pyci.add_hci(hams, sel, c, eps=0.02)
assert nadd == len(new) and np.array_equal(sel.to_det_array()[len(fd):], new)
# the grown space builds and solves sharded
op2 = pyci.sparse_op(hams, sel)
e2, _ = op2.solve(n=1, tol=1e-9)
ip2, ix2, dv2 = O.sparse_op(O.FULLCI, n, occ[0], occ[1], sel.to_det_array(), (s1, s2))
e2o, _ = O.lowest_eigenpair(ip2, ix2, dv2, len(sel))
assert abs(e2[0] - e2o) < 1e-9, (e2, e2o)
# ---- nnz-balanced row partition (rebalance.cu): the grown space has long rows first (the determinants the selection
# started from) and short ones behind; any imbalance is evened out here (PYCI_B200_REBALANCE_MIN=1)
os.environ["PYCI_B200_REBALANCE_MIN"] = "1.0"
op3 = pyci.sparse_op(hams, sel)
del os.environ["PYCI_B200_REBALANCE_MIN"]
st2, st3 = op2.stats(), op3.stats()
parts = [None] * world
dist.all_gather_object(parts, (st2["stored_nnz"], st3["row_begin"], st3["row_count"], st3["stored_nnz"]))
assert parts[0][1] == 0 and sum(p[2] for p in parts) == len(sel)
assert all(parts[k][1] + parts[k][2] == parts[k + 1][1] for k in range(world - 1))  # contiguous, in rank order
assert sum(p[3] for p in parts) == sum(p[0] for p in parts)
uniform_worst, balanced_worst = max(p[0] for p in parts), max(p[3] for p in parts)
longest_row = int(np.max(np.diff(ip2))) * 2
assert balanced_worst <= uniform_worst and balanced_worst <= sum(p[3] for p in parts) / world + longest_row, parts
lo3, cnt3 = st3["row_begin"], st3["row_count"]
assert np.array_equal(op3.indptr(), ip2[lo3:lo3 + cnt3 + 1] - ip2[lo3])
assert np.array_equal(op3.indices(), ix2[ip2[lo3]:ip2[lo3 + cnt3]])
assert np.array_equal(op3.data(), dv2[ip2[lo3]:ip2[lo3 + cnt3]])
x3 = seeded_vec(len(sel), 8)
y3, y3o = op3(x3), O.matvec(ip2, ix2, dv2, x3, True)
assert np.max(np.abs(y3 - y3o)) <= 1e-11 * np.max(np.abs(y3o))
e3, c3 = op3.solve(n=2, tol=1e-9)
assert abs(e3[-1] - e2o) < 1e-9 and e3[0] > e3[-1] and c3.shape == (2, len(sel)), (e3, e2o)  # largest of the n lowest first
r3 = op3(c3[-1]) - (e3[-1] - 0.0) * c3[-1]
assert np.linalg.norm(r3) < 1e-6
if rank == 0:
    print("rebalance: worst rank %d -> %d stored entries of %d" % (uniform_worst, balanced_worst, sum(p[3] for p in parts)), flush=True)
# ---- selected GenCI space through the segment-pair join (join.cuh), row-sharded: this rank's rows x all determinants
os.environ["PYCI_B200_FORCE_JOIN"] = "1"
ng, og = 16, 5
_, g1, g2 = O.synthetic_integrals(ng, 77)
alld = O.all_dets(O.GENCI, ng, og)
gd = np.ascontiguousarray(alld[np.random.default_rng(12).permutation(len(alld))[:3000]])
gw = pyci.genci_wfn(ng, og, 0, gd)
gh = pyci.hamiltonian(0.0, g1, g2)
opg = pyci.sparse_op(gh, gw)
gi, gx, gv = O.sparse_op(O.GENCI, ng, og, 0, gd, (g1, g2))
stg = opg.stats()  # (a selected space may have been re-partitioned to equal stored entries per rank)
lo, cnt = stg["row_begin"], stg["row_count"]
assert np.array_equal(opg.indptr(), gi[lo:lo + cnt + 1] - gi[lo])
assert np.array_equal(opg.indices(), gx[gi[lo]:gi[lo + cnt]])
assert np.array_equal(opg.data(), gv[gi[lo]:gi[lo + cnt]])
cg = seeded_vec(len(gd), 5)
cg /= np.linalg.norm(cg)
r1, r2 = pyci.compute_rdms(gw, cg)
q1, q2 = O.compute_rdms(O.GENCI, ng, og, 0, gd, cg)
assert np.max(np.abs(r1 - q1)) < 1e-12 and np.max(np.abs(r2 - q2)) < 1e-12
del os.environ["PYCI_B200_FORCE_JOIN"]
dist.barrier()
print("rank %d of %d ok: launches %d" % (rank, world, pyci.launch_count()), flush=True)
dist.destroy_process_group()
