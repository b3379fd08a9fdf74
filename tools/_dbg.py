import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import pyci_b200 as pyci
from conftest import datafile
small = np.load("/root/repo/tests/golden/small.npz")
fn, kind, occ = "h4_sto3g", "fullci", (2, 2)
ham = pyci.hamiltonian(datafile(fn)); wfn = pyci.fullci_wfn(ham.nbasis, *occ); wfn.add_all_dets()
tag = f"{fn}.{kind}{occ[0]}{occ[1]}"
op = pyci.sparse_op(ham, wfn, symmetric=False)
ip, ix, d = small[tag + ".nonsym.indptr"], small[tag + ".nonsym.indices"], small[tag + ".nonsym.data"]
g = op.data()
dets = wfn.to_det_array()
bad = np.nonzero(g != d)[0]
print("nbad", len(bad), "of", len(d))
rows = np.searchsorted(ip, bad, side="right") - 1
for b, r in list(zip(bad, rows))[:40]:
    c = ix[b]
    da, db = dets[r]; ca, cb = dets[c]
    ea = bin(int(da) ^ int(ca)).count("1") // 2; eb = bin(int(db) ^ int(cb)).count("1") // 2
    print(r, c, "exc a,b =", ea, eb, "got", g[b], "want", d[b])
