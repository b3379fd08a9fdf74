"""Developer probe: add_hci / compute_enpt2 on the device vs the CPU oracle on a selected FullCI space."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from pyci_b200 import cabi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 33
eps = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0e-4
occ = (4, 4)
ecore, one, two = O.synthetic_integrals(n, 1234)
dets = np.ascontiguousarray(O.all_dets(O.FULLCI, n, *occ)[::stride])
c = np.random.default_rng(1).standard_normal(len(dets))
c /= np.linalg.norm(c)
ctx = cabi.Context(0)
ham = cabi.Ham(ctx, n, ecore, one, two)
out = dict(n=n, ndet=len(dets), eps=eps)
for rep in range(2):
    wfn = cabi.Wfn(ctx, cabi.FULLCI, n, occ[0], occ[1], dets)
    t0 = time.time()
    pt, nt = wfn.compute_enpt2(ham, c, -10.0, eps)
    out["enpt2_wall_s"], out["enpt2_dev_s"], out["external"] = time.time() - t0, wfn.ext_seconds(), nt
    t0 = time.time()
    new = wfn.add_hci(ham, c, eps)
    out["hci_wall_s"], out["hci_dev_s"], out["added"] = time.time() - t0, wfn.ext_seconds(), len(new)
    wfn.close()
# CPU oracle on a row sample (cost per row is uniform)
k = min(len(dets), 2000)
t0 = time.time()
O.add_hci(O.FULLCI, n, occ[0], occ[1], dets[:k], (one, two), c[:k], eps)
out["oracle_rows_per_s"] = k / (time.time() - t0)
out["device_rows_per_s_hci"] = len(dets) / out["hci_dev_s"]
out["device_rows_per_s_enpt2"] = len(dets) / out["enpt2_dev_s"]
cand = 4 * (n - 4) * 2 + 2 * 6 * ((n - 4) * (n - 5) // 2) + (4 * (n - 4)) ** 2
out["candidates_per_s_hci"] = cand * len(dets) / out["hci_dev_s"]
print(json.dumps(out))
