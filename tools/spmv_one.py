"""Developer probe: build one workload and launch the SpMV kernel a few times in a given shape (for ncu)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pyci_b200 as pyci  # noqa: E402
from pyci_b200 import cabi  # noqa: E402

name, tpr, ctas, blk = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
depth = int(sys.argv[5]) if len(sys.argv) > 5 else 2
spec = bench.workload_spec(name)
ham, wfn = bench.make_problem(pyci, spec)
ctx = cabi.Context(0)
kind = {"doci": cabi.DOCI, "fullci": cabi.FULLCI}[spec["kind"]]
dham = cabi.Ham(ctx, ham.nbasis, ham.ecore, ham.one_mo, ham.two_mo, ham.h, ham.v, ham.w)
dwfn = cabi.Wfn(ctx, kind, ham.nbasis, wfn.nocc_up, wfn.nocc_dn, wfn.to_det_array())
op = cabi.Op(ctx, dham, dwfn)
op.set_spmv_shape(tpr, ctas)
op.set_spmv_block(blk, depth)
ms = op.time_spmv(1, 2, 0)
print(name, tpr, ctas, blk, float(np.mean(ms)))
