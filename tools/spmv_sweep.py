"""Developer probe: SpMV launch-shape sweep (threads per row x CTAs per SM) on one workload."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pyci_b200 as pyci  # noqa: E402
from pyci_b200 import cabi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
spec = bench.workload_spec(name)
ham, wfn = bench.make_problem(pyci, spec)
ctx = cabi.Context(0)
kind = {"doci": cabi.DOCI, "fullci": cabi.FULLCI, "genci": cabi.GENCI}[spec["kind"]]
dham = cabi.Ham(ctx, ham.nbasis, ham.ecore, ham.one_mo, ham.two_mo, ham.h, ham.v, ham.w)
dwfn = cabi.Wfn(ctx, kind, ham.nbasis, wfn.nocc_up, wfn.nocc_dn, wfn.to_det_array())
op = cabi.Op(ctx, dham, dwfn)
print(json.dumps(dict(workload=name, build=op.build_times(), nnz=op.stored_nnz)))
byt = op.stored_nnz * 12 + (op.row_count + 1) * 8 + op.row_count * 8 + op.ncol * 8
x = np.random.default_rng(0).standard_normal(op.ncol)
yref = None
shapes = [(64, 4, 256, 2), (64, 4, 256, 3), (32, 2, 512, 2), (32, 2, 512, 3), (32, 2, 512, 4), (32, 1, 1024, 3),
          (64, 2, 512, 3), (32, 4, 256, 3)]
if name == "cfg5" or (len(sys.argv) > 2 and sys.argv[2] == "short"):
    shapes = [(1, 8, 256, 2), (1, 6, 256, 2), (32, 4, 256, 2), (-8, 3, 256, 2), (-8, 6, 256, 2), (-16, 2, 256, 2), (-16, 4, 256, 2),
              (-32, 1, 256, 2), (-32, 2, 256, 2)]
for tpr, ctas, blk, depth in shapes:
    if True:
        op.set_spmv_shape(tpr, ctas)
        op.set_spmv_block(blk, depth)
        ms = op.time_spmv(3, 10, 0)
        y = op.matvec(x)
        if yref is None:
            yref = y
        err = float(np.max(np.abs(y - yref)) / np.max(np.abs(yref)))
        print(json.dumps(dict(tpr=tpr, ctas=ctas, block=blk, depth=depth, ms=float(np.mean(ms)), min_ms=float(ms.min()),
                              gbs=byt / (np.mean(ms) * 1e-3) / 1e9, relerr_vs_first=err)), flush=True)
