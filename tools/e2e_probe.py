"""Developer probe: where the end-to-end time of pyci_b200.sparse_op(ham, wfn) goes (host wall clock per C-ABI call)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pyci_b200 as pyci  # noqa: E402
from pyci_b200 import cabi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
spec = bench.workload_spec(name)
ham, wfn = bench.make_problem(pyci, spec)
kind = {"doci": cabi.DOCI, "fullci": cabi.FULLCI, "genci": cabi.GENCI}[spec["kind"]]
ctx = cabi.Context(0)
dets = wfn.to_det_array()
out = []
for rep in range(4):
    t = [time.perf_counter()]
    dets2 = wfn.to_det_array(); t.append(time.perf_counter())
    dham = cabi.Ham(ctx, ham.nbasis, ham.ecore, ham.one_mo, ham.two_mo, ham.h, ham.v, ham.w); t.append(time.perf_counter())
    dwfn = cabi.Wfn(ctx, kind, ham.nbasis, wfn.nocc_up, wfn.nocc_dn, dets); t.append(time.perf_counter())
    op = cabi.Op(ctx, dham, dwfn); t.append(time.perf_counter())
    ip = np.empty(op.row_count + 1, dtype=np.int64)
    cabi.check(cabi.lib().pyci_op_export_csr(op.handle, ip.ctypes.data, None, None)); t.append(time.perf_counter())
    op.close(); dwfn.close(); dham.close(); t.append(time.perf_counter())
    t0 = time.perf_counter()
    o = pyci.sparse_op(ham, wfn); t1 = time.perf_counter(); o.indptr(); t2 = time.perf_counter(); del o; t3 = time.perf_counter()
    out.append(dict(to_det_array=t[1] - t[0], ham=t[2] - t[1], wfn=t[3] - t[2], build=t[4] - t[3], indptr=t[5] - t[4],
                    close=t[6] - t[5], api_sparse_op=t1 - t0, api_indptr=t2 - t1, api_del=t3 - t2,
                    dev=op.build_times() if False else None))
for o in out:
    print(json.dumps({k: (round(1e3 * v, 3) if v is not None else None) for k, v in o.items()}))
