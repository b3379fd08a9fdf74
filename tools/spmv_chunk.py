"""Developer probe: SpMV of a config-5-style operator (short rows) for one value of PYCI_B200_SPMV_CHUNK.
    PYCI_B200_SPMV_CHUNK=64 python tools/spmv_chunk.py [K=32] [npair=10] [ndet=5000000]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyci_b200 import cabi  # noqa: E402
from pyci_b200.synthetic import seniority_zero_genci_dets, spin_orbital_integrals, synthetic_integrals  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 32
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ND = int(sys.argv[3]) if len(sys.argv) > 3 else 5000000
_, one, two = synthetic_integrals(K, 1234)
h2, g2 = spin_orbital_integrals(one, two)
ctx = cabi.Context(0)
ham = cabi.Ham(ctx, 2 * K, 0.0, h2, g2)
wfn = cabi.Wfn(ctx, cabi.GENCI, 2 * K, 2 * NP, 0, seniority_zero_genci_dets(K, NP, ND))
op = cabi.Op(ctx, ham, wfn)
ms = op.time_spmv(3, 20, 0)
nbytes = op.stored_nnz * 12 + (op.row_count + 1) * 8 + op.row_count * 8 + op.ncol * 8
print(json.dumps({"chunk": os.environ.get("PYCI_B200_SPMV_CHUNK", "0"), "ndet": wfn.ndet, "nnz": int(op.stored_nnz),
                  "ms": float(np.mean(ms)), "gbs": nbytes / (float(np.mean(ms)) * 1e-3) / 1e9,
                  "build_s": float(op.build_times()["total"])}))
