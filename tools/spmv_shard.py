"""Developer probe: SpMV of ONE rank's shard of the full-size config 5 operator on a single GPU (the columns span all
50 M determinants, so x is 400 MB and does not fit L2 -- the regime the 8-GPU run is in).
    PYCI_B200_SPMV_SEQ=k python tools/spmv_shard.py [ndet=50000000] [nranks=8] [rank=3]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyci_b200 import cabi  # noqa: E402
from pyci_b200.synthetic import seniority_zero_genci_dets, spin_orbital_integrals, synthetic_integrals  # noqa: E402

ND = int(sys.argv[1]) if len(sys.argv) > 1 else 50000000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
RK = int(sys.argv[3]) if len(sys.argv) > 3 else 3
K, NP = 32, 10
_, one, two = synthetic_integrals(K, 1234)
h2, g2 = spin_orbital_integrals(one, two)
ctx = cabi.Context(0)
ham = cabi.Ham(ctx, 2 * K, 0.0, h2, g2)
wfn = cabi.Wfn(ctx, cabi.GENCI, 2 * K, 2 * NP, 0, seniority_zero_genci_dets(K, NP, ND))
op = cabi.Op(ctx, ham, wfn, shard=(RK, R))
nbytes = op.stored_nnz * 12 + (op.row_count + 1) * 8 + op.row_count * 8 + op.ncol * 8
ms = float(np.mean(op.time_spmv(3, 10, 0)))
print(json.dumps({"seq": os.environ.get("PYCI_B200_SPMV_SEQ", "default"), "chunk": os.environ.get("PYCI_B200_SPMV_CHUNK", "default"),
                  "rows": int(op.row_count), "ncol": int(op.ncol), "nnz": int(op.stored_nnz), "ms": ms,
                  "gbs": nbytes / (ms * 1e-3) / 1e9}))
