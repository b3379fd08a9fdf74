"""Developer probe: Davidson on a config-5-style operator (short rows) for several subspace sizes.
    python tools/solver_probe.py [K=32] [npair=10] [ndet=5000000] [ncv,ncv,...]
Environment (read by the library): PYCI_B200_SOLVER_GS2=1 (always two Gram-Schmidt passes), PYCI_B200_SOLVER_KEEP=k."""
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
NCVS = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [-1]
_, one, two = synthetic_integrals(K, 1234)
h2, g2 = spin_orbital_integrals(one, two)
ctx = cabi.Context(0)
ham = cabi.Ham(ctx, 2 * K, 0.0, h2, g2)
wfn = cabi.Wfn(ctx, cabi.GENCI, 2 * K, 2 * NP, 0, seniority_zero_genci_dets(K, NP, ND))
op = cabi.Op(ctx, ham, wfn)
for ncv in NCVS:
    for rep in range(2):
        es, cs, st = op.solve(n=1, ncv=ncv, tol=1e-9)
    print(json.dumps({"ncv": ncv, "gs2": os.environ.get("PYCI_B200_SOLVER_GS2", ""), "keep": os.environ.get("PYCI_B200_SOLVER_KEEP", ""),
                      "E0": float(es[0]), "matvecs": st["matvecs"], "restarts": st["restarts"], "seconds": st["seconds"],
                      "spmv_seconds": st["spmv_seconds"], "non_spmv_share": 1.0 - st["spmv_seconds"] / st["seconds"],
                      "residual": st["residual"]}), flush=True)
