#!/usr/bin/env python
r"""A PyCI script run on the B200 path: the only change against the reference is the import line.

    python examples/selected_ci.py [fcidump] [nocc_up] [nocc_dn]

FullCI of the FCIDUMP in one go (construction, lowest eigenpair, 1-/2-RDM) and then the same state by heat-bath
selection: HF -> [solve -> add_hci -> op.update]* with the Epstein-Nesbet PT2 estimate of every intermediate space.
Needs a CUDA device (there is no CPU fallback).  Mirrors pyci/test/test_routines.py:435-460 (run_hci) of the reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyci_b200 as pyci  # was: import pyci  # noqa: E402


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "data", "be_ccpvdz.fcidump")
    nup = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    ndn = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    ham = pyci.hamiltonian(path)

    # ---- the whole space at once
    wfn = pyci.fullci_wfn(ham.nbasis, nup, ndn)
    wfn.add_all_dets()
    op = pyci.sparse_op(ham, wfn)
    es, cs = op.solve(n=1, tol=1.0e-9)
    d1, d2 = pyci.compute_rdms(wfn, cs[0])
    print("FullCI: %d determinants, %d stored elements, E0 = %.10f, tr(rdm1) = %.6f"
          % (len(wfn), op.size, es[0], np.trace(d1[0]) + np.trace(d1[1])))
    e_fci = es[0]

    # ---- heat-bath selected CI towards the same state
    sel = pyci.fullci_wfn(ham.nbasis, nup, ndn)
    sel.add_hartreefock_det()
    op = pyci.sparse_op(ham, sel)
    es, cs = op.solve(n=1, tol=1.0e-9)
    for it in range(20):
        ept2 = pyci.compute_enpt2(ham, sel, cs[0], es[0], 1.0e-6)
        print("iter %2d: %7d determinants  E_var = %.10f  E_var+PT2 = %.10f  (E_FCI - E_var = %.2e)"
              % (it, len(sel), es[0], ept2, es[0] - e_fci))
        if pyci.add_hci(ham, sel, cs[0], eps=1.0e-4) == 0:
            break
        op.update(ham, sel)             # only the new determinants are enumerated
        es, cs = op.solve(n=1, tol=1.0e-9)
    print("overlap with the FullCI state: %.8f" % abs(pyci.compute_overlap(sel, wfn, cs[0], cs_full(wfn, ham))))


def cs_full(wfn, ham):
    return pyci.sparse_op(ham, wfn).solve(n=1, tol=1.0e-9)[1][0]


if __name__ == "__main__":
    main()
