"""Golden vectors for compute_transition_rdms / compute_overlap from the REFERENCE'S OWN compiled sources.

Run in the build container only (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden_trdm.py        ->  tests/golden/trdm.npz

Two different selections of the same space with seeded coefficient vectors; stored: the inputs (determinant
arrays, coefficients) and the reference's outputs (transition 1-/2-RDMs, overlap).  DOCI and FullCI only: the
reference's GenCI routine is defective (rdm.cpp:904-1009; see oracle/pyci_oracle.c).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
DATA = os.path.join(ROOT, "tests", "data")

CASES = [("h6_fullci", "h6_sto_3g", "fullci_wfn", (3, 3)), ("lih_fullci", "lih_sto6g", "fullci_wfn", (2, 1)),
         ("be_doci", "be_ccpvdz", "doci_wfn", (2, 2)), ("h4_fullci", "h4_sto3g", "fullci_wfn", (2, 2))]


def main():
    import pyci_ref as pyci
    rng = np.random.default_rng(2024)
    out = {}
    for tag, fn, cls, occ in CASES:
        ham = pyci.secondquant_op(os.path.join(DATA, fn + ".fcidump"))
        w = getattr(pyci, cls)(ham.nbasis, *occ)
        w.add_all_dets()
        d = w.to_det_array()
        i1 = rng.permutation(len(d))[: max(1, len(d) * 2 // 3)]
        i2 = rng.permutation(len(d))[: max(1, len(d) * 3 // 4)]
        w1 = getattr(pyci, cls)(ham.nbasis, occ[0], occ[1], d[i1])
        w2 = getattr(pyci, cls)(ham.nbasis, occ[0], occ[1], d[i2])
        c1, c2 = rng.standard_normal(len(w1)), rng.standard_normal(len(w2))
        r1, r2 = pyci.compute_transition_rdms(w1, w2, c1, c2)
        out[tag + ".dets1"], out[tag + ".dets2"] = w1.to_det_array(), w2.to_det_array()
        out[tag + ".c1"], out[tag + ".c2"] = c1, c2
        out[tag + ".rdm1"], out[tag + ".rdm2"] = r1, r2
        out[tag + ".overlap"] = np.array(pyci.compute_overlap(w1, w2, c1, c2))
        print(tag, len(w1), len(w2), r1.shape, r2.shape, float(out[tag + ".overlap"]))
    np.savez_compressed(os.path.join(HERE, "trdm.npz"), **out)


if __name__ == "__main__":
    main()
