"""Golden vectors for SparseOp::update (sparseop.cpp:175-201) from the REFERENCE'S OWN compiled sources
(oracle/_ref/pyci_ref; build container only):

    python tests/golden/make_golden_update.py   ->  tests/golden/update.npz

Per case: the final determinant list, the sizes the wave function had at construction and at every update, and the
operator's CSR after the last update (row pointer + sha256 of indices and data) -- for symmetric, non-symmetric and rectangular non-symmetric operators.  What the
reference does to a non-symmetric operator is the point: rows it already has keep the columns they were built with,
only the appended rows see the grown wave function.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
DATA = "/root/reference/pyci/test/data/"

CASES = [  # tag, fcidump, kind, occupations, excitation levels added per stage
    ("be.fullci22", "be_ccpvdz", "fullci", (2, 2), [(0, 1), (2,), (3,)]),
    ("be.doci22", "be_ccpvdz", "doci", (2, 2), [(0, 1), (2,)]),
]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def add_excitations(wfn, *levels):  # pyci.add_excitations (pyci/utility.py) on the compiled module
    for e in levels:
        wfn.add_excited_dets(e)


def main():
    import pyci_ref as pyci

    out = {}
    for tag, fn, kind, occ, stages in CASES:
        ham = pyci.secondquant_op(DATA + fn + ".fcidump")
        wfn = getattr(pyci, kind + "_wfn")(ham.nbasis, *occ)
        for mode, (rows_short, symm) in {"sym": (0, True), "nonsym": (0, False), "rect": (7, False)}.items():
            w = type(wfn)(wfn)
            add_excitations(w, *stages[0])
            sizes = [len(w)]
            nrow0 = len(w) - rows_short
            op = pyci.sparse_op(ham, w, nrow0, len(w), symmetric=symm)
            for st in stages[1:]:
                add_excitations(w, *st)
                sizes.append(len(w))
                op.update(ham, w)
            key = "%s.%s" % (tag, mode)
            out[key + ".dets"] = w.to_det_array()
            out[key + ".sizes"] = np.array(sizes, dtype=np.int64)
            out[key + ".nrow0"] = np.array(nrow0, dtype=np.int64)
            out[key + ".shape"] = np.array(op.shape, dtype=np.int64)
            out[key + ".nnz"] = np.array(op.size, dtype=np.int64)
            out[key + ".indptr"] = op.indptr()
            # (the larger arrays as digests: the fixture stays small)
            out[key + ".indices.sha256"] = np.array(sha(op.indices().astype(np.int64)))
            out[key + ".data.sha256"] = np.array(sha(op.data().astype(np.float64)))
            print(key, sizes, nrow0, op.shape, op.size)
    np.savez_compressed(os.path.join(HERE, "update.npz"), **out)


if __name__ == "__main__":
    main()
