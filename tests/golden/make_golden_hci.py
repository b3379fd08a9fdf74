"""Golden vectors for add_hci / compute_enpt2 from the REFERENCE'S OWN compiled sources (oracle/_ref/pyci_ref).

Run in the build container only (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden_hci.py        ->  tests/golden/hci.npz

For every case a short heat-bath trajectory is run with the reference (sparse_op -> dense ground state ->
compute_enpt2 -> add_hci).  Stored per step: the INPUTS (determinant array in the reference's order,
coefficient vector, energy, eps) and the reference's OUTPUTS (the set of appended determinants, sorted --
the reference's own append order is the iteration order of its hash map and is not reproducible -- and the
ENPT2 energy).  GenCI is absent: the reference's GenCI loops read past the virtual list (hci.cpp:205,223 use
nvir = 2*nbasis - nocc); it is covered by the GenCI == FullCI identity in the tests instead.
"""
import gzip
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
DATA = os.path.join(ROOT, "tests", "data")

CASES = [  # tag, file, wfn class, occ, eps of the trajectory, steps
    ("be_fullci", "be_ccpvdz", "fullci_wfn", (2, 2), 1.0e-3, 3),
    ("h6_fullci", "h6_sto_3g", "fullci_wfn", (3, 3), 2.0e-2, 3),
    ("lih_fullci", "lih_sto6g", "fullci_wfn", (2, 2), 1.0e-2, 3),
    ("li2_doci", "li2_ccpvdz", "doci_wfn", (3, 3), 1.0e-3, 3),
    ("be_doci", "be_ccpvdz", "doci_wfn", (2, 2), 1.0e-4, 2),
]


def datafile(name):
    plain = os.path.join(DATA, name + ".fcidump")
    if os.path.exists(plain):
        return plain
    out = os.path.join(tempfile.gettempdir(), name + ".fcidump")
    with gzip.open(plain + ".gz", "rb") as src, open(out, "wb") as dst:
        shutil.copyfileobj(src, dst)
    return out


def sorted_rows(d):
    if d.shape[0] == 0:
        return d
    flat = d.reshape(d.shape[0], -1)
    order = np.lexsort(flat.T[::-1])
    return d[order]


def main():
    import scipy.sparse as sp

    import pyci_ref as pyci

    out = {}
    for tag, fn, cls, occ, eps, steps in CASES:
        ham = pyci.secondquant_op(datafile(fn))
        wfn = getattr(pyci, cls)(ham.nbasis, *occ)
        wfn.add_hartreefock_det()
        for it in range(steps):
            op = pyci.sparse_op(ham, wfn)
            low = sp.csr_matrix((op.data(), op.indices(), op.indptr()), shape=op.shape)
            full = (low + sp.tril(low, -1).T).toarray()
            ev, evec = np.linalg.eigh(full)
            c = np.ascontiguousarray(evec[:, 0])
            e = float(ev[0] + ham.ecore)
            key = "%s.%d." % (tag, it)
            out[key + "dets"] = wfn.to_det_array()
            out[key + "coeffs"] = c
            out[key + "energy"] = np.array(e)
            out[key + "eps"] = np.array(eps)
            out[key + "enpt2"] = np.array(pyci.compute_enpt2(ham, wfn, c, e, eps))
            out[key + "enpt2_tight"] = np.array(pyci.compute_enpt2(ham, wfn, c, e, eps * 1e-2))
            before = len(wfn)
            nadd = pyci.add_hci(ham, wfn, c, eps=eps)
            new = wfn.to_det_array()[before:]
            assert new.shape[0] == nadd
            out[key + "new_sorted"] = sorted_rows(new)
            print(tag, it, before, "->", len(wfn), "E", e, "PT2", float(out[key + "enpt2"]))
    np.savez_compressed(os.path.join(HERE, "hci.npz"), **out)


if __name__ == "__main__":
    main()
