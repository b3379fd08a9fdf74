"""Generate the golden vectors under tests/golden/ from the REFERENCE'S OWN compiled sources.

Run in the build container only (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

It imports oracle/_ref/pyci_ref (the unmodified reference sources compiled against oracle/shim) and,
in a second process, oracle/_ref/pyci_ref_gencifix (the same with the two GenCI loop bounds
sparseop.cpp:453,476 reading nvir_up).  Outputs:

  small.npz      full CSR arrays / RDM tensors / matvec results for small systems
  digests.json   nnz + sha256 of (indptr, indices, data) + E0 for the larger systems (configs 1-2 ...)

Nothing here runs on the GPU box; the committed outputs are what the tests read.
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
DATA = "/root/reference/pyci/test/data/"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def e0_of(op, ecore):
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    n = op.shape[0]
    low = sp.csr_matrix((op.data(), op.indices(), op.indptr()), shape=op.shape)
    full = low + sp.tril(low, -1).T
    if n <= 64:
        return float(np.linalg.eigvalsh(full.toarray())[0] + ecore)
    return float(sla.eigsh(full, k=1, which="SA", tol=1e-13, ncv=min(40, n - 1))[0][0] + ecore)


def seeded_vec(n, seed):
    return np.random.default_rng(seed).standard_normal(n)


def main_reference():
    import pyci_ref as pyci
    from oracle import oracle as O

    small = {}
    digests = {}

    def put_csr(tag, op):
        small[tag + ".indptr"] = op.indptr()
        small[tag + ".indices"] = op.indices()
        small[tag + ".data"] = op.data()

    def put_digest(tag, op, **extra):
        d = dict(nnz=int(op.size), shape=list(op.shape), indptr=sha(op.indptr()), indices=sha(op.indices()),
                 data=sha(op.data()), data_abs_sum=float(np.abs(op.data()).sum()))
        d.update(extra)
        digests[tag] = d

    # ---- small systems: whole arrays
    for fn, kind, occ in [("h4_sto3g", "fullci", (2, 2)), ("lih_sto6g", "fullci", (2, 2)),
                          ("BH_sto-3g_eq", "fullci", (3, 3)), ("h6_sto_3g", "fullci", (4, 2)),
                          ("be_ccpvdz", "doci", (2, 2)), ("h2_sto3g", "fullci", (1, 1))]:
        ham = pyci.secondquant_op(DATA + fn + ".fcidump")
        wfn = getattr(pyci, kind + "_wfn")(ham.nbasis, *occ)
        wfn.add_all_dets()
        tag = f"{fn}.{kind}{occ[0]}{occ[1]}"
        small[tag + ".dets"] = wfn.to_det_array()
        op = pyci.sparse_op(ham, wfn)
        put_csr(tag + ".sym", op)
        x = seeded_vec(len(wfn), 11)
        small[tag + ".sym.y"] = op(x)
        small[tag + ".E0"] = np.array(e0_of(op, ham.ecore))
        opn = pyci.sparse_op(ham, wfn, symmetric=False)
        put_csr(tag + ".nonsym", opn)
        small[tag + ".nonsym.y"] = opn(x)
        if len(wfn) > 20:
            opr = pyci.sparse_op(ham, wfn, len(wfn) - 10, symmetric=False)
            put_csr(tag + ".rect", opr)
            small[tag + ".rect.y"] = opr(x)
        c = seeded_vec(len(wfn), 12)
        c /= np.linalg.norm(c)
        r1, r2 = pyci.compute_rdms(wfn, c)
        small[tag + ".rdm1"] = r1
        small[tag + ".rdm2"] = r2

    # ---- a selected (incomplete, shuffled) FullCI space with nocc_up != nocc_dn, synthetic integrals
    n = 8
    ec, one, two = O.synthetic_integrals(n, 1234)
    ham = pyci.secondquant_op(ec, one, two)
    dets = O.all_dets(O.FULLCI, n, 3, 2)
    sel = np.random.default_rng(5).permutation(len(dets))[:600]
    sd = np.ascontiguousarray(dets[sel])
    wfn = pyci.fullci_wfn(n, 3, 2, sd)
    small["syn8.fullci32.sel.dets"] = sd
    put_csr("syn8.fullci32.sel.sym", pyci.sparse_op(ham, wfn))
    put_csr("syn8.fullci32.sel.nonsym", pyci.sparse_op(ham, wfn, symmetric=False))
    c = seeded_vec(len(wfn), 12)
    c /= np.linalg.norm(c)
    r1, r2 = pyci.compute_rdms(wfn, c)
    small["syn8.fullci32.sel.rdm1"] = r1
    small["syn8.fullci32.sel.rdm2"] = r2

    # ---- larger systems: digests (configs 1 and 2 of BASELINE.json, Li2 DOCI, synthetic 4a4b, multiword)
    for fn, kind, occ in [("be_ccpvdz", "fullci", (2, 2)), ("h2o_ccpvdz", "doci", (5, 5)),
                          ("li2_ccpvdz", "doci", (3, 3))]:
        ham = pyci.secondquant_op(DATA + fn + ".fcidump")
        wfn = getattr(pyci, kind + "_wfn")(ham.nbasis, *occ)
        wfn.add_all_dets()
        tag = f"{fn}.{kind}{occ[0]}{occ[1]}"
        op = pyci.sparse_op(ham, wfn)
        x = seeded_vec(len(wfn), 11)
        put_digest(tag + ".sym", op, E0=e0_of(op, ham.ecore), ndet=len(wfn), y_sha=sha(op(x)),
                   y_norm=float(np.linalg.norm(op(x))))
        opn = pyci.sparse_op(ham, wfn, symmetric=False)
        put_digest(tag + ".nonsym", opn, ndet=len(wfn), y_norm=float(np.linalg.norm(opn(x))))
        c = seeded_vec(len(wfn), 12)
        c /= np.linalg.norm(c)
        r1, r2 = pyci.compute_rdms(wfn, c)
        digests[tag + ".rdm"] = dict(rdm1=sha(r1), rdm2=sha(r2), rdm1_sum=float(r1.sum()),
                                     rdm2_abs_sum=float(np.abs(r2).sum()))
    for n, occ in [(10, (4, 4)), (9, (4, 3))]:
        ec, one, two = O.synthetic_integrals(n, 1234)
        ham = pyci.secondquant_op(ec, one, two)
        wfn = pyci.fullci_wfn(n, *occ)
        wfn.add_all_dets()
        tag = f"syn{n}.fullci{occ[0]}{occ[1]}"
        op = pyci.sparse_op(ham, wfn)
        put_digest(tag + ".sym", op, E0=e0_of(op, ham.ecore), ndet=len(wfn))
        put_digest(tag + ".nonsym", pyci.sparse_op(ham, wfn, symmetric=False), ndet=len(wfn))
    n = 66
    ec, one, two = O.synthetic_integrals(n, 7)
    ham = pyci.secondquant_op(ec, one, two)
    wfn = pyci.doci_wfn(n, 2, 2)
    wfn.add_all_dets()
    put_digest("syn66.doci22.sym", pyci.sparse_op(ham, wfn), ndet=len(wfn))

    np.savez_compressed(os.path.join(HERE, "small.npz"), **small)
    return digests


def main_gencifix():
    import pyci_ref_gencifix as pyci
    from oracle import oracle as O

    out = {}
    for fn, occ in [("h4_sto3g", (2, 2)), ("BH_sto-3g_eq", (3, 3)), ("h6_sto_3g", (4, 2))]:
        ecore, one, two = O.read_fcidump(DATA + fn + ".fcidump")
        n = one.shape[0]
        fd = O.all_dets(O.FULLCI, n, *occ)
        gd = (fd[:, 0, :] | (fd[:, 1, :] << np.uint64(n))).astype(np.uint64)
        h2, g2 = O.spin_orbital_integrals(one, two)
        ham = pyci.secondquant_op(ecore, h2, g2)
        wfn = pyci.genci_wfn(2 * n, sum(occ), 0, gd)
        tag = f"{fn}.genci{sum(occ)}"
        out[tag + ".dets"] = gd
        for sym, name in ((True, "sym"), (False, "nonsym")):
            op = pyci.sparse_op(ham, wfn, symmetric=sym)
            out[f"{tag}.{name}.indptr"] = op.indptr()
            out[f"{tag}.{name}.indices"] = op.indices()
            out[f"{tag}.{name}.data"] = op.data()
    np.savez_compressed(os.path.join(HERE, "genci.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--gencifix":
        main_gencifix()
    else:
        dig = main_reference()
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--gencifix"])
        with open(os.path.join(HERE, "digests.json"), "w") as f:
            json.dump(dig, f, indent=1, sort_keys=True)
        print("wrote", os.listdir(HERE))
