r"""Small pure-numpy helpers that sit next to the hot path in the reference's Python layer
(``/root/reference/pyci/utility.py:24-150``, ``pyci/excitation_ci.py:21-39``): seniority-zero integral
transforms, spin-block -> spin-orbital RDM conversion, excitation-level determinant selection.
They only call the public ``_pyci`` surface."""
import numpy as np

__all__ = ["make_senzero_integrals", "reduce_senzero_integrals", "spinize_rdms", "spin_free_rdms", "add_excitations",
           "odometer_one_spin", "odometer_two_spin"]


def __getattr__(name):
    # pyci/utility.py:424-505 keeps the odometers in this namespace (`from pyci.utility import odometer_one_spin`,
    # pyci/test/test_odometer.py:21); they live in pyci_b200/selectors.py, which imports the extension module
    if name in ("odometer_one_spin", "odometer_two_spin"):
        from pyci_b200 import selectors
        return getattr(selectors, name)
    raise AttributeError(name)


def make_senzero_integrals(one_mo, two_mo):
    r"""Return the seniority-zero integrals (h, v, w) of a restricted Hamiltonian:
    ``h[p] = t_pp``, ``v[p,q] = <pp|qq>``, ``w[p,q] = 2<pq|pq> - <pq|qp>`` (two_mo in physicist order)."""
    one_mo = np.asarray(one_mo, dtype=np.double)
    two_mo = np.asarray(two_mo, dtype=np.double)
    h = np.copy(np.diagonal(one_mo))
    v = np.copy(np.diagonal(np.diagonal(two_mo)))
    w = np.diagonal(np.diagonal(two_mo, axis1=0, axis2=2), axis1=0, axis2=1) * 2 \
        - np.diagonal(np.diagonal(two_mo, axis1=0, axis2=3), axis1=0, axis2=1)
    return h, np.ascontiguousarray(v), np.ascontiguousarray(w)


def reduce_senzero_integrals(h, v, w, nocc):
    r"""Fold the one-particle seniority-zero integrals into the two-particle ones for ``nocc`` pairs, so that
    ``E = sum(rv * d0) + sum(rw * d2)`` with the DOCI matrices of ``compute_rdms``."""
    factor = 2.0 / (nocc * 2 - 1)
    rv = np.diag(h) * factor + v
    rw = (h[:, None] + h[None, :]) * factor + w
    return rv, rw


def spinize_rdms(d1, d2):
    r"""Convert DOCI matrices (d0, d2) or FullCI spin blocks (rdm1[2,n,n], rdm2[3,n,n,n,n]) to generalised
    spin-orbital RDMs (alpha orbitals first): ``rdm1[2n,2n]``, ``rdm2[2n,2n,2n,2n]`` antisymmetric."""
    d1 = np.asarray(d1)
    d2 = np.asarray(d2)
    n = d1.shape[1]
    rdm1 = np.zeros((2 * n, 2 * n), dtype=np.double)
    rdm2 = np.zeros((2 * n,) * 4, dtype=np.double)
    a, b = slice(0, n), slice(n, 2 * n)
    if d1.ndim == 2:
        p = np.arange(n)
        rdm1[a, a][p, p] = d1[p, p]
        rdm1[b, b][p, p] = d1[p, p]
        ppqq = (p[:, None], p[:, None], p[None, :], p[None, :])
        pqpq = (p[:, None], p[None, :], p[:, None], p[None, :])
        for s, t in ((a, b), (b, a)):
            blk = rdm2[s, t, s, t]
            blk[ppqq] += d1
            blk[pqpq] += d2
        for s in (a, b):
            rdm2[s, s, s, s][pqpq] += d2
        rdm2 -= np.transpose(rdm2, axes=(1, 0, 2, 3))
        rdm2 -= np.transpose(rdm2, axes=(0, 1, 3, 2))
        rdm2 *= 0.5
    else:
        rdm1[a, a] += d1[0]
        rdm1[b, b] += d1[1]
        rdm2[a, a, a, a] += d2[0]
        rdm2[b, b, b, b] += d2[1]
        rdm2[a, b, a, b] += d2[2]
        rdm2[b, a, b, a] += d2[2].transpose(1, 0, 3, 2)
        rdm2[a, b, b, a] -= d2[2].transpose(0, 1, 3, 2)
        rdm2[b, a, a, b] -= d2[2].transpose(1, 0, 2, 3)
    return rdm1, rdm2


def spin_free_rdms(d1, d2, d3=None, d4=None, d5=None, d6=None, d7=None, flag="3RDM"):
    r"""Spin-summed 1- and 2-RDM of FullCI spin blocks (``pyci/utility.py:326-421``, its FullCI branch):
    ``rdm1 = aa + bb``, ``rdm2 = aaaa + abab + baba + bbbb`` of the spin-orbital matrices of ``spinize_rdms``.
    The DOCI branch of the reference needs the 3-/4-RDM intermediates of ``compute_rdms_1234`` (out of scope here)."""
    d1 = np.asarray(d1)
    if d1.ndim == 2:
        raise NotImplementedError("spin_free_rdms of DOCI matrices needs the 3-/4-RDM terms (compute_rdms_1234), "
                                  "which pyci_b200 does not provide")
    n = d1.shape[1]
    rdm1, rdm2 = spinize_rdms(d1, d2)
    a, b = slice(0, n), slice(n, 2 * n)
    return rdm1[a, a] + rdm1[b, b], rdm2[a, a, a, a] + rdm2[a, b, a, b] + rdm2[b, a, b, a] + rdm2[b, b, b, b]


def add_excitations(wfn, *excitations, ref=None):
    r"""Add the determinants of the given excitation levels (relative to ``ref``, default Hartree-Fock)."""
    for e in excitations:
        wfn.add_excited_dets(e, ref=ref)
