r"""Determinant selectors of the reference's Python layer: the callers that fill a wave function before
``sparse_op`` is built from it.  Pure host code over the public wave-function surface (``add_occs``, ``add_det``,
``to_det_array``); nothing here touches the device.

Mirrors, with the same names, arguments, insertion order and error types:

* ``add_seniorities``           ``/root/reference/pyci/seniority_ci.py:30-96``
* ``odometer_one_spin/_two_spin`` ``/root/reference/pyci/utility.py:424-505``
* ``add_gkci`` and its node models ``/root/reference/pyci/gkci.py:30-227``
* ``add_cost``                  ``/root/reference/pyci/cost_ci.py:28-55`` — the reference passes ``q_max=`` to functions
  whose parameter is ``qmax`` and raises ``TypeError`` in this snapshot; here the call it means to make is made.

Pinned by ``tests/golden/selectors.npz`` (the reference's own functions driving its compiled classes,
``tests/golden/make_golden_selectors.py``): determinant arrays equal, order included."""
from itertools import combinations, product

import numpy as np
from scipy.special import gammaln, polygamma

from pyci_b200 import _pyci

__all__ = ["add_seniorities", "odometer_one_spin", "odometer_two_spin", "add_gkci", "add_cost",
           "compute_nodes_cntsp", "compute_nodes_gamma", "compute_nodes_interval"]


def add_seniorities(wfn, *seniorities):
    r"""Add every determinant with the given numbers of unpaired electrons to a FullCI wave function.

    For each seniority the up-strings run in ``doci_wfn.add_all_dets`` order; the down-string takes ``pairs =
    (nocc - s) / 2`` orbitals of the up-string and the rest outside it, both in ``itertools.combinations`` order."""
    if not isinstance(wfn, _pyci.fullci_wfn):
        raise TypeError(f"invalid `wfn` type `{type(wfn)}`; must be `pyci.fullci_wfn`")
    nup, ndn, n = wfn.nocc_up, wfn.nocc_dn, wfn.nbasis
    lowest = nup - ndn
    highest = n - nup + ndn  # the reference's bound (seniority_ci.py:49-51), kept as it is
    for s in seniorities:
        if s < lowest or s > highest or (s - lowest) % 2:
            raise ValueError(f"invalid seniority number in `seniorities = {seniorities}`")
    strings = _pyci.doci_wfn(n, nup, nup)
    strings.add_all_dets()
    ups = strings.to_occ_array()
    occs = np.empty((2, nup), dtype=_pyci.c_long)
    for s in seniorities:
        pairs = (nup + ndn - s) // 2
        for up in ups:
            occs[0] = up
            outside = np.setdiff1d(np.arange(n, dtype=_pyci.c_long), up, assume_unique=True)
            for shared, rest in product(combinations(up, pairs), combinations(outside, ndn - pairs)):
                occs[1, :pairs] = shared
                occs[1, pairs:ndn] = rest
                wfn.add_occs(occs)


def odometer_one_spin(wfn, cost, t, qmax):
    r"""Odometer walk over the occupation lists of a one-spin wave function: a list is added when its last orbital
    exists and ``sum(cost[occ]) + t * cost[occ[-1]] < qmax``; a rejected list makes the next-lower position advance
    (the walk assumes ascending costs, like the reference's)."""
    k, n = wfn.nocc_up, wfn.nbasis
    cost = np.asarray(cost)
    occ = np.arange(k, dtype=_pyci.c_long)
    prev = occ.copy()
    pos = k - 1
    while True:
        if occ[-1] < n and (np.sum(cost[occ]) + t * cost[occ[-1]]) < qmax:
            wfn.add_occs(occ)
            pos = k - 1
        else:
            occ[:] = prev
            pos -= 1
        if pos < 0:
            return
        prev[:] = occ
        occ[pos:] = np.arange(occ[pos] + 1, occ[pos] + 1 + k - pos)


def odometer_two_spin(wfn, cost, t, qmax):
    r"""The same walk for each spin; every accepted up-string is paired with every accepted down-string (up-string
    major).  Nothing is added when either spin has no accepted string."""
    n = wfn.nbasis
    up = _pyci.doci_wfn(n, wfn.nocc_up, wfn.nocc_up)
    odometer_one_spin(up, cost, t, qmax)
    if not len(up):
        return
    ups = up.to_det_array()
    if wfn.nocc_dn:
        dn = _pyci.doci_wfn(n, wfn.nocc_dn, wfn.nocc_dn)
        odometer_one_spin(dn, cost, t, qmax)
        if not len(dn):
            return
        dns = dn.to_det_array()
    else:
        dns = np.zeros((1, ups.shape[1]), dtype=ups.dtype)
    det = np.empty((2, ups.shape[1]), dtype=ups.dtype)
    for a in ups:
        det[0] = a
        for b in dns:
            det[1] = b
            wfn.add_det(det)


def compute_nodes_cntsp(nbasis):
    r"""Nodes of hydrogen-like shells: shell ``s`` (1, 2, ...) holds ``s**2`` functions with ``s - 1`` nodes each."""
    nodes = np.zeros(nbasis)
    first, shell = 0, 1
    while first < nbasis:
        nodes[first:first + shell * shell] = shell - 1
        first += shell * shell
        shell += 1
    return nodes


def compute_nodes_gamma(nbasis, d, maxiter=100, tol=1.0e-9):
    r"""Nodes ``n_k`` solving ``Gamma(n + d + 1) / (Gamma(d + 1) Gamma(n + 1)) = k + 1`` for function ``k``, by the
    reference's iteration (gkci.py:146-171, its step formula kept so that the iterates agree), each solve started from
    the linear extrapolation of the two before it."""
    nodes = np.zeros(nbasis)
    d1 = d + 1.0
    lg_d = gammaln(d1)
    n = 0.0
    for k in range(1, nbasis):
        for _ in range(maxiter):
            lg_n, lg_nd = gammaln((n + 1, n + d1))
            psi_n, psi_nd = polygamma(0, (n + 1, n + d1))
            tri_n, tri_nd = polygamma(1, (n + 1, n + d1))
            a = np.exp(lg_nd - lg_n - lg_d)
            b = a - k - 1
            c = psi_n - psi_nd
            step = (2 * b * c) / (2 * a * c * c + b * (tri_nd - tri_n - c * (psi_n + psi_nd)))
            n += step
            if np.abs(step) < tol:
                break
        else:
            raise RuntimeError(f"Did not converge in {maxiter} iterations")
        nodes[k] = n
        n += n - nodes[k - 1]
    return nodes


def compute_nodes_interval(nbasis, es, width):
    r"""Nodes from orbital energies: intervals of ``width`` centred on each energy are merged where they overlap; a
    function's node count is the merged length below its energy in units of ``width`` (plus one half)."""
    es = np.asarray(es, dtype=float)
    half = width * 0.5
    lower, upper = es - half, es + half
    last = 0
    for k in range(1, nbasis):
        if es[k] - half < es[k - 1] + half:
            upper[last] = es[k] + half
        else:
            last += 1
            lower[last], upper[last] = es[k] - half, es[k] + half
    nodes = np.zeros(nbasis)
    for k in range(nbasis):
        for i in range(last + 1):
            if es[k] > upper[i]:
                nodes[k] += (upper[i] - lower[i]) / width
            else:
                nodes[k] += (es[k] - lower[i]) / width + 0.5
                break
    return nodes


def _odometer_for(wfn):
    if isinstance(wfn, (_pyci.doci_wfn, _pyci.genci_wfn)):
        return odometer_one_spin
    if isinstance(wfn, _pyci.fullci_wfn):
        return odometer_two_spin
    raise TypeError(f"invalid `wfn` type `{type(wfn)}`; must be `pyci.wavefunction`")


def add_gkci(wfn, t=-0.5, p=1.0, mode="cntsp", dim=3, energies=None, width=None):
    r"""Griebel-Knapek CI: odometer selection with node counts as costs.  ``mode`` names a node model over
    ``nbasis + 1`` functions ('cntsp', 'gamma' with ``dim``, 'interval' with ``energies`` and ``width``) or is the
    node array itself; the threshold is ``(sum(nodes[:nocc_up - 1]) + (t + 1) * nodes[-1]) * p``."""
    if isinstance(mode, str):
        if mode == "cntsp":
            nodes = compute_nodes_cntsp(wfn.nbasis + 1)
        elif mode == "gamma":
            nodes = compute_nodes_gamma(wfn.nbasis + 1, dim)
        elif mode == "interval":
            nodes = compute_nodes_interval(wfn.nbasis + 1, energies, width)
        else:
            raise ValueError(f"invalid `mode` value `{mode}`; must be one of ('cntsp', 'gamma', 'interval')")
    else:
        nodes = np.asarray(mode)
    qmax = (np.sum(nodes[: wfn.nocc_up - 1]) + (t + 1) * nodes[-1]) * p
    _odometer_for(wfn)(wfn, nodes, t, qmax)


def add_cost(wfn, cost, q_max, t=-0.5):
    r"""Odometer selection with the given orbital costs and threshold ``q_max``."""
    _odometer_for(wfn)(wfn, cost, t, q_max)
