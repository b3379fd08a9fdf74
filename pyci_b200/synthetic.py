r"""Synthetic workloads for the configurations BASELINE.json names without an integral file (configs 3-5):
random-symmetric integrals of a given orbital count.
Pure numpy host code; the oracle (oracle/oracle.py) restates the same integrals so both sides of a parity
test see identical inputs."""
import numpy as np

__all__ = ["synthetic_integrals", "spin_orbital_integrals"]


def synthetic_integrals(n, seed=1234):
    r"""(ecore, one_mo[n,n], two_mo[n,n,n,n]) with every element non-zero: one_mo symmetric with a spread
    diagonal (h[p,p] += p keeps the ground state well separated), two-electron integrals 0.1*N(0,1)
    symmetrised to the 8-fold symmetry of real orbitals and returned in physicist order
    ``two_mo[i,k,j,l] = <ik|jl>`` (the layout of squantop.cpp:134-141 in the reference)."""
    rng = np.random.default_rng(seed)
    h = rng.standard_normal((n, n))
    h = (h + h.T) / 2
    h[np.arange(n), np.arange(n)] += np.arange(n)
    g = 0.1 * rng.standard_normal((n, n, n, n))
    g = g + g.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    two_mo = np.ascontiguousarray(g.transpose(0, 2, 1, 3))
    return 0.0, np.ascontiguousarray(h), two_mo


def spin_orbital_integrals(one_mo, two_mo):
    r"""Spatial (n) -> spin-orbital (2n, alpha block first) integrals: GenCI(2n, N) on the result equals
    FullCI(n, na, nb) on the input."""
    n = one_mo.shape[0]
    h = np.zeros((2 * n, 2 * n))
    h[:n, :n] = one_mo
    h[n:, n:] = one_mo
    g = np.zeros((2 * n,) * 4)
    for s in (0, n):
        for t in (0, n):
            g[s:s + n, t:t + n, s:s + n, t:t + n] = two_mo
    return h, g
