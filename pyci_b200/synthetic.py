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


def seniority_zero_genci_dets(nspatial, npair, ndet=None):
    r"""A deterministic *selected* GenCI space for config 5: the closed-shell (seniority-zero) determinants of
    ``npair`` electron pairs in ``nspatial`` spatial orbitals, written as strings over ``2*nspatial <= 64``
    spin-orbitals (alpha block first, so spatial orbital p is spin-orbitals p and p + nspatial), in colex
    order of the pair string and truncated to the first ``ndet``.  Inside the full spin-orbital space it is a
    sparse selection: of the O(10^5) single and double excitations of a row only the ~npair*(nspatial-npair)
    pair excitations stay in the space -- the regime of heat-bath / selected CI, where construction is bound
    by index probes that miss.  Returns uint64[ndet, 1]."""
    import math
    if 2 * nspatial > 64:
        raise ValueError("2 * nspatial must be <= 64")
    total = math.comb(nspatial, npair)
    ndet = total if ndet is None else min(ndet, total)
    out = _colex_unrank_masks(nspatial, npair, ndet)
    return ((out << np.uint64(nspatial)) | out).reshape(-1, 1)


def _colex_unrank_masks(n, k, count, chunk=1 << 22):
    r"""Bit masks of the k-subsets of range(n) with colex ranks 0..count-1 (combinatorial number system:
    the j-th largest element is the largest c with C(c, j) <= remaining rank), vectorised over the ranks."""
    import math
    out = np.empty(count, dtype=np.uint64)
    tables = [np.array([math.comb(c, j) for c in range(n + 1)], dtype=np.int64) for j in range(k + 1)]
    for lo in range(0, count, chunk):
        r = np.arange(lo, min(count, lo + chunk), dtype=np.int64)
        mask = np.zeros(len(r), dtype=np.uint64)
        for j in range(k, 0, -1):
            c = np.searchsorted(tables[j], r, side="right") - 1
            r -= tables[j][c]
            mask |= np.uint64(1) << c.astype(np.uint64)
        out[lo:lo + len(mask)] = mask
    return out
