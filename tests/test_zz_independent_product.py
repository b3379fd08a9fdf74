"""Full-size products and energies against an INDEPENDENT CPU algorithm: the string-driven direct-CI product of
tests/golden/make_golden_e0_direct.py (one-spin Hamiltonian over the alpha strings + alpha-beta term from
single-excitation lists and one dense product; it shares no row code with the product or with the oracle, and it
reproduces the oracle-built matrices' energies).  Fixtures: tests/golden/spmv_direct.npz (y = H x of a seeded x as 4096
sampled entries, |y| and x.y) and the `syn16` entry of tests/golden/e0_syn.json.

The file sorts last on purpose: these are the only checks of the device at config 3 / config 4 size that do not go
through the oracle's row code."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, seeded_vec
from oracle import oracle as O

TOL = 1e-10  # relative to |y| (entries), to |y| (norm) and to |x||y| (x.y): rounding of a 3193-term fp64 row sum is ~1e-14


@pytest.fixture(scope="module")
def fixture():
    with np.load(os.path.join(GOLDEN, "spmv_direct.npz")) as f:
        return {k: f[k] for k in f.files}


def compare(y, x, fx, key):
    idx, yref = fx[key + ".idx"], fx[key + ".y"]
    ynorm, xy, xnorm = float(fx[key + ".norm"]), float(fx[key + ".xdoty"]), float(fx[key + ".xnorm"])
    assert abs(np.linalg.norm(x) - xnorm) <= 1e-12 * xnorm  # the same seeded x
    assert np.max(np.abs(y[idx] - yref)) <= TOL * np.abs(yref).max()
    assert abs(np.linalg.norm(y) - ynorm) <= TOL * ynorm
    assert abs(float(x @ y) - xy) <= TOL * xnorm * ynorm


def test_fixture_against_the_oracle_matrix(fixture):
    """CPU: the syn10 entry of the fixture equals the oracle-built FullCI(10, 4a4b) matrix applied to the same x."""
    n, occ = 10, (4, 4)
    _, one, two = O.synthetic_integrals(n, 1234)
    dets = O.all_dets(O.FULLCI, n, *occ)
    ip, ix, dv = O.sparse_op(O.FULLCI, n, occ[0], occ[1], dets, (one, two), symmetric=True)
    x = seeded_vec(len(dets), 8)
    compare(O.full_symmetric(ip, ix, dv, len(dets)) @ x, x, fixture, "syn10")


@pytest.fixture(scope="module")
def cabi():
    from pyci_b200 import cabi as C
    assert C.lib().pyci_device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return C


@pytest.fixture(scope="module")
def ctx(cabi):
    c = cabi.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n,key", [(14, "syn14"), (16, "syn16")])
def test_full_size_product_and_energy_against_direct_ci(cabi, ctx, fixture, n, key):
    """BASELINE config 3 (n = 14) and config 4 on one GPU (n = 16, 127 GB): op(x) of the whole operator against the
    direct-CI product, and E0 against the direct-CI + ARPACK energy (config 3's matrix-based golden equals it)."""
    import torch
    if n == 16 and torch.cuda.mem_get_info(0)[1] < 170e9:
        pytest.skip("needs a 180 GB device")
    occ = (4, 4)
    ecore, one, two = O.synthetic_integrals(n, 1234)
    ham = cabi.Ham(ctx, n, ecore, one, two)
    w = cabi.Wfn(ctx, cabi.FULLCI, n, occ[0], occ[1])  # generated on the device in add_all_dets order
    op = cabi.Op(ctx, ham, w)
    nd = op.nrow
    assert nd == {14: 1002001, 16: 3312400}[n] and fixture[key + ".idx"][-1] == nd - 1
    x = seeded_vec(nd, 8)
    compare(op.matvec(x), x, fixture, key)
    es = op.solve(n=1, tol=1e-9)[0]
    with open(os.path.join(GOLDEN, "e0_syn.json")) as f:
        g = json.load(f)[key]
    e0 = g.get("direct_ci_E0", g["E0"])
    assert g["ndet"] == nd and abs(es[0] - (e0 + ecore)) <= 1e-10, (es[0], e0)
    op.close()
    w.close()
    ham.close()
