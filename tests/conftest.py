"""Shared fixtures.  `-m "not gpu"` runs here on CPU; `-m gpu` runs on a B200 box where /root/reference
does not exist: nothing below reads it."""
import gzip
import hashlib
import json
import os
import shutil
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATA = os.path.join(ROOT, "tests", "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")
_TMP = {}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def datafile(name):
    """Path of tests/data/<name>.fcidump, gunzipping `<name>.fcidump.gz` into a temp dir on first use."""
    plain = os.path.join(DATA, name + ".fcidump")
    if os.path.exists(plain):
        return plain
    if name not in _TMP:
        import tempfile
        d = tempfile.mkdtemp(prefix="pyci_b200_data_")
        out = os.path.join(d, name + ".fcidump")
        with gzip.open(plain + ".gz", "rb") as src, open(out, "wb") as dst:
            shutil.copyfileobj(src, dst)
        _TMP[name] = out
    return _TMP[name]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def seeded_vec(n, seed):
    return np.random.default_rng(seed).standard_normal(n)


@pytest.fixture(scope="session")
def small():
    with np.load(os.path.join(GOLDEN, "small.npz")) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def genci_golden():
    with np.load(os.path.join(GOLDEN, "genci.npz")) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def digests():
    with open(os.path.join(GOLDEN, "digests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def hci_golden():
    with np.load(os.path.join(GOLDEN, "hci.npz")) as f:
        return {k: f[k] for k in f.files}


HCI_CASES = [("be_fullci", "be_ccpvdz", "fullci", (2, 2), 3), ("h6_fullci", "h6_sto_3g", "fullci", (3, 3), 3),
             ("lih_fullci", "lih_sto6g", "fullci", (2, 2), 3), ("li2_doci", "li2_ccpvdz", "doci", (3, 3), 3),
             ("be_doci", "be_ccpvdz", "doci", (2, 2), 2)]


def sorted_rows(d):
    """determinant array sorted lexicographically by its words (order-free comparison of determinant sets)"""
    if d.shape[0] == 0:
        return d
    flat = d.reshape(d.shape[0], -1)
    return d[np.lexsort(flat.T[::-1])]


@pytest.fixture(scope="session")
def trdm_golden():
    with np.load(os.path.join(GOLDEN, "trdm.npz")) as f:
        return {k: f[k] for k in f.files}


TRDM_CASES = [("h6_fullci", "fullci", 6, (3, 3)), ("lih_fullci", "fullci", 6, (2, 1)), ("be_doci", "doci", 14, (2, 2)),
              ("h4_fullci", "fullci", 4, (2, 2))]


@pytest.fixture(scope="session")
def update_golden():
    """SparseOp::update applied by the compiled reference (tests/golden/make_golden_update.py)."""
    with np.load(os.path.join(GOLDEN, "update.npz")) as f:
        return {k: f[k] for k in f.files}


UPDATE_CASES = [("be.fullci22", "be_ccpvdz", "fullci", (2, 2)), ("be.doci22", "be_ccpvdz", "doci", (2, 2))]
UPDATE_MODES = [("sym", True), ("nonsym", False), ("rect", False)]

