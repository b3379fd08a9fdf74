"""The drop-in boundary on a box WITHOUT a GPU: libpyci_b200.so loads, exports every function that
include/pyci_b200.h declares, reports errors the way the header says, and has no CPU fallback."""
import ctypes
import os
import re

import numpy as np
import pytest

from pyci_b200 import cabi


def declared_functions():
    text = open(cabi.HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"PYCI_API\s+[\w\s\*]+?\b(pyci_\w+)\s*\(", text)))


def test_header_and_ctypes_table_agree():
    names = declared_functions()
    assert len(names) >= 30
    assert sorted(cabi.PROTOTYPES) == names


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(cabi.LIBRARY)
    for name in declared_functions():
        assert hasattr(L, name), name
    assert cabi.lib().pyci_abi_version() == 1


def test_only_c_abi_symbols_are_exported():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", cabi.LIBRARY], capture_output=True, text=True).stdout
    exported = [ln.split()[-1] for ln in out.splitlines() if " T " in ln]
    foreign = [s for s in exported if not s.startswith("pyci_") and s not in ("_init", "_fini")]
    assert not foreign, foreign[:10]


def test_built_for_sm_100a_only():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", cabi.LIBRARY], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(cabi.lib().pyci_device_count() > 0, reason="checks the no-GPU behaviour")
def test_compute_entry_points_fail_loudly_without_a_device():
    with pytest.raises(cabi.PyciError) as e:
        cabi.Context(0)
    assert e.value.status == cabi.ERR_CUDA
    assert "no CPU fallback" in str(e.value)
    import pyci_b200 as pyci
    ham = pyci.hamiltonian(0.0, np.eye(2), np.zeros((2, 2, 2, 2)))
    wfn = pyci.doci_wfn(2, 1, 1)
    wfn.add_all_dets()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pyci.sparse_op(ham, wfn)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pyci.compute_rdms(wfn, np.ones(len(wfn)))
    # the selected-CI, transition-RDM and overlap entry points are device-only as well
    c = np.ones(len(wfn))
    for call in (lambda: pyci.add_hci(ham, wfn, c), lambda: pyci.compute_enpt2(ham, wfn, c, 0.0),
                 lambda: pyci.compute_transition_rdms(wfn, wfn, c, c), lambda: pyci.compute_overlap(wfn, wfn, c, c)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
    assert len(wfn) == 2  # a failed add_hci leaves the host wave function untouched


def test_product_package_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "pyci_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|liboracle|oracle/_ref|pyci_ref", text, flags=re.M):
                    bad.append(fn)
    assert not bad, bad
