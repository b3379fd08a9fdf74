"""Runs files of the reference's OWN test suite, unchanged, with `pyci` bound to pyci_b200.

Usage: python tests/reference_suite_runner.py <dir holding the reference's pyci/test> <pytest args...>
       python tests/reference_suite_runner.py --fanci <dir holding fanci_src/ and fanci_test/> <pytest args...>

With --fanci the reference's FanCI Python layer (pyci/fanci/fanci.py, detratio.py: a CALLER of the rectangular
`sparse_op(ham, wfn, nrow=nproj, ncol=len(wfn), symmetric=False)` and of `op(x, out=...)`, fanci.py:203,442,511) is
loaded from its copy as `pyci.fanci` and its own tests run on top of pyci_b200.  Only the pure-Python models load:
AP1roG / APIG (and pCCDS, whose test imports AP1roG) need the reference's C++ objective classes (fanci.cpp, out of scope).

`import pyci` / `from pyci.test import datafile` / `from pyci.utility import ...` inside those files then resolve to
this package, to the suite's own `datafile` (its data/ directory) and to pyci_b200.utility.  Nothing of the oracle is
imported: this is the drop-in check of the boundary, the numbers the files assert are the reference's.
Test infrastructure only (called by tests/test_reference_suite.py in a subprocess so that the alias stays local)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fanci_main(argv):
    import types
    base = os.path.abspath(argv[0])
    suite = os.path.join(base, "fanci_test")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, base)
    import pyci_b200
    sys.modules["pyci"] = pyci_b200
    sys.modules["pyci.utility"] = pyci_b200.utility
    pkg = types.ModuleType("pyci.fanci")           # stands in for pyci/fanci/__init__.py, which also imports the
    pkg.__path__ = [os.path.join(base, "fanci_src")]  # models built on the C++ objectives
    sys.modules["pyci.fanci"] = pkg
    pyci_b200.fanci = pkg
    pkg.FanCI = importlib.import_module("pyci.fanci.fanci").FanCI
    pkg.DetRatio = importlib.import_module("pyci.fanci.detratio").DetRatio
    tests = importlib.import_module("fanci_test")   # its __init__.py: find_datafile, assert_deriv
    sys.modules["pyci.fanci.test"] = tests
    pkg.test = tests
    import pytest
    args = ["-q", "-p", "no:cacheprovider", "-c", os.devnull, "--rootdir", suite, "-W", "ignore"]
    files = [os.path.join(suite, a) if a.split("::")[0].endswith(".py") and not os.path.isabs(a) else a for a in argv[1:]]
    return pytest.main(args + files)


def main(argv):
    if argv and argv[0] == "--fanci":
        return fanci_main(argv[1:])
    suite = os.path.abspath(argv[0])
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.dirname(suite))
    import pyci_b200
    sys.modules["pyci"] = pyci_b200
    sys.modules["pyci.utility"] = pyci_b200.utility
    pkg = importlib.import_module(os.path.basename(suite))  # the suite's __init__.py: datafile()
    sys.modules["pyci.test"] = pkg
    pyci_b200.test = pkg
    import pytest
    args = ["-q", "-p", "no:cacheprovider", "-c", os.devnull, "--rootdir", suite, "-W", "ignore::pytest.PytestUnknownMarkWarning"]
    files = [os.path.join(suite, a) if a.split("::")[0].endswith(".py") and not os.path.isabs(a) else a for a in argv[1:]]
    return pytest.main(args + files)


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
