"""Runs files of the reference's OWN test suite, unchanged, with `pyci` bound to pyci_b200.

Usage: python tests/reference_suite_runner.py <dir holding the reference's pyci/test> <pytest args...>

`import pyci` / `from pyci.test import datafile` / `from pyci.utility import ...` inside those files then resolve to
this package, to the suite's own `datafile` (its data/ directory) and to pyci_b200.utility.  Nothing of the oracle is
imported: this is the drop-in check of the boundary, the numbers the files assert are the reference's.
Test infrastructure only (called by tests/test_reference_suite.py in a subprocess so that the alias stays local)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(argv):
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
