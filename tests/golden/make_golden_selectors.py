"""Golden determinant lists for the pure-Python selectors of the reference's Python layer (pyci/seniority_ci.py,
pyci/cost_ci.py, pyci/gkci.py, the odometers of pyci/utility.py), produced by THE REFERENCE'S OWN Python functions
driving the reference's own compiled wave-function classes (oracle/_ref/pyci_ref).

Run in the build container only (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden_selectors.py        ->  tests/golden/selectors.npz

The reference's modules are imported from where they lie under a stand-in package `pyci` whose `_pyci` is the compiled
reference; nothing is copied.  Stored: for each case the determinant array in insertion order (the order matters:
`to_det_array` / `to_occ_array` expose it, pyci/test/test_odometer.py:58-60)."""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PYCI_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

# (tag, wfn class, (nbasis, nocc_up, nocc_dn))
SENIORITY = [("sen.f6_3_3", (6, 3, 3), (0, 2, 4, 6)), ("sen.f6_3_2", (6, 3, 2), (1, 3, 5)), ("sen.f7_3_1", (7, 3, 1), (2, 4)),
             ("sen.f5_2_2.s2", (5, 2, 2), (2,)), ("sen.f8_4_2.s2", (8, 4, 2), (2,))]
ODOMETER = [("odo.doci8_3", "doci_wfn", (8, 3, 3), -0.5, 4.5), ("odo.genci9_4", "genci_wfn", (9, 4, 0), 0.25, 9.5),
            ("odo.fullci6_2_2", "fullci_wfn", (6, 2, 2), -0.5, 4.2), ("odo.fullci6_3_1", "fullci_wfn", (6, 3, 1), 0.0, 6.1),
            ("odo.fullci5_2_0", "fullci_wfn", (5, 2, 0), -0.5, 3.0), ("odo.doci6_2.none", "doci_wfn", (6, 2, 2), -0.5, -1.0)]
GKCI = [("gk.doci10_3.cntsp", "doci_wfn", (10, 3, 3), dict()), ("gk.fullci8_2_2.cntsp", "fullci_wfn", (8, 2, 2), dict(t=-0.5, p=1.5)),
        ("gk.genci12_4.cntsp", "genci_wfn", (12, 4, 0), dict(t=0.0, p=1.2)),
        ("gk.doci9_3.interval", "doci_wfn", (9, 3, 3), dict(mode="interval", width=0.6, p=1.3)),
        ("gk.doci9_3.nodes", "doci_wfn", (9, 3, 3), dict(mode="nodes", p=1.1)),
        ("gk.fullci7_2_1.gamma", "fullci_wfn", (7, 2, 1), dict(mode="gamma", dim=3, p=1.4))]


def costs(n, seed):
    """ascending pseudo orbital energies"""
    return np.sort(np.random.default_rng(seed).uniform(0.0, 4.0, n))


def reference_python_layer():
    import pyci_ref
    pkg = types.ModuleType("pyci")
    pkg.__path__ = [os.path.join(REF, "pyci")]
    pkg._pyci = pyci_ref._pyci
    sys.modules["pyci"] = pkg
    sys.modules["pyci._pyci"] = pyci_ref._pyci
    mods = {}
    for name in ("utility", "seniority_ci", "cost_ci", "gkci"):
        spec = importlib.util.spec_from_file_location("pyci." + name, os.path.join(REF, "pyci", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["pyci." + name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return pyci_ref, mods


def main():
    pyci, m = reference_python_layer()
    out = {}
    for tag, shape, sens in SENIORITY:
        w = pyci.fullci_wfn(*shape)
        m["seniority_ci"].add_seniorities(w, *sens)
        out[tag] = w.to_det_array()
    for tag, cls, shape, t, qmax in ODOMETER:
        w = getattr(pyci, cls)(*shape)
        # pyci/cost_ci.py:51,53 passes `q_max=` to functions whose parameter is `qmax` (TypeError in this snapshot): the
        # odometers it means to call are driven directly
        odo = m["utility"].odometer_two_spin if cls == "fullci_wfn" else m["utility"].odometer_one_spin
        odo(w, costs(shape[0], 7), t, qmax)
        out[tag] = w.to_det_array()
    for tag, cls, shape, kw in GKCI:
        w = getattr(pyci, cls)(*shape)
        kw = dict(kw)
        if kw.get("mode") == "interval":
            kw["energies"] = costs(shape[0] + 1, 11)
        if kw.get("mode") == "nodes":
            kw["mode"] = costs(shape[0] + 1, 13)
        m["gkci"].add_gkci(w, **kw)
        out[tag] = w.to_det_array()
    for k, v in out.items():
        print(k, v.shape)
    np.savez_compressed(os.path.join(HERE, "selectors.npz"), **out)


if __name__ == "__main__":
    main()
