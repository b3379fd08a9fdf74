r"""pyci_b200: B200-native (sm_100a) implementation of PyCI's CI-Hamiltonian hot path.

Drop-in for the part of ``pyci`` that builds the sparse CI matrix, applies it, finds its lowest
eigenpairs and contracts reduced density matrices::

    import pyci_b200 as pyci
    ham = pyci.hamiltonian("be_ccpvdz.fcidump")
    wfn = pyci.fullci_wfn(ham.nbasis, 2, 2); wfn.add_all_dets()
    op = pyci.sparse_op(ham, wfn)            # built on the GPU (CSR stays in HBM)
    es, cs = op.solve(n=1, tol=1e-6)         # block Davidson on the GPU
    d1, d2 = pyci.compute_rdms(wfn, cs[0])

The names mirror ``pyci/__init__.py`` of the reference (``/root/reference/pyci/__init__.py:18-100``) for
everything on the path.  There is no CPU fallback: the extension modules must be built
(``python -c "import __graft_entry__ as g; g.build()"``) and compute calls need a CUDA device.
"""
import os as _os

try:
    from pyci_b200 import _pyci
except ImportError as _exc:  # fail loudly: the product is the native code
    raise ImportError(
        "pyci_b200._pyci (pybind11 host module) or libpyci_b200.so (CUDA library) is not built or not "
        "loadable: run `make -C %s` first. Original error: %s"
        % (_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "csrc"), _exc)
    ) from _exc

from pyci_b200._pyci import __version__, c_long, c_ulong, c_double
from pyci_b200._pyci import secondquant_op, wavefunction, one_spin_wfn, two_spin_wfn
from pyci_b200._pyci import doci_wfn, fullci_wfn, genci_wfn, sparse_op
from pyci_b200._pyci import get_num_threads, set_num_threads, popcnt, ctz
from pyci_b200._pyci import compute_rdms, add_hci, compute_enpt2, compute_transition_rdms, compute_overlap
from pyci_b200._pyci import device_count, set_device, nccl_unique_id, init_comm
from pyci_b200._pyci import launch_count, reset_launch_count, synchronize, release_memory

from pyci_b200.utility import make_senzero_integrals, reduce_senzero_integrals, spinize_rdms, spin_free_rdms
from pyci_b200.utility import add_excitations
from pyci_b200.selectors import add_seniorities, add_gkci, add_cost, odometer_one_spin, odometer_two_spin

# Alias kept by the reference for compatibility with old versions (pyci/__init__.py:100)
hamiltonian = secondquant_op

__all__ = [
    "__version__", "c_long", "c_ulong", "c_double",
    "secondquant_op", "hamiltonian", "wavefunction", "one_spin_wfn", "two_spin_wfn",
    "doci_wfn", "fullci_wfn", "genci_wfn", "sparse_op",
    "get_num_threads", "set_num_threads", "popcnt", "ctz", "compute_rdms", "add_hci", "compute_enpt2", "compute_transition_rdms", "compute_overlap",
    "make_senzero_integrals", "reduce_senzero_integrals", "spinize_rdms", "spin_free_rdms", "add_excitations",
    "add_seniorities", "add_gkci", "add_cost", "odometer_one_spin", "odometer_two_spin",
    "device_count", "set_device", "nccl_unique_id", "init_comm", "launch_count", "reset_launch_count",
    "synchronize", "release_memory",
]
