// pybind11 module pyci_b200._pyci: the reference-facing surface of the B200 hot path.  Class, method
// and keyword names, defaults and exception types follow /root/reference/pyci/src/binding.cpp:29-1291
// for everything on or next to the path (secondquant_op, the wave-function classes, sparse_op,
// compute_rdms, thread/bit helpers); callers outside the path (add_hci, compute_enpt2, FanCI
// objectives, transition RDMs) are not provided by this build.
#include <pybind11/stl.h>

#include "pyci_host.h"

namespace py = pybind11;
using namespace pyci_host;

namespace {

template<class W, class Base>
void bind_spin_wfn(py::class_<W, Base> &c) {
    c.def("__getitem__", [](const W &w, long i) { return w.py_getitem(i); }, py::arg("index"),
          "Return the determinant at position ``index`` as an array of unsigned words.");
    c.def("to_file", [](const W &w, const std::string &f) { w.to_file(f); }, py::arg("filename"),
          "Write the wave function to a binary file.");
    c.def("to_det_array", [](const W &w, long lo, long hi) { return w.py_to_det_array(lo, hi); },
          py::arg("low") = -1, py::arg("high") = -1, "Return the determinant bit-strings as an array.");
    c.def("to_occ_array", [](const W &w, long lo, long hi) { return w.py_to_occ_array(lo, hi); },
          py::arg("low") = -1, py::arg("high") = -1, "Return the occupied-orbital indices as an array.");
    c.def("index_det", [](const W &w, const Array<ulong> d) { return w.py_index_det(d); }, py::arg("det"),
          "Return the index of a determinant, or -1 if it is not in the wave function.");
    c.def("index_det_from_rank", [](const W &w, const Hash r) { return w.index_det_from_rank(r); },
          py::arg("rank"), "Return the index of the determinant with the given rank, or -1.");
    c.def("rank_det", [](const W &w, const Array<ulong> d) { return w.py_rank_det(d); }, py::arg("det"),
          "Return the rank (128-bit hash) of a determinant.");
    c.def("add_det", [](W &w, const Array<ulong> d) { return w.py_add_det(d); }, py::arg("det"),
          "Add a determinant; returns its index, or -1 if it was already present.");
    c.def("add_occs", [](W &w, const Array<long> o) { return w.py_add_occs(o); }, py::arg("occs"),
          "Add a determinant from occupied-orbital indices; returns its index or -1.");
    c.def("add_hartreefock_det", [](W &w) { w.add_hartreefock_det(); }, "Add the Hartree-Fock determinant.");
    c.def("add_all_dets", [](W &w, long nthread) { w.add_all_dets(nthread); }, py::arg("nthread") = -1,
          "Replace the contents with all determinants of the full space.");
    c.def("add_excited_dets", [](W &w, long exc, py::object ref) { return w.py_add_excited_dets(exc, ref); },
          py::arg("exc"), py::arg("ref") = py::none(),
          "Add all determinants of excitation level ``exc`` from ``ref`` (default Hartree-Fock).");
    c.def("add_dets_from_wfn", [](W &w, const W &o) { w.add_dets_from(o); }, py::arg("wfn"),
          "Add the determinants of another wave function.");
    c.def("reserve", [](W &w, long n) { w.reserve(n); }, py::arg("n"), "Reserve space for ``n`` determinants.");
    c.def("_append_new_dets", [](W &w, const Array<ulong> a) {
        if (a.size() % w.nw)
            throw std::invalid_argument("array size is not a multiple of the words per determinant");
        w.append_new_dets(a.data(), (long)(a.size() / w.nw));
    }, py::arg("dets"), "Bulk append of determinants known to be distinct and absent (what add_hci uses for the "
                        "determinants selected on the device); no duplicate check.");
}

} // namespace

PYBIND11_MODULE(_pyci, m) {
    m.doc() = "pyci_b200._pyci: B200-native CI Hamiltonian hot path behind the PyCI C extension surface.";
    m.attr("__version__") = "0.6.1+b200";
    m.attr("c_long") = py::dtype::of<long>();
    m.attr("c_ulong") = py::dtype::of<ulong>();
    m.attr("c_double") = py::dtype::of<double>();

    // ---- secondquant_op (binding.cpp:58-206)
    py::class_<SQuantOp> ham(m, "secondquant_op", "Second-quantized operator (Hamiltonian) class.");
    ham.def_readonly("nbasis", &SQuantOp::nbasis);
    ham.def_readonly("ecore", &SQuantOp::ecore);
    ham.def_readonly("one_mo", &SQuantOp::one_mo_array);
    ham.def_readonly("two_mo", &SQuantOp::two_mo_array);
    ham.def_readonly("h", &SQuantOp::h_array);
    ham.def_readonly("v", &SQuantOp::v_array);
    ham.def_readonly("w", &SQuantOp::w_array);
    ham.def(py::init<const std::string &>(), py::arg("filename"));
    ham.def(py::init<const double, const Array<double>, const Array<double>>(), py::arg("ecore"), py::arg("one_mo"),
            py::arg("two_mo"));
    ham.def("to_file", &SQuantOp::to_file, py::arg("filename"), py::arg("nelec") = 0, py::arg("ms2") = 0,
            py::arg("tol") = 0.0);

    // ---- wave functions (binding.cpp:212-878)
    py::class_<Wfn> wfn(m, "wavefunction", "Wave function base class.");
    wfn.def_readonly("nbasis", &Wfn::nbasis);
    wfn.def_readonly("nocc", &Wfn::nocc);
    wfn.def_readonly("nocc_up", &Wfn::nocc_up);
    wfn.def_readonly("nocc_dn", &Wfn::nocc_dn);
    wfn.def_readonly("nvir", &Wfn::nvir);
    wfn.def_readonly("nvir_up", &Wfn::nvir_up);
    wfn.def_readonly("nvir_dn", &Wfn::nvir_dn);
    wfn.def("__len__", &Wfn::length);
    wfn.def("squeeze", &Wfn::squeeze, "Free any unused memory allocated to this object.");

    py::class_<OneSpinWfn, Wfn> one(m, "one_spin_wfn", "One-spin wave function base class.");
    bind_spin_wfn(one);
    py::class_<TwoSpinWfn, Wfn> two(m, "two_spin_wfn", "Two-spin wave function base class.");
    bind_spin_wfn(two);

    py::class_<DOCIWfn, OneSpinWfn> doci(m, "doci_wfn", "DOCI wave function class.");
    doci.def(py::init<const DOCIWfn &>(), py::arg("wfn"));
    doci.def(py::init<const std::string &>(), py::arg("filename"));
    doci.def(py::init<const long, const long, const long>(), py::arg("nbasis"), py::arg("nocc_up"), py::arg("nocc_dn"));
    doci.def(py::init<const long, const long, const long, const Array<ulong>>(), py::arg("nbasis"),
             py::arg("nocc_up"), py::arg("nocc_dn"), py::arg("array"));
    doci.def(py::init<const long, const long, const long, const Array<long>>(), py::arg("nbasis"), py::arg("nocc_up"),
             py::arg("nocc_dn"), py::arg("array"));

    py::class_<FullCIWfn, TwoSpinWfn> fullci(m, "fullci_wfn", "FullCI wave function class.");
    fullci.def(py::init<const DOCIWfn &>(), py::arg("wfn"));
    fullci.def(py::init<const FullCIWfn &>(), py::arg("wfn"));
    fullci.def(py::init<const std::string &>(), py::arg("filename"));
    fullci.def(py::init<const long, const long, const long>(), py::arg("nbasis"), py::arg("nocc_up"),
               py::arg("nocc_dn"));
    fullci.def(py::init<const long, const long, const long, const Array<ulong>>(), py::arg("nbasis"),
               py::arg("nocc_up"), py::arg("nocc_dn"), py::arg("array"));
    fullci.def(py::init<const long, const long, const long, const Array<long>>(), py::arg("nbasis"),
               py::arg("nocc_up"), py::arg("nocc_dn"), py::arg("array"));

    py::class_<GenCIWfn, OneSpinWfn> genci(m, "genci_wfn", "Generalized CI wave function class.");
    genci.def(py::init<const DOCIWfn &>(), py::arg("wfn"));
    genci.def(py::init<const FullCIWfn &>(), py::arg("wfn"));
    genci.def(py::init<const GenCIWfn &>(), py::arg("wfn"));
    genci.def(py::init<const std::string &>(), py::arg("filename"));
    genci.def(py::init<const long, const long, const long>(), py::arg("nbasis"), py::arg("nocc_up"), py::arg("nocc_dn"));
    genci.def(py::init<const long, const long, const long, const Array<ulong>>(), py::arg("nbasis"),
              py::arg("nocc_up"), py::arg("nocc_dn"), py::arg("array"));
    genci.def(py::init<const long, const long, const long, const Array<long>>(), py::arg("nbasis"),
              py::arg("nocc_up"), py::arg("nocc_dn"), py::arg("array"));

    // ---- sparse_op (binding.cpp:884-1088)
    py::class_<SparseOp> op(m, "sparse_op", "Sparse matrix operator class (CSR resident in GPU memory).");
    op.def_readonly("ecore", &SparseOp::ecore);
    op.def_readonly("symmetric", &SparseOp::symmetric);
    op.def_property_readonly("size", &SparseOp::size);
    op.def_readonly("shape", &SparseOp::shape);
    op.def_property_readonly("dtype", &SparseOp::dtype);
    op.def(py::init([](const SQuantOp &h, const DOCIWfn &w, long r, long c, bool s) { return new SparseOp(h, w, r, c, s); }),
           py::arg("ham"), py::arg("wfn"), py::arg("nrow") = -1, py::arg("ncol") = -1, py::arg("symmetric") = true);
    op.def(py::init([](const SQuantOp &h, const FullCIWfn &w, long r, long c, bool s) { return new SparseOp(h, w, r, c, s); }),
           py::arg("ham"), py::arg("wfn"), py::arg("nrow") = -1, py::arg("ncol") = -1, py::arg("symmetric") = true);
    op.def(py::init([](const SQuantOp &h, const GenCIWfn &w, long r, long c, bool s) { return new SparseOp(h, w, r, c, s); }),
           py::arg("ham"), py::arg("wfn"), py::arg("nrow") = -1, py::arg("ncol") = -1, py::arg("symmetric") = true);
    op.def("update", [](SparseOp &o, const SQuantOp &h, const DOCIWfn &w) { o.update(h, w); }, py::arg("ham"), py::arg("wfn"));
    op.def("update", [](SparseOp &o, const SQuantOp &h, const FullCIWfn &w) { o.update(h, w); }, py::arg("ham"), py::arg("wfn"));
    op.def("update", [](SparseOp &o, const SQuantOp &h, const GenCIWfn &w) { o.update(h, w); }, py::arg("ham"), py::arg("wfn"));
    op.def("__call__", &SparseOp::py_matvec, py::arg("x"));
    op.def("__call__", &SparseOp::py_matvec_out, py::arg("x"), py::arg("out"));
    op.def("matvec", &SparseOp::py_matvec, py::arg("x"));
    op.def("matvec", &SparseOp::py_matvec_out, py::arg("x"), py::arg("out"));
    op.def("get_element", &SparseOp::get_element, py::arg("i"), py::arg("j"));
    op.def("solve", &SparseOp::py_solve_ci, py::arg("n") = 1, py::arg("c0") = py::none(), py::arg("ncv") = -1,
           py::arg("maxiter") = -1, py::arg("tol") = 1.0e-12);
    op.def("reserve", &SparseOp::reserve, py::arg("n"));
    op.def("squeeze", &SparseOp::squeeze);
    op.def("data", &SparseOp::py_data, "Return CSR matrix data vector");
    op.def("indices", &SparseOp::py_indices, "Return CSR matrix indices vector");
    op.def("indptr", &SparseOp::py_indptr, "Return CSR matrix index pointer vector");
    op.def("time_matvec", [](SparseOp &o, int warmup, int reps, long flush_bytes) {
        std::vector<double> ms((size_t)std::max(reps, 1));
        int rc;
        {
            py::gil_scoped_release nogil;
            rc = pyci_op_time_spmv(o.handle, warmup, reps, flush_bytes, ms.data());
        }
        check(rc);
        return ms;
    }, py::arg("warmup") = 3, py::arg("reps") = 10, py::arg("flush_bytes") = 0,
           "Per-launch device milliseconds of the SpMV kernel (CUDA events on the launching stream).");
    op.def("stats", &SparseOp::py_stats, "Device timings and counters of the last build / solve (pyci_b200 extension).");

    // ---- free functions (binding.cpp:1094-1145, 1208-1291)
    m.def("get_num_threads", &get_num_threads);
    m.def("set_num_threads", &set_num_threads, py::arg("n"));
    m.def("popcnt", [](const Array<ulong> det) { return popcnt_det(det.size(), det.data()); }, py::arg("det"));
    m.def("ctz", [](const Array<ulong> det) { return ctz_det(det.size(), det.data()); }, py::arg("det"));
    m.def("compute_rdms", [](const DOCIWfn &w, const Array<double> c) { return py_compute_rdms(w, c); }, py::arg("wfn"),
          py::arg("coeffs"));
    m.def("compute_rdms", [](const FullCIWfn &w, const Array<double> c) { return py_compute_rdms(w, c); },
          py::arg("wfn"), py::arg("coeffs"));
    m.def("compute_rdms", [](const GenCIWfn &w, const Array<double> c) { return py_compute_rdms(w, c); },
          py::arg("wfn"), py::arg("coeffs"));

    // ---- pyci_b200 extensions: device context, row sharding, launch accounting
    // compute_transition_rdms / compute_overlap (binding.cpp:1183-1206, 1293-1342)
    m.def("compute_transition_rdms", [](const DOCIWfn &a, const DOCIWfn &b, const Array<double> c1, const Array<double> c2) { return py_compute_transition_rdms(a, b, c1, c2); },
          py::arg("wfn1"), py::arg("wfn2"), py::arg("coeffs1"), py::arg("coeffs2"),
          "Compute the transition one- and two- particle reduced density matrices of two wave functions (on the GPU).");
    m.def("compute_transition_rdms", [](const FullCIWfn &a, const FullCIWfn &b, const Array<double> c1, const Array<double> c2) { return py_compute_transition_rdms(a, b, c1, c2); },
          py::arg("wfn1"), py::arg("wfn2"), py::arg("coeffs1"), py::arg("coeffs2"));
    m.def("compute_transition_rdms", [](const GenCIWfn &a, const GenCIWfn &b, const Array<double> c1, const Array<double> c2) { return py_compute_transition_rdms(a, b, c1, c2); },
          py::arg("wfn1"), py::arg("wfn2"), py::arg("coeffs1"), py::arg("coeffs2"));
    m.def("compute_overlap", [](const OneSpinWfn &a, const OneSpinWfn &b, const Array<double> c1, const Array<double> c2) { return py_compute_overlap(a, b, c1, c2); },
          py::arg("wfn1"), py::arg("wfn2"), py::arg("coeffs1"), py::arg("coeffs2"),
          "Compute the overlap of two wave functions (on the GPU).");
    m.def("compute_overlap", [](const TwoSpinWfn &a, const TwoSpinWfn &b, const Array<double> c1, const Array<double> c2) { return py_compute_overlap(a, b, c1, c2); },
          py::arg("wfn1"), py::arg("wfn2"), py::arg("coeffs1"), py::arg("coeffs2"));
    // add_hci / compute_enpt2 (binding.cpp:1147-1181, 1344-1375): same overloads, keywords and defaults
    m.def("add_hci", [](const SQuantOp &h, DOCIWfn &w, const Array<double> c, double eps, long nt) { return py_add_hci(h, w, c, eps, nt); },
          py::arg("ham"), py::arg("wfn"), py::arg("coeffs"), py::arg("eps") = 1.0e-5, py::arg("nthread") = -1,
          "Add determinants to a wave function by running an iteration of Heat-Bath CI (on the GPU).");
    m.def("add_hci", [](const SQuantOp &h, FullCIWfn &w, const Array<double> c, double eps, long nt) { return py_add_hci(h, w, c, eps, nt); },
          py::arg("ham"), py::arg("wfn"), py::arg("coeffs"), py::arg("eps") = 1.0e-5, py::arg("nthread") = -1);
    m.def("add_hci", [](const SQuantOp &h, GenCIWfn &w, const Array<double> c, double eps, long nt) { return py_add_hci(h, w, c, eps, nt); },
          py::arg("ham"), py::arg("wfn"), py::arg("coeffs"), py::arg("eps") = 1.0e-5, py::arg("nthread") = -1);
    m.def("compute_enpt2", [](const SQuantOp &h, const DOCIWfn &w, const Array<double> c, double e, double eps, long nt) { return py_compute_enpt2(h, w, c, e, eps, nt); },
          py::arg("ham"), py::arg("wfn"), py::arg("coeffs"), py::arg("energy"), py::arg("eps") = 1.0e-5, py::arg("nthread") = -1,
          "Compute the second-order Epstein-Nesbet perturbation theory correction to the energy (on the GPU).");
    m.def("compute_enpt2", [](const SQuantOp &h, const FullCIWfn &w, const Array<double> c, double e, double eps, long nt) { return py_compute_enpt2(h, w, c, e, eps, nt); },
          py::arg("ham"), py::arg("wfn"), py::arg("coeffs"), py::arg("energy"), py::arg("eps") = 1.0e-5, py::arg("nthread") = -1);
    m.def("compute_enpt2", [](const SQuantOp &h, const GenCIWfn &w, const Array<double> c, double e, double eps, long nt) { return py_compute_enpt2(h, w, c, e, eps, nt); },
          py::arg("ham"), py::arg("wfn"), py::arg("coeffs"), py::arg("energy"), py::arg("eps") = 1.0e-5, py::arg("nthread") = -1);
    m.def("device_count", []() { return pyci_device_count(); });
    m.def("set_device", [](int device, uintptr_t stream) { set_device_context(device, stream); }, py::arg("device"),
          py::arg("stream") = 0, "Bind this process to a CUDA device (and optionally an existing cudaStream_t).");
    m.def("nccl_unique_id", []() {
        char id[128];
        check(pyci_nccl_unique_id(id));
        return py::bytes(id, 128);
    });
    m.def("init_comm", [](int rank, int nranks, py::bytes id) {
        std::string s = id;
        if (nranks > 1 && s.size() != 128)
            throw std::invalid_argument("NCCL unique id must be 128 bytes");
        check(pyci_ctx_init_comm(device_context(), rank, nranks, s.data()));
    }, py::arg("rank"), py::arg("nranks"), py::arg("unique_id"));
    m.def("launch_count", []() { return pyci_ctx_launch_count(device_context()); });
    m.def("reset_launch_count", []() { pyci_ctx_reset_launch_count(device_context()); });
    m.def("synchronize", []() { check(pyci_ctx_synchronize(device_context())); });
    m.def("release_memory", []() { check(pyci_ctx_release_memory(device_context())); },
          "Return the device memory the library's pool holds but does not use to the driver.");

    const char *env = std::getenv("PYCI_NUM_THREADS");
    if (env)
        set_num_threads(std::atol(env));
}
