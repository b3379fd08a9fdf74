// Host side of pyci_b200._pyci: C++ objects with the same public surface as the reference's
// SQuantOp / Wfn hierarchy / SparseOp (/root/reference/pyci/include/pyci.h:285-693), re-implemented
// here.  They own the host data (integrals, determinant list, determinant dictionary) and feed the
// CUDA library through the C ABI of include/pyci_b200.h; no compute happens on this side.
#pragma once

#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <cstdint>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "pyci_b200.h"

namespace pyci_host {

namespace py = pybind11;

typedef unsigned long ulong;
typedef std::pair<ulong, ulong> Hash;

template<typename T>
using Array = py::array_t<T, py::array::c_style | py::array::forcecast>;

// ---- bit-string helpers (semantics of common.cpp) -------------------------------------------------
long binomial(long n, long k);
long nword_det(long nbasis);
void fill_det(long nocc, const long *occs, ulong *det);
void fill_hartreefock_det(long nocc, ulong *det);
void fill_occs(long nword, const ulong *det, long *occs);
void fill_virs(long nword, long nbasis, const ulong *det, long *virs);
void next_colex(long *indices);
void unrank_colex(long nbasis, long nocc, long rank, long *occs);
long popcnt_det(long nword, const ulong *det);
long ctz_det(long nword, const ulong *det);
void excite_det(long i, long a, ulong *det);
long get_num_threads();
void set_num_threads(long n);

// SpookyHash V2 (Bob Jenkins, public domain) 128-bit digest with PyCI's seeds (pyci.h:123-128)
Hash spooky_rank(const ulong *words, long nwords);

// ---- determinant dictionary: bit-string -> position ----------------------------------------------
class DetTable {
public:
    void reset(long nw) {
        nw_ = nw;
        slots_.clear();
        count_ = 0;
    }
    void reserve(long n);
    // position of `key` in dets, or -1
    long find(const std::vector<ulong> &dets, const ulong *key) const;
    // map key -> idx, overwriting an existing mapping (dict[rank] = i); returns true if it was new
    bool assign(const std::vector<ulong> &dets, const ulong *key, long idx);
    // map determinants [first, first + n) of dets, known to be distinct and absent from the table (the device's
    // add_hci guarantees both), to their positions: no key comparisons, slots prefetched a block ahead
    void insert_new_bulk(const std::vector<ulong> &dets, long first, long n);
    // assign() for determinants [first, first + n) in order (duplicates allowed: the later position wins), with
    // the home slots prefetched a block ahead
    void assign_bulk(const std::vector<ulong> &dets, long first, long n);
    long size() const { return count_; }

private:
    long nw_ = 1;
    long count_ = 0;
    std::vector<long> slots_;
    void grow(const std::vector<ulong> &dets, long newcap);
    static uint64_t mix(const ulong *k, long nw);
};

// ---- device context (one per process) ---------------------------------------------------------------
pyci_ctx *device_context();
void set_device_context(int device, uintptr_t stream);
void check(int status); // maps C-ABI status codes to the reference's Python exception types

// ---- SQuantOp ------------------------------------------------------------------------------------------
struct SQuantOp {
    long nbasis = 0;
    double ecore = 0.0;
    Array<double> one_mo_array, two_mo_array, h_array, v_array, w_array;

    SQuantOp(const std::string &filename);
    SQuantOp(double ecore, const Array<double> one_mo, const Array<double> two_mo);
    void to_file(const std::string &filename, long nelec, long ms2, double tol) const;

    const double *one_mo() const { return one_mo_array.data(); }
    const double *two_mo() const { return two_mo_array.data(); }

private:
    void derive_senzero();
};

// ---- wave functions -----------------------------------------------------------------------------------
struct Wfn {
    long nbasis = 0, nocc = 0, nocc_up = 0, nocc_dn = 0, nvir = 0, nvir_up = 0, nvir_dn = 0;
    long ndet = 0, nword = 0, nword2 = 0, maxrank_up = 0, maxrank_dn = 0;
    int nspin = 1; // strings per determinant: 1 (DOCI, GenCI) or 2 (FullCI)
    long nw = 1;   // words per determinant = nspin * nword
    std::vector<ulong> dets;
    DetTable dict;
    // true while the contents are exactly what add_all_dets produced (nothing added since): such a wave function is
    // defined by (nbasis, nocc_up, nocc_dn), and the device unranks its determinants itself instead of receiving them
    bool full_space = false;

    virtual ~Wfn() = default;
    virtual int kind() const = 0;

    long length() const { return ndet; }
    void squeeze() { dets.shrink_to_fit(); }

    void init(long nb, long nu, long nd, int nspin_);
    void load_file(const std::string &filename, int nspin_);
    void set_dets(long n, const ulong *ptr);
    void set_new_dets(long n, const ulong *ptr);
    void set_occs(long n, const long *ptr);

    const ulong *det_ptr(long i) const { return &dets[i * nw]; }
    void to_file(const std::string &filename) const;
    long index_det(const ulong *det) const { return dict.find(dets, det); }
    long index_det_from_rank(const Hash rank) const;
    Hash rank_det(const ulong *det) const { return spooky_rank(det, nw); }
    long add_det(const ulong *det);
    void append_new_dets(const ulong *ptr, long n); // distinct determinants that are not in the wave function yet
    long add_det_from_occs(const long *occs);
    void add_hartreefock_det();
    void add_all_dets(long nthread);
    void add_dets_from(const Wfn &other);
    void reserve(long n);

    // python-facing helpers shared by the one- and two-spin classes
    py::array py_getitem(long index) const;
    py::array py_to_det_array(long low, long high) const;
    py::array py_to_occ_array(long low, long high) const;
    long py_index_det(const Array<ulong> det) const;
    Hash py_rank_det(const Array<ulong> det) const;
    long py_add_det(const Array<ulong> det);
    long py_add_occs(const Array<long> occs);
    long py_add_excited_dets(long exc, const py::object ref);

protected:
    void onespin_excited(const ulong *rdet, long e, long nocc_s, std::vector<ulong> &out) const;
    void check_det_arg(const Array<ulong> &det) const;
    mutable std::map<Hash, long> rank_index_; // lazily built for index_det_from_rank
    mutable long rank_index_ndet_ = -1;
};

struct OneSpinWfn : Wfn {};
struct TwoSpinWfn : Wfn {};

struct DOCIWfn final : OneSpinWfn {
    int kind() const override { return PYCI_DOCI; }
    DOCIWfn(const DOCIWfn &) = default;
    DOCIWfn(const std::string &filename);
    DOCIWfn(long nb, long nu, long nd);
    DOCIWfn(long nb, long nu, long nd, const Array<ulong> array);
    DOCIWfn(long nb, long nu, long nd, const Array<long> array);
};

struct FullCIWfn final : TwoSpinWfn {
    int kind() const override { return PYCI_FULLCI; }
    FullCIWfn(const FullCIWfn &) = default;
    FullCIWfn(const DOCIWfn &);
    FullCIWfn(const std::string &filename);
    FullCIWfn(long nb, long nu, long nd);
    FullCIWfn(long nb, long nu, long nd, const Array<ulong> array);
    FullCIWfn(long nb, long nu, long nd, const Array<long> array);
};

struct GenCIWfn final : OneSpinWfn {
    int kind() const override { return PYCI_GENCI; }
    GenCIWfn(const GenCIWfn &) = default;
    GenCIWfn(const DOCIWfn &);
    GenCIWfn(const FullCIWfn &);
    GenCIWfn(const std::string &filename);
    GenCIWfn(long nb, long nu, long nd);
    GenCIWfn(long nb, long nu, long nd, const Array<ulong> array);
    GenCIWfn(long nb, long nu, long nd, const Array<long> array);
};

// ---- SparseOp -----------------------------------------------------------------------------------------
struct SparseOp {
    long nrow = 0, ncol = 0;
    double ecore = 0.0;
    bool symmetric = true;
    py::tuple shape;
    pyci_op *handle = nullptr;
    pyci_solve_stats last_stats{};

    SparseOp(const SQuantOp &ham, const Wfn &wfn, long rows, long cols, bool symm);
    ~SparseOp();
    SparseOp(const SparseOp &) = delete;
    SparseOp &operator=(const SparseOp &) = delete;

    void build(const SQuantOp &ham, const Wfn &wfn, long rows, long cols);
    void update(const SQuantOp &ham, const Wfn &wfn);
    py::object dtype() const { return py::dtype::of<double>(); }
    long size() const { return handle ? pyci_op_size(handle) : 0; } // SparseOp::size, summed on the device at first use
    double get_element(long i, long j) const;
    // SparseOp::perform_op / perform_op_symm (sparseop.cpp:96-112): y = A x on raw host pointers, the entry the reference's
    // C++ callers use (FanCI objectives, fanci.cpp:144,196).  Full rows are stored, so both are the same product.
    void perform_op(const double *x, double *y) const;
    void perform_op_symm(const double *x, double *y) const { perform_op(x, y); }
    Array<double> py_matvec(const Array<double> x) const;
    Array<double> py_matvec_out(const Array<double> x, Array<double> y) const;
    py::tuple py_solve_ci(long n, py::object c0, long ncv, long maxiter, double tol);
    void reserve(long) {}
    void squeeze() {}
    Array<double> py_data() const;
    Array<long> py_indices() const;
    Array<long> py_indptr() const;
    py::dict py_stats() const;
};

py::tuple py_compute_rdms(const Wfn &wfn, const Array<double> coeffs);
py::tuple py_compute_transition_rdms(const Wfn &wfn1, const Wfn &wfn2, const Array<double> coeffs1,
                                     const Array<double> coeffs2);
double py_compute_overlap(const Wfn &wfn1, const Wfn &wfn2, const Array<double> coeffs1, const Array<double> coeffs2);
long py_add_hci(const SQuantOp &ham, Wfn &wfn, const Array<double> coeffs, double eps, long nthread);
double py_compute_enpt2(const SQuantOp &ham, const Wfn &wfn, const Array<double> coeffs, double energy, double eps,
                        long nthread);

} // namespace pyci_host
