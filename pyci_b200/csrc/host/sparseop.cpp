// sparse_op / compute_rdms of pyci_b200._pyci: numpy marshalling around the C ABI, with the method
// names, defaults and error behaviour of the reference's SparseOp
// (/root/reference/pyci/src/sparseop.cpp:49-178,504-514; binding.cpp:884-1088) and py_compute_rdms_*
// (rdm.cpp:1011-1055).  All numerical work happens in libpyci_b200.so on the GPU.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "pyci_host.h"

namespace pyci_host {

namespace {

pyci_ctx *g_ctx = nullptr;

// Page-locked buffers behind the numpy arrays the export methods return (the row pointer: 8 MB at a million rows).
// A copy from the device into a fresh pageable array pays the page faults of the array and the driver's bounce
// buffers; into page-locked memory it runs at the PCIe rate.  A page-locked allocation costs about a millisecond, so
// the buffers are recycled: the capsule that owns an array's memory hands it back here when numpy drops the array.
// (Never destroyed: buffers still cached at interpreter exit are left to the process teardown.)
struct PinnedPool {
    static constexpr size_t MAX_CACHED = (size_t)256 << 20, MAX_ONE = (size_t)64 << 20;
    std::mutex m;
    std::vector<std::pair<size_t, void *>> cache;
    size_t cached = 0;
    void *get(size_t bytes, size_t *cap) {
        size_t c = 4096;
        while (c < bytes)
            c <<= 1;
        *cap = c;
        {
            std::lock_guard<std::mutex> lock(m);
            for (size_t i = 0; i < cache.size(); ++i)
                if (cache[i].first == c) {
                    void *p = cache[i].second;
                    cache[i] = cache.back();
                    cache.pop_back();
                    cached -= c;
                    return p;
                }
        }
        void *p = nullptr;
        return pyci_host_alloc(&p, c) == PYCI_OK ? p : nullptr;
    }
    void put(void *p, size_t cap) {
        {
            std::lock_guard<std::mutex> lock(m);
            if (cached + cap <= MAX_CACHED) {
                cache.emplace_back(cap, p);
                cached += cap;
                return;
            }
        }
        pyci_host_free(p);
    }
};
PinnedPool &pinned_pool() {
    static PinnedPool *pool = new PinnedPool;
    return *pool;
}
struct PinnedBlock {
    void *p;
    size_t cap;
};

// a 1-d int64 array of n elements over page-locked memory (plain numpy memory when it is large or unavailable)
Array<long> pinned_long_array(long n) {
    const size_t bytes = sizeof(long) * (size_t)std::max<long>(n, 1);
    if (bytes <= PinnedPool::MAX_ONE && !std::getenv("PYCI_B200_NO_PINNED_RESULTS")) {
        size_t cap = 0;
        if (void *p = pinned_pool().get(bytes, &cap)) {
            py::capsule owner(new PinnedBlock{p, cap}, [](void *q) {
                PinnedBlock *b = static_cast<PinnedBlock *>(q);
                pinned_pool().put(b->p, b->cap);
                delete b;
            });
            return py::array_t<long>({(py::ssize_t)n}, {(py::ssize_t)sizeof(long)}, static_cast<long *>(p), owner);
        }
    }
    return Array<long>(n);
}

struct DeviceHam {
    pyci_ham *h = nullptr;
    DeviceHam(const SQuantOp &ham) {
        check(pyci_ham_upload(device_context(), ham.nbasis, ham.ecore, ham.one_mo(), ham.two_mo(),
                              ham.h_array.data(), ham.v_array.data(), ham.w_array.data(), &h));
    }
    ~DeviceHam() { pyci_ham_destroy(h); }
};

struct DeviceWfn {
    pyci_wfn *w = nullptr;
    DeviceWfn(const Wfn &wfn) {
        if (wfn.dict.size() != wfn.ndet)
            throw std::invalid_argument("wave function contains duplicate determinants");
        if (wfn.full_space && !std::getenv("PYCI_B200_UPLOAD_DETS")) {
            // the contents are add_all_dets' output: the device generates them (no determinant crosses PCIe)
            check(pyci_wfn_create_all_dets(device_context(), wfn.kind(), wfn.nbasis, wfn.nocc_up, wfn.nocc_dn, &w));
            return;
        }
        check(pyci_wfn_upload(device_context(), wfn.kind(), wfn.nbasis, wfn.nocc_up, wfn.nocc_dn, wfn.ndet,
                              reinterpret_cast<const uint64_t *>(wfn.dets.data()), &w));
    }
    ~DeviceWfn() { pyci_wfn_destroy(w); }
};

} // namespace

void check(int status) {
    if (status == PYCI_OK)
        return;
    const std::string msg = pyci_last_error();
    switch (status) {
    case PYCI_ERR_VALUE:
        throw std::invalid_argument(msg);
    case PYCI_ERR_TYPE:
        throw py::type_error(msg);
    case PYCI_ERR_MEMORY:
        throw std::bad_alloc();
    default:
        throw std::runtime_error(msg);
    }
}

pyci_ctx *device_context() {
    if (!g_ctx) {
        int device = 0;
        const char *env = std::getenv("PYCI_B200_DEVICE");
        if (!env)
            env = std::getenv("LOCAL_RANK");
        if (env)
            device = std::atoi(env);
        check(pyci_ctx_create(device, nullptr, &g_ctx));
    }
    return g_ctx;
}

void set_device_context(int device, uintptr_t stream) {
    // sparse_op objects built so far keep a pointer to the context they were built on (their kernels, frees and
    // destructors run there): a replaced context is retired, not destroyed -- a few hundred bytes and one stream per
    // set_device() call, against a use-after-free in every live operator
    static std::vector<pyci_ctx *> retired;
    pyci_ctx *fresh = nullptr;
    check(pyci_ctx_create(device, reinterpret_cast<void *>(stream), &fresh));
    if (g_ctx)
        retired.push_back(g_ctx);
    g_ctx = fresh;
}

SparseOp::SparseOp(const SQuantOp &ham, const Wfn &wfn, long rows, long cols, bool symm)
    : nrow((rows > -1) ? rows : wfn.ndet), ncol((cols > -1) ? cols : wfn.ndet), ecore(ham.ecore), symmetric(symm) {
    build(ham, wfn, nrow, ncol);
}

SparseOp::~SparseOp() { pyci_op_destroy(handle); }

void SparseOp::build(const SQuantOp &ham, const Wfn &wfn, long rows, long cols) {
    DeviceHam dham(ham);
    DeviceWfn dwfn(wfn);
    pyci_op *fresh = nullptr;
    pyci_ctx *ctx = device_context();
    int rc;
    {
        py::gil_scoped_release nogil;
        rc = pyci_op_build(ctx, dham.h, dwfn.w, rows, cols, symmetric ? 1 : 0, &fresh);
    }
    check(rc);
    pyci_op_destroy(handle);
    handle = fresh;
    nrow = rows;
    ncol = cols;
    ecore = ham.ecore;
    shape = py::make_tuple(py::cast(nrow), py::cast(ncol));
}

// SparseOp::py_update (sparseop.cpp:175-178): extend to all determinants now in wfn.  The reference appends rows
// [old nrow, ndet) to its storage and leaves the rows it has alone.  Symmetric: the device keeps full rows, so
// pyci_op_update builds the new rows and transposes their old-column entries into the old rows (only the new
// determinants are enumerated).  Non-symmetric: the new rows are appended, the old ones keep the columns they were
// built with, like the reference's.  Row-sharded operators are rebuilt (the uniform partition moves rows between
// ranks): identical for symmetric operators; a non-symmetric one then has complete rows where the reference has the
// old ones.
void SparseOp::update(const SQuantOp &ham, const Wfn &wfn) {
    if (handle && (!symmetric || nrow == ncol) && wfn.ndet >= nrow && pyci_ctx_nranks(device_context()) == 1) {
        DeviceHam dham(ham);
        DeviceWfn dwfn(wfn);
        int rc;
        {
            py::gil_scoped_release nogil;
            rc = pyci_op_update(handle, dham.h, dwfn.w);
        }
        if (rc == PYCI_OK) {
            nrow = ncol = wfn.ndet;
            ecore = ham.ecore;
                    shape = py::make_tuple(py::cast(nrow), py::cast(ncol));
            return;
        }
        if (rc != PYCI_ERR_UNSUPPORTED)
            check(rc);
    }
    build(ham, wfn, wfn.ndet, wfn.ndet);
}

double SparseOp::get_element(long i, long j) const {
    if (i < 0 || i >= nrow || j < 0 || j >= ncol)
        throw py::index_error("matrix index out of range");
    double v = 0.0;
    check(pyci_op_get_element(handle, i, j, &v));
    return v;
}

void SparseOp::perform_op(const double *x, double *y) const { check(pyci_op_matvec(handle, x, y)); }

Array<double> SparseOp::py_matvec(const Array<double> x) const {
    Array<double> y(nrow);
    return py_matvec_out(x, y);
}

Array<double> SparseOp::py_matvec_out(const Array<double> x, Array<double> y) const {
    if ((long)x.size() < ncol)
        throw std::invalid_argument("x has fewer elements than the operator has columns");
    if ((long)y.size() < nrow)
        throw std::invalid_argument("out has fewer elements than the operator has rows");
    const double *xp = x.data();
    double *yp = y.mutable_data();
    int rc;
    {
        py::gil_scoped_release nogil;
        rc = pyci_op_matvec(handle, xp, yp);
    }
    check(rc);
    return y;
}

py::tuple SparseOp::py_solve_ci(long n, py::object c0, long ncv, long maxiter, double tol) {
    if (n < 1)
        throw std::invalid_argument("cannot find >=n eigenpairs for sparse operator with n rows");
    Array<double> eigvals(n);
    Array<double> eigvecs({n, nrow});
    Array<double> guess;
    const double *cptr = nullptr;
    if (!c0.is_none()) {
        guess = c0.cast<Array<double>>();
        if ((long)guess.size() < nrow)
            throw std::invalid_argument("c0 has fewer elements than the operator has rows");
        cptr = guess.data();
    }
    double *ev = eigvals.mutable_data(), *ec = eigvecs.mutable_data();
    int rc;
    {
        py::gil_scoped_release nogil;
        rc = pyci_op_solve(handle, n, cptr, ncv, maxiter, tol, ev, ec, &last_stats);
    }
    check(rc);
    return py::make_tuple(eigvals, eigvecs);
}

Array<long> SparseOp::py_indptr() const {
    Array<long> a = pinned_long_array(pyci_op_row_count(handle) + 1);
    check(pyci_op_export_csr(handle, a.mutable_data(), nullptr, nullptr));
    return a;
}

Array<long> SparseOp::py_indices() const {
    std::vector<long> ptr((size_t)pyci_op_row_count(handle) + 1);
    Array<long> a(pyci_op_size(handle));
    check(pyci_op_export_csr(handle, ptr.data(), a.mutable_data(), nullptr));
    return a;
}

Array<double> SparseOp::py_data() const {
    std::vector<long> ptr((size_t)pyci_op_row_count(handle) + 1);
    Array<double> a(pyci_op_size(handle));
    check(pyci_op_export_csr(handle, ptr.data(), nullptr, a.mutable_data()));
    return a;
}

py::dict SparseOp::py_stats() const {
    py::dict d;
    double t[4];
    pyci_op_build_times(handle, t);
    d["hash_seconds"] = t[0];
    d["count_seconds"] = t[1];
    d["fill_seconds"] = t[2];
    d["build_seconds"] = t[3];
    d["fill_kernel"] = std::string(pyci_op_fill_kernel(handle));
    d["stored_nnz"] = pyci_op_stored_nnz(handle);
    d["row_begin"] = pyci_op_row_begin(handle);
    d["row_count"] = pyci_op_row_count(handle);
    d["matvecs"] = last_stats.matvecs;
    d["iterations"] = last_stats.iterations;
    d["restarts"] = last_stats.restarts;
    d["residual"] = last_stats.residual;
    d["solve_seconds"] = last_stats.seconds;
    d["spmv_seconds"] = last_stats.spmv_seconds;
    return d;
}

// py_compute_rdms_{doci,fullci,genci} (rdm.cpp:1011-1055): output shapes follow the reference
py::tuple py_compute_rdms(const Wfn &wfn, const Array<double> coeffs) {
    if ((long)coeffs.size() < wfn.ndet)
        throw std::invalid_argument("coeffs has fewer elements than the wave function has determinants");
    const long n = wfn.nbasis;
    Array<double> r1, r2;
    switch (wfn.kind()) {
    case PYCI_DOCI:
        r1 = Array<double>({n, n});
        r2 = Array<double>({n, n});
        break;
    case PYCI_FULLCI:
        r1 = Array<double>({2L, n, n});
        r2 = Array<double>({3L, n, n, n, n});
        break;
    default:
        r1 = Array<double>({n, n});
        r2 = Array<double>({n, n, n, n});
        break;
    }
    DeviceWfn dwfn(wfn);
    const double *cp = coeffs.data();
    double *p1 = r1.mutable_data(), *p2 = r2.mutable_data();
    pyci_ctx *ctx = device_context();
    int rc;
    {
        py::gil_scoped_release nogil;
        rc = pyci_compute_rdms(ctx, dwfn.w, cp, p1, p2);
    }
    check(rc);
    return py::make_tuple(r1, r2);
}

// py_compute_transition_rdms_{doci,fullci,genci} (rdm.cpp:1057-1088): shapes as compute_rdms
py::tuple py_compute_transition_rdms(const Wfn &wfn1, const Wfn &wfn2, const Array<double> coeffs1,
                                     const Array<double> coeffs2) {
    if ((long)coeffs1.size() < wfn1.ndet || (long)coeffs2.size() < wfn2.ndet)
        throw std::invalid_argument("coeffs has fewer elements than the wave function has determinants");
    const long n = wfn1.nbasis;
    Array<double> r1, r2;
    switch (wfn1.kind()) {
    case PYCI_DOCI:
        r1 = Array<double>({n, n});
        r2 = Array<double>({n, n});
        break;
    case PYCI_FULLCI:
        r1 = Array<double>({2L, n, n});
        r2 = Array<double>({3L, n, n, n, n});
        break;
    default:
        r1 = Array<double>({n, n});
        r2 = Array<double>({n, n, n, n});
        break;
    }
    DeviceWfn d1(wfn1), d2(wfn2);
    const double *c1 = coeffs1.data(), *c2 = coeffs2.data();
    double *p1 = r1.mutable_data(), *p2 = r2.mutable_data();
    pyci_ctx *ctx = device_context();
    int rc;
    {
        py::gil_scoped_release nogil;
        rc = pyci_compute_transition_rdms(ctx, d1.w, d2.w, c1, c2, p1, p2);
    }
    check(rc);
    return py::make_tuple(r1, r2);
}

// py_compute_overlap (overlap.cpp:60-76)
double py_compute_overlap(const Wfn &wfn1, const Wfn &wfn2, const Array<double> coeffs1, const Array<double> coeffs2) {
    if ((long)coeffs1.size() < wfn1.ndet || (long)coeffs2.size() < wfn2.ndet)
        throw std::invalid_argument("coeffs has fewer elements than the wave function has determinants");
    DeviceWfn d1(wfn1), d2(wfn2);
    const double *c1 = coeffs1.data(), *c2 = coeffs2.data();
    pyci_ctx *ctx = device_context();
    double out = 0.0;
    int rc;
    {
        py::gil_scoped_release nogil;
        rc = pyci_compute_overlap(ctx, d1.w, d2.w, c1, c2, &out);
    }
    check(rc);
    return out;
}

// py_add_hci (hci.cpp:282-301; binding.cpp:1147-1181): one heat-bath iteration on the device; the selected
// determinants are read back and appended to the host wave function so that it stays the single owner of
// the determinant list.  `nthread` is accepted for signature parity and ignored.
long py_add_hci(const SQuantOp &ham, Wfn &wfn, const Array<double> coeffs, double eps, long) {
    if ((long)coeffs.size() < wfn.ndet)
        throw std::invalid_argument("coeffs has fewer elements than the wave function has determinants");
    if (wfn.ndet == 0)
        return 0;
    DeviceHam dham(ham);
    DeviceWfn dwfn(wfn);
    const double *cp = coeffs.data();
    pyci_ctx *ctx = device_context();
    long nnew = 0;
    std::vector<ulong> fresh;
    int rc;
    {
        py::gil_scoped_release nogil;
        rc = pyci_wfn_add_hci(ctx, dham.h, dwfn.w, cp, eps, &nnew);
        if (rc == PYCI_OK && nnew > 0) {
            fresh.resize((size_t)(nnew * wfn.nw));
            rc = pyci_wfn_download_dets(dwfn.w, wfn.ndet, nnew, reinterpret_cast<uint64_t *>(fresh.data()));
        }
    }
    check(rc);
    const long before = wfn.ndet;
    wfn.append_new_dets(fresh.data(), nnew); // distinct and absent by construction: no per-determinant look-up
    return wfn.ndet - before;
}

// py_compute_enpt2 (enpt2.cpp:382-398; binding.cpp:1344-1375); DOCI goes through its FullCI image like the
// reference (enpt2.cpp:376-380)
double py_compute_enpt2(const SQuantOp &ham, const Wfn &wfn, const Array<double> coeffs, double energy, double eps,
                        long) {
    if ((long)coeffs.size() < wfn.ndet)
        throw std::invalid_argument("coeffs has fewer elements than the wave function has determinants");
    if (wfn.kind() == PYCI_DOCI) {
        const FullCIWfn image(static_cast<const DOCIWfn &>(wfn));
        return py_compute_enpt2(ham, image, coeffs, energy, eps, -1);
    }
    DeviceHam dham(ham);
    DeviceWfn dwfn(wfn);
    const double *cp = coeffs.data();
    pyci_ctx *ctx = device_context();
    double out = energy;
    int rc;
    {
        py::gil_scoped_release nogil;
        rc = pyci_compute_enpt2(ctx, dham.h, dwfn.w, cp, energy, eps, &out, nullptr);
    }
    check(rc);
    return out;
}

} // namespace pyci_host
