// Wave-function classes of pyci_b200._pyci: determinant storage, dictionary and generators with the
// behaviour of the reference's Wfn / OneSpinWfn / TwoSpinWfn / DOCIWfn / FullCIWfn / GenCIWfn
// (/root/reference/pyci/src/wfn.cpp, onespinwfn.cpp, twospinwfn.cpp, dociwfn.cpp, fullciwfn.cpp,
// genciwfn.cpp).  One generic implementation parameterised by the number of strings per determinant.
#include <algorithm>
#include <cstring>
#include <fstream>

#include "pyci_host.h"

namespace pyci_host {

// wfn.cpp:51-71
void Wfn::init(long nb, long nu, long nd, int nspin_) {
    if (nd < 0)
        throw std::domain_error("nocc_dn is < 0");
    else if (nu < nd)
        throw std::domain_error("nocc_up is < nocc_dn");
    else if (nb < nu)
        throw std::domain_error("nbasis is < nocc_up");
    nbasis = nb;
    nocc = nu + nd;
    nocc_up = nu;
    nocc_dn = nd;
    nvir = nb * 2 - nu - nd;
    nvir_up = nb - nu;
    nvir_dn = nb - nd;
    ndet = 0;
    nword = nword_det(nb);
    nword2 = nword * 2;
    maxrank_up = binomial(nb, nu);
    maxrank_dn = binomial(nb, nd);
    nspin = nspin_;
    nw = nword * nspin;
    dets.clear();
    dict.reset(nw);
    full_space = false;
}

// binary format: 4 x int64 header (ndet, nbasis, nocc_up, nocc_dn) + raw uint64 determinants
// (onespinwfn.cpp:26-50,92-104; twospinwfn.cpp:26-50,95-107)
void Wfn::load_file(const std::string &filename, int nspin_) {
    std::ifstream file(filename, std::ios::in | std::ios::binary);
    long hdr[4];
    if (!file.read(reinterpret_cast<char *>(hdr), sizeof(hdr)))
        throw std::ios_base::failure("error in file");
    if (hdr[0] < 0 || hdr[1] < 0)
        throw std::ios_base::failure("error in file");
    std::vector<ulong> buf((size_t)(hdr[0] * nword_det(hdr[1]) * nspin_));
    if (!buf.empty() && !file.read(reinterpret_cast<char *>(buf.data()), sizeof(ulong) * buf.size()))
        throw std::ios_base::failure("error in file");
    init(hdr[1], hdr[2], hdr[3], nspin_);
    set_dets(hdr[0], buf.data());
}

void Wfn::to_file(const std::string &filename) const {
    std::ofstream file(filename, std::ios::out | std::ios::binary);
    const long hdr[4] = {ndet, nbasis, nocc_up, nocc_dn};
    bool ok = static_cast<bool>(file.write(reinterpret_cast<const char *>(hdr), sizeof(hdr)));
    if (ok && ndet)
        ok = static_cast<bool>(file.write(reinterpret_cast<const char *>(dets.data()), sizeof(ulong) * nw * ndet));
    file.close();
    if (!ok)
        throw std::ios_base::failure("error writing file");
}

// constructor from a determinant array: every row is kept, the dictionary maps a repeated string to
// its last position (onespinwfn.cpp:55-63, twospinwfn.cpp:55-63)
void Wfn::set_dets(long n, const ulong *ptr) {
    full_space = false;
    ndet = n;
    dets.assign(ptr, ptr + n * nw);
    dict.reset(nw);
    dict.assign_bulk(dets, 0, n);
}

// constructor from an occupation array: [n][nocc_up] or [n][2][nocc_up] (onespinwfn.cpp:65-78,
// twospinwfn.cpp:65-83: both spin rows have stride nocc_up)
void Wfn::set_occs(long n, const long *ptr) {
    std::vector<ulong> buf((size_t)(n * nw), 0UL);
    for (long i = 0; i < n; ++i) {
        fill_det(nocc_up, ptr + i * nspin * nocc_up, &buf[i * nw]);
        if (nspin == 2)
            fill_det(nocc_dn, ptr + (i * 2 + 1) * nocc_up, &buf[i * nw + nword]);
    }
    set_dets(n, buf.data());
}

long Wfn::index_det_from_rank(const Hash rank) const {
    if (rank_index_ndet_ != ndet) {
        rank_index_.clear();
        for (long i = 0; i < ndet; ++i)
            rank_index_[rank_det(det_ptr(i))] = i;
        rank_index_ndet_ = ndet;
    }
    auto it = rank_index_.find(rank);
    return it == rank_index_.end() ? -1 : it->second;
}

// onespinwfn.cpp:141-148
long Wfn::add_det(const ulong *det) {
    if (dict.find(dets, det) >= 0)
        return -1;
    // `det` may alias our own storage; copy before growing
    std::vector<ulong> tmp(det, det + nw);
    dets.insert(dets.end(), tmp.begin(), tmp.end());
    dict.assign(dets, &dets[ndet * nw], ndet);
    full_space = false;
    return ndet++;
}

// Bulk form of add_det for determinants the caller guarantees to be distinct and new (py_add_hci: the device
// de-duplicates and excludes the determinants already present).
void Wfn::append_new_dets(const ulong *ptr, long n) {
    if (n <= 0)
        return;
    dets.insert(dets.end(), ptr, ptr + (size_t)(n * nw));
    dict.insert_new_bulk(dets, ndet, n);
    full_space = false;
    ndet += n;
}

long Wfn::add_det_from_occs(const long *occs) {
    std::vector<ulong> det((size_t)nw, 0UL);
    fill_det(nocc_up, occs, &det[0]);
    if (nspin == 2)
        fill_det(nocc_dn, occs + nocc_up, &det[nword]);
    return add_det(det.data());
}

void Wfn::add_hartreefock_det() {
    std::vector<ulong> det((size_t)nw, 0UL);
    fill_hartreefock_det(nocc_up, &det[0]);
    if (nspin == 2)
        fill_hartreefock_det(nocc_dn, &det[nword]);
    add_det(det.data());
}

namespace {

// all C(nbasis, nocc) strings in colexicographic order (onespinwfn.cpp:173-185)
void colex_strings(long nbasis, long nocc, long nword, long count, std::vector<ulong> &out) {
    out.assign((size_t)(count * nword), 0UL);
    std::vector<long> occ((size_t)nocc + 2);
    for (long i = 0; i < nocc; ++i)
        occ[i] = i;
    occ[nocc] = nbasis + 1;
    occ[nocc + 1] = nbasis + 3;
    for (long idx = 0; idx < count; ++idx) {
        fill_det(nocc, occ.data(), &out[idx * nword]);
        if (nocc == 0)
            break;
        next_colex(occ.data());
    }
}

} // namespace

// replaces the contents with the full space: colex order (one-spin, onespinwfn.cpp:187-217) or
// idx = colex(alpha) * C(n, nocc_dn) + colex(beta) (two-spin, twospinwfn.cpp:181-245)
void Wfn::add_all_dets(long /*nthread*/) {
    if (maxrank_up == std::numeric_limits<long>::max() || maxrank_dn == std::numeric_limits<long>::max())
        throw std::domain_error("cannot generate > 2 ** 63 determinants");
    std::vector<ulong> up, dn;
    colex_strings(nbasis, nocc_up, nword, maxrank_up, up);
    if (nspin == 1) {
        set_new_dets(maxrank_up, up.data());
        full_space = nword == 1;
        return;
    }
    colex_strings(nbasis, nocc_dn, nword, maxrank_dn, dn);
    if (maxrank_up > std::numeric_limits<long>::max() / std::max(maxrank_dn, 1L))
        throw std::domain_error("cannot generate > 2 ** 63 determinants");
    const long n = maxrank_up * maxrank_dn;
    std::vector<ulong> all((size_t)(n * nw));
    for (long a = 0; a < maxrank_up; ++a)
        for (long b = 0; b < maxrank_dn; ++b) {
            ulong *d = &all[(a * maxrank_dn + b) * nw];
            std::memcpy(d, &up[a * nword], sizeof(ulong) * nword);
            std::memcpy(d + nword, &dn[b * nword], sizeof(ulong) * nword);
        }
    set_new_dets(n, all.data());
    full_space = nword == 1;
}

// replace the contents by n determinants that are distinct by construction (add_all_dets): no key comparisons
void Wfn::set_new_dets(long n, const ulong *ptr) {
    ndet = n;
    dets.assign(ptr, ptr + n * nw);
    dict.reset(nw);
    dict.insert_new_bulk(dets, 0, n);
}

// all e-fold excitations of one string, in the reference's order: outer loop over the colex
// combinations of virtuals, inner loop over the colex combinations of occupieds (onespinwfn.cpp:219-246)
void Wfn::onespin_excited(const ulong *rdet, long e, long nocc_s, std::vector<ulong> &out) const {
    const long nvir_s = nbasis - nocc_s;
    out.clear();
    if (e < 0 || e > nocc_s || e > nvir_s)
        return;
    const long no = binomial(nocc_s, e), nv = binomial(nvir_s, e);
    std::vector<long> occs((size_t)nocc_s + 1), virs((size_t)nvir_s + 1), oi((size_t)e + 2), vi((size_t)e + 2);
    std::vector<ulong> det((size_t)nword);
    fill_occs(nword, rdet, occs.data());
    fill_virs(nword, nbasis, rdet, virs.data());
    for (long k = 0; k < e; ++k)
        vi[k] = k;
    vi[e] = nvir_s + 1;
    vi[e + 1] = nvir_s + 3;
    for (long i = 0; i < nv; ++i) {
        for (long k = 0; k < e; ++k)
            oi[k] = k;
        oi[e] = nocc_s + 1;
        oi[e + 1] = nocc_s + 3;
        for (long j = 0; j < no; ++j) {
            std::memcpy(det.data(), rdet, sizeof(ulong) * nword);
            for (long k = 0; k < e; ++k)
                excite_det(occs[oi[k]], virs[vi[k]], det.data());
            out.insert(out.end(), det.begin(), det.end());
            if (e)
                next_colex(oi.data());
        }
        if (e)
            next_colex(vi.data());
    }
}

// onespinwfn.cpp:301-314 and twospinwfn.cpp:247-263,321-341
long Wfn::py_add_excited_dets(long exc, const py::object ref) {
    std::vector<ulong> rdet((size_t)nw, 0UL);
    if (ref.is_none()) {
        fill_hartreefock_det(nocc_up, &rdet[0]);
        if (nspin == 2)
            fill_hartreefock_det(nocc_dn, &rdet[nword]);
    } else {
        Array<ulong> a = ref.cast<Array<ulong>>();
        if ((long)a.size() < nw)
            throw std::invalid_argument("reference determinant has the wrong size");
        std::memcpy(rdet.data(), a.data(), sizeof(ulong) * nw);
    }
    const long before = ndet;
    std::vector<ulong> up, dn, det((size_t)nw);
    if (nspin == 1) {
        onespin_excited(rdet.data(), exc, nocc_up, up);
        for (size_t i = 0; i + nword <= up.size(); i += nword)
            add_det(&up[i]);
        return ndet - before;
    }
    const long maxup = std::min(nocc_up, nvir_up), maxdn = std::min(nocc_dn, nvir_dn);
    long a = std::min(exc, maxup), b = exc - a;
    while (a >= 0 && b <= maxdn) {
        if (a == 0 && b == 0) {
            add_det(rdet.data());
        } else {
            onespin_excited(&rdet[0], a, nocc_up, up);
            onespin_excited(&rdet[nword], b, nocc_dn, dn);
            for (size_t i = 0; i + nword <= up.size(); i += nword) {
                std::memcpy(&det[0], &up[i], sizeof(ulong) * nword);
                for (size_t j = 0; j + nword <= dn.size(); j += nword) {
                    std::memcpy(&det[nword], &dn[j], sizeof(ulong) * nword);
                    add_det(det.data());
                }
            }
        }
        --a;
        ++b;
    }
    return ndet - before;
}

// The reference iterates its hash map here (onespinwfn.cpp:246-249), which makes the resulting order
// an artefact of the map; we append in the other wave function's own order.
void Wfn::add_dets_from(const Wfn &other) {
    if (other.nw != nw || other.nbasis != nbasis)
        throw std::invalid_argument("wave functions are not compatible");
    for (long i = 0; i < other.ndet; ++i)
        add_det(other.det_ptr(i));
}

void Wfn::reserve(long n) {
    dets.reserve((size_t)(n * nw));
    dict.reserve(n);
}

// ---- python-facing helpers -------------------------------------------------------------------------

void Wfn::check_det_arg(const Array<ulong> &det) const {
    if ((long)det.size() < nw)
        throw std::invalid_argument("determinant array has the wrong size");
}

py::array Wfn::py_getitem(long index) const {
    if (index < 0)
        index += ndet;
    if (index < 0 || index >= ndet)
        throw py::index_error("determinant index out of range");
    if (nspin == 1) {
        Array<ulong> a(nword);
        std::memcpy(a.mutable_data(), det_ptr(index), sizeof(ulong) * nw);
        return a;
    }
    Array<ulong> a({2L, nword});
    std::memcpy(a.mutable_data(), det_ptr(index), sizeof(ulong) * nw);
    return a;
}

static void slice_args(long ndet, long &start, long &end) {
    // onespinwfn.cpp:259-267
    if (start == -1) {
        start = 0;
        if (end == -1)
            end = ndet;
    } else if (end == -1) {
        end = start;
        start = 0;
    }
    if (start < 0 || end > ndet || end < start)
        throw py::index_error("determinant range out of bounds");
}

py::array Wfn::py_to_det_array(long start, long end) const {
    slice_args(ndet, start, end);
    const long n = end - start;
    Array<ulong> a = (nspin == 1) ? Array<ulong>({n, nword}) : Array<ulong>({n, 2L, nword});
    if (n)
        std::memcpy(a.mutable_data(), &dets[start * nw], sizeof(ulong) * n * nw);
    return a;
}

py::array Wfn::py_to_occ_array(long start, long end) const {
    slice_args(ndet, start, end);
    const long n = end - start;
    Array<long> a = (nspin == 1) ? Array<long>({n, nocc_up}) : Array<long>({n, 2L, nocc_up});
    long *p = a.mutable_data();
    std::fill(p, p + a.size(), 0L);
    for (long i = 0; i < n; ++i) {
        fill_occs(nword, det_ptr(start + i), p + i * nspin * nocc_up);
        if (nspin == 2)
            fill_occs(nword, det_ptr(start + i) + nword, p + (i * 2 + 1) * nocc_up);
    }
    return a;
}

long Wfn::py_index_det(const Array<ulong> det) const {
    check_det_arg(det);
    return index_det(det.data());
}

Hash Wfn::py_rank_det(const Array<ulong> det) const {
    check_det_arg(det);
    return rank_det(det.data());
}

long Wfn::py_add_det(const Array<ulong> det) {
    check_det_arg(det);
    return add_det(det.data());
}

long Wfn::py_add_occs(const Array<long> occs) {
    if ((long)occs.size() < ((nspin == 2) ? nocc_up + nocc_dn : nocc_up))
        throw std::invalid_argument("occupation array has the wrong size");
    return add_det_from_occs(occs.data());
}

// ---- concrete classes ----------------------------------------------------------------------------------

static long rows_of(const py::array &a) { return a.ndim() ? a.shape(0) : 0; }

DOCIWfn::DOCIWfn(long nb, long nu, long nd) {
    init(nb, nu, nd, 1);
    if (nocc_up != nocc_dn)
        throw std::invalid_argument("nocc_up != nocc_dn");
}
DOCIWfn::DOCIWfn(const std::string &filename) {
    load_file(filename, 1);
    if (nocc_up != nocc_dn)
        throw std::invalid_argument("nocc_up != nocc_dn");
}
DOCIWfn::DOCIWfn(long nb, long nu, long nd, const Array<ulong> array) : DOCIWfn(nb, nu, nd) {
    set_dets(rows_of(array), array.data());
}
DOCIWfn::DOCIWfn(long nb, long nu, long nd, const Array<long> array) : DOCIWfn(nb, nu, nd) {
    set_occs(rows_of(array), array.data());
}

FullCIWfn::FullCIWfn(long nb, long nu, long nd) { init(nb, nu, nd, 2); }
FullCIWfn::FullCIWfn(const std::string &filename) { load_file(filename, 2); }
FullCIWfn::FullCIWfn(long nb, long nu, long nd, const Array<ulong> array) : FullCIWfn(nb, nu, nd) {
    set_dets(rows_of(array), array.data());
}
FullCIWfn::FullCIWfn(long nb, long nu, long nd, const Array<long> array) : FullCIWfn(nb, nu, nd) {
    set_occs(rows_of(array), array.data());
}
// fullciwfn.cpp:26-35: the same string for both spins
FullCIWfn::FullCIWfn(const DOCIWfn &wfn) {
    init(wfn.nbasis, wfn.nocc_up, wfn.nocc_dn, 2);
    std::vector<ulong> buf((size_t)(wfn.ndet * nw));
    for (long i = 0; i < wfn.ndet; ++i) {
        std::memcpy(&buf[i * nw], wfn.det_ptr(i), sizeof(ulong) * nword);
        std::memcpy(&buf[i * nw + nword], wfn.det_ptr(i), sizeof(ulong) * nword);
    }
    set_dets(wfn.ndet, buf.data());
}

GenCIWfn::GenCIWfn(long nb, long nu, long nd) {
    init(nb, nu, nd, 1);
    if (nocc_dn)
        throw std::invalid_argument("nocc_dn != 0");
}
GenCIWfn::GenCIWfn(const std::string &filename) {
    load_file(filename, 1);
    if (nocc_dn)
        throw std::invalid_argument("nocc_dn != 0");
}
GenCIWfn::GenCIWfn(long nb, long nu, long nd, const Array<ulong> array) : GenCIWfn(nb, nu, nd) {
    set_dets(rows_of(array), array.data());
}
GenCIWfn::GenCIWfn(long nb, long nu, long nd, const Array<long> array) : GenCIWfn(nb, nu, nd) {
    set_occs(rows_of(array), array.data());
}
// Spin-orbital form of a FullCI wave function: alpha orbitals first, beta orbital p at nbasis + p.
// (The reference's converter, genciwfn.cpp:29-45, reads the beta occupations from the alpha string and
// strides the output by nword2; this is the conversion it evidently intends -- see DESIGN.md.)
GenCIWfn::GenCIWfn(const FullCIWfn &wfn) {
    init(wfn.nbasis * 2, wfn.nocc, 0, 1);
    std::vector<ulong> buf((size_t)(wfn.ndet * nw), 0UL);
    std::vector<long> occs((size_t)wfn.nocc + 1);
    for (long i = 0; i < wfn.ndet; ++i) {
        fill_occs(wfn.nword, wfn.det_ptr(i), occs.data());
        fill_occs(wfn.nword, wfn.det_ptr(i) + wfn.nword, occs.data() + wfn.nocc_up);
        for (long j = 0; j < wfn.nocc_dn; ++j)
            occs[wfn.nocc_up + j] += wfn.nbasis;
        fill_det(wfn.nocc, occs.data(), &buf[i * nw]);
    }
    set_dets(wfn.ndet, buf.data());
}
GenCIWfn::GenCIWfn(const DOCIWfn &wfn) : GenCIWfn(FullCIWfn(wfn)) {}

} // namespace pyci_host
