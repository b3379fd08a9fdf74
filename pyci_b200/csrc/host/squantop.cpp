// Second-quantised Hamiltonian container of pyci_b200._pyci: FCIDUMP reader/writer and the
// seniority-zero integrals, with the behaviour of the reference's SQuantOp
// (/root/reference/pyci/src/squantop.cpp:79-223).  Stays on the host; the integrals are uploaded to
// HBM when an operator is built (pyci_ham_upload).
#include <cmath>
#include <fstream>
#include <iomanip>
#include <regex>
#include <sstream>

#include "pyci_host.h"

namespace pyci_host {

// h[p] = t_pp, v[p,q] = <pp|qq>, w[p,q] = 2<pq|pq> - <pq|qp> (squantop.cpp:152-160,171-182)
void SQuantOp::derive_senzero() {
    const long n1 = nbasis, n2 = n1 * n1, n3 = n2 * n1;
    h_array = Array<double>(nbasis);
    v_array = Array<double>({nbasis, nbasis});
    w_array = Array<double>({nbasis, nbasis});
    const double *t = one_mo_array.data(), *g = two_mo_array.data();
    double *h = h_array.mutable_data(), *v = v_array.mutable_data(), *w = w_array.mutable_data();
    for (long i = 0; i < n1; ++i) {
        h[i] = t[i * (n1 + 1)];
        for (long j = 0; j < n1; ++j) {
            v[i * n1 + j] = g[i * n3 + i * n2 + j * n1 + j];
            w[i * n1 + j] = g[i * n3 + j * n2 + i * n1 + j] * 2 - g[i * n3 + j * n2 + j * n1 + i];
        }
    }
}

SQuantOp::SQuantOp(double e, const Array<double> mo1, const Array<double> mo2)
    : nbasis(mo1.ndim() ? mo1.shape(0) : 0), ecore(e), one_mo_array(mo1), two_mo_array(mo2) {
    const long n2 = nbasis * nbasis;
    if (nbasis < 1 || one_mo_array.size() != n2 || two_mo_array.size() != n2 * n2)
        throw std::invalid_argument("one_mo must be (n, n) and two_mo (n, n, n, n)");
    derive_senzero();
}

SQuantOp::SQuantOp(const std::string &filename) {
    std::ifstream f(filename);
    if (f.fail())
        throw std::ios_base::failure("Failed to read the FCIDUMP file " + filename);
    std::string header, line;
    bool ended = false;
    while (std::getline(f, line)) {
        if (line.find("&END") != std::string::npos || line.find("/") != std::string::npos) {
            ended = true;
            break;
        }
        header += " ";
        header += line;
    }
    if (!ended)
        throw std::ios_base::failure("FCIDUMP has the wrong header");
    std::smatch m;
    if (!std::regex_search(header, m, std::regex(R"(NORB[ ]*=[ ]*(\d+))")))
        throw std::invalid_argument("NORB is not found.");
    nbasis = std::stol(m[1]);
    if (std::regex_search(header, m, std::regex(R"(UHF[ ]*=[ .]*(FALSE|TRUE))", std::regex::icase))) {
        std::string s = m[1];
        if (s[0] == 'T' || s[0] == 't')
            throw std::runtime_error("Unrestricted FCIDUMP not implemented");
    }
    const long n1 = nbasis, n2 = n1 * n1, n3 = n2 * n1;
    one_mo_array = Array<double>({nbasis, nbasis});
    two_mo_array = Array<double>({nbasis, nbasis, nbasis, nbasis});
    double *one = one_mo_array.mutable_data(), *two = two_mo_array.mutable_data();
    std::fill(one, one + n2, 0.0);
    std::fill(two, two + n2 * n2, 0.0);
    ecore = 0.0;
    double x;
    long i, j, k, l;
    while (f >> x >> i >> j >> k >> l) {
        if (i > n1 || j > n1 || k > n1 || l > n1 || i < 0 || j < 0 || k < 0 || l < 0)
            throw std::ios_base::failure("FCIDUMP orbital index out of range");
        if (i && j && k && l) {
            --i, --j, --k, --l;
            // chemists' (ij|kl) with its 8-fold symmetry, stored in physicist order <ik|jl>
            two[i * n3 + k * n2 + j * n1 + l] = x;
            two[k * n3 + i * n2 + l * n1 + j] = x;
            two[j * n3 + k * n2 + i * n1 + l] = x;
            two[i * n3 + l * n2 + j * n1 + k] = x;
            two[j * n3 + l * n2 + i * n1 + k] = x;
            two[l * n3 + j * n2 + k * n1 + i] = x;
            two[k * n3 + j * n2 + l * n1 + i] = x;
            two[l * n3 + i * n2 + k * n1 + j] = x;
        } else if (i && j) {
            --i, --j;
            one[i * n1 + j] = x;
            one[j * n1 + i] = x;
        } else {
            ecore = x;
        }
    }
    derive_senzero();
}

void SQuantOp::to_file(const std::string &filename, long nelec, long ms2, double tol) const {
    const long n1 = nbasis, n2 = n1 * n1, n3 = n2 * n1;
    std::ofstream f(filename);
    if (f.fail())
        throw std::ios_base::failure("Failed to open the FCIDUMP file " + filename);
    f << "&FCIDUMP\nNORB=" << nbasis << ",\nNELEC=" << nelec << ",\nMS2=" << ms2 << ",\nUHF=.FALSE.,\nORBSYM=";
    for (long i = 0; i < nbasis; ++i)
        f << "1,";
    f << "\nISYM=1,\n&END\n";
    const double *one = one_mo(), *two = two_mo();
    auto put = [&](double val) -> std::ostream & {
        return f << std::setw(28) << std::setprecision(20) << std::scientific << val;
    };
    for (long i = 0; i < n1; ++i)
        for (long j = 0; j <= i; ++j)
            for (long k = 0; k < n1; ++k)
                for (long l = 0; l <= k; ++l)
                    if ((i * (i + 1)) / 2 + j >= (k * (k + 1)) / 2 + l) {
                        const double val = two[i * n3 + k * n2 + j * n1 + l];
                        if (std::abs(val) > tol)
                            put(val) << ' ' << i + 1 << ' ' << j + 1 << ' ' << k + 1 << ' ' << l + 1 << "\n";
                    }
    for (long i = 0; i < n1; ++i)
        for (long j = 0; j <= i; ++j) {
            const double val = one[i * n1 + j];
            if (std::abs(val) > tol)
                put(val) << ' ' << i + 1 << ' ' << j + 1 << " 0 0\n";
        }
    put(ecore) << " 0 0 0 0\n";
}

} // namespace pyci_host
