// TEST INFRASTRUCTURE ONLY (oracle/): Spectra is an un-vendored, unpinned dependency of the
// reference (Makefile:110-120) and is absent here.  This stub lets sparseop.cpp:114-146 compile;
// calling solve() on the compiled reference raises.  Eigenvalue oracle = scipy eigsh on the
// exported CSR (see oracle/README.md).
#pragma once
#include <stdexcept>
#include <Eigen/Core>

namespace Spectra {

enum class SortRule { SmallestAlge };
enum class CompInfo { Successful, NotComputed };

template<class Op>
class SymEigsSolver {
public:
    template<class... A>
    SymEigsSolver(A &&...) {}
    void init() {}
    void init(const double *) {}
    template<class... A>
    long compute(A &&...) {
        throw std::runtime_error("oracle/_ref: Spectra is not available; use scipy eigsh on the CSR");
    }
    CompInfo info() const { return CompInfo::NotComputed; }
    Eigen::Placeholder eigenvalues() const { return {}; }
    Eigen::Placeholder eigenvectors() const { return {}; }
};

} // namespace Spectra
