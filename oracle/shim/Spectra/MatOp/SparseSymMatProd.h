// TEST INFRASTRUCTURE ONLY (oracle/): see SymEigsSolver.h in this directory tree.
#pragma once
namespace Spectra {
template<class T, int UpLo, int Order, class I>
class SparseSymMatProd {
public:
    template<class M>
    explicit SparseSymMatProd(const M &) {}
};
} // namespace Spectra
