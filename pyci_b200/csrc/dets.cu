// Determinant generation on the device: the B200 form of Wfn::add_all_dets
// (/root/reference/pyci/src/onespinwfn.cpp:173-217, twospinwfn.cpp:181-245).  A complete space is defined by three
// integers (nbasis, nocc_up, nocc_dn); unranking it in HBM means a construction from such a wave function moves no
// determinant over PCIe (one upload per process otherwise: 53 MB per rank at FullCI(16, 4a4b)).
#include <vector>

#include "common.cuh"

namespace {

// string of colex rank r among the C(n, k) strings: the combinatorial number system r = sum_j C(c_j, j),
// c_k > ... > c_1 >= 0 (colex order = ascending integer value, unrank_colex of common.cpp:115-135)
__global__ void colex_strings_kernel(int n, int k, long count, const u64 *__restrict__ binom, u64 *__restrict__ out) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= count)
        return;
    u64 s = 0ULL, rem = (u64)r;
    int p = n - 1;
    for (int j = k; j >= 1; --j) {
        while (binom[p * 65 + j] > rem)
            --p;
        s |= 1ULL << p;
        rem -= binom[p * 65 + j];
        --p;
    }
    out[r] = s;
}

// two-spin order: idx = colex(alpha) * Nb + colex(beta) (twospinwfn.cpp:195-218)
__global__ void product_dets_kernel(const u64 *__restrict__ A, const u64 *__restrict__ B, long nb, long ndet,
                                    u64 *__restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ndet)
        return;
    const long a = i / nb, b = i - a * nb;
    reinterpret_cast<ulonglong2 *>(out)[i] = make_ulonglong2(A[a], B[b]);
}

} // namespace

// dets of the complete space of wfn (kind / nbasis / occupations set, ndet = size of the space) into wfn->dets
int wfn_generate_all_dets(pyci_wfn *wfn, long na, long nb) {
    pyci_ctx *ctx = wfn->ctx;
    cudaStream_t st = ctx->stream;
    std::vector<u64> hb(65 * 65, 0ULL);
    for (int p = 0; p < 65; ++p) {
        hb[p * 65] = 1ULL;
        for (int j = 1; j <= p; ++j) {
            const unsigned __int128 v = (p - 1 >= j ? (unsigned __int128)hb[(p - 1) * 65 + j] : 0) +
                                        (unsigned __int128)hb[(p - 1) * 65 + j - 1];
            hb[p * 65 + j] = v > (unsigned __int128)0xffffffffffffffffULL ? 0xffffffffffffffffULL : (u64)v;
        }
    }
    u64 *dbinom = nullptr, *sa = nullptr, *sb = nullptr;
    auto body = [&]() -> int {
        PYCI_CUDA(dev_malloc(&dbinom, sizeof(u64) * hb.size()));
        PYCI_CUDA(cudaMemcpyAsync(dbinom, hb.data(), sizeof(u64) * hb.size(), cudaMemcpyHostToDevice, st));
        PYCI_CUDA(dev_malloc(&wfn->dets, sizeof(u64) * (size_t)std::max<long>(wfn->ndet * wfn->nwords, 1)));
        if (wfn->ndet == 0)
            return PYCI_OK;
        if (wfn->kind != PYCI_FULLCI) {
            colex_strings_kernel<<<(unsigned)((na + 255) / 256), 256, 0, st>>>((int)wfn->nbasis, (int)wfn->nocc_up, na, dbinom,
                                                                             wfn->dets);
            ctx->launches++;
        } else {
            PYCI_CUDA(dev_malloc(&sa, sizeof(u64) * (size_t)na));
            PYCI_CUDA(dev_malloc(&sb, sizeof(u64) * (size_t)nb));
            colex_strings_kernel<<<(unsigned)((na + 255) / 256), 256, 0, st>>>((int)wfn->nbasis, (int)wfn->nocc_up, na, dbinom, sa);
            colex_strings_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>((int)wfn->nbasis, (int)wfn->nocc_dn, nb, dbinom, sb);
            product_dets_kernel<<<(unsigned)((wfn->ndet + 255) / 256), 256, 0, st>>>(sa, sb, nb, wfn->ndet, wfn->dets);
            ctx->launches += 3;
        }
        PYCI_CUDA(cudaGetLastError());
        PYCI_CUDA(cudaStreamSynchronize(st)); // hb is read by the copy above
        return PYCI_OK;
    };
    const int rc = body();
    dev_free(dbinom);
    dev_free(sa);
    dev_free(sb);
    return rc;
}
