// 1- and 2-particle reduced density matrices: the device form of compute_rdms
// (/root/reference/pyci/src/rdm.cpp:20-65 DOCI, :269-530 FullCI, :532-632 GenCI).
//
// Same excitation enumerator and determinant index as the Hamiltonian construction (enumerate.cuh).
// One CTA per determinant row; every connected pair idet < jdet is visited once from the idet side and
// c_i c_j sign is scattered to all symmetry-related tensor positions with fp64 atomics (RED in L2).
// GenCI follows the fully antisymmetric convention of the FullCI same-spin blocks: the reference's GenCI
// routine is defective in this snapshot (buffer sizes, index typos, nvir bound; DESIGN.md, "GenCI").
#include <algorithm>
#include <cstring>

#include "enumerate.cuh"

namespace {

// eight symmetry-related positions of a same-spin contribution G[p,q,r,s] (rdm.cpp:399-414)
__device__ __forceinline__ void scatter8(double *G, long n, long p, long q, long r, long s, double x) {
    const long n1 = n, n2 = n * n, n3 = n2 * n;
    atomicAdd(G + p * n3 + q * n2 + r * n1 + s, x);
    atomicAdd(G + p * n3 + q * n2 + s * n1 + r, -x);
    atomicAdd(G + q * n3 + p * n2 + r * n1 + s, -x);
    atomicAdd(G + q * n3 + p * n2 + s * n1 + r, x);
    atomicAdd(G + r * n3 + s * n2 + p * n1 + q, x);
    atomicAdd(G + r * n3 + s * n2 + q * n1 + p, -x);
    atomicAdd(G + s * n3 + r * n2 + p * n1 + q, -x);
    atomicAdd(G + s * n3 + r * n2 + q * n1 + p, x);
}

// diagonal same-spin pair (rdm.cpp:308-318)
__device__ __forceinline__ void scatter4_diag(double *G, long n, long p, long q, double x) {
    const long n1 = n, n2 = n * n, n3 = n2 * n;
    atomicAdd(G + p * n3 + q * n2 + p * n1 + q, x);
    atomicAdd(G + p * n3 + q * n2 + q * n1 + p, -x);
    atomicAdd(G + q * n3 + p * n2 + p * n1 + q, -x);
    atomicAdd(G + q * n3 + p * n2 + q * n1 + p, x);
}

constexpr int RDM_UNROLL = 4;

template<int KIND, int KM>
__global__ void __launch_bounds__(256) rdm_kernel(BuildParams P, DetIndex<KM> index) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uchar2 *pairs = reinterpret_cast<uchar2 *>(smem_raw);
    __shared__ RowShared rs;
    fill_pairs(pairs, P.npairs_dim);
    const int nspin = (KIND == PYCI_FULLCI) ? 2 : 1;
    const long n = P.n, n1 = n, n2 = n * n, n3 = n2 * n, n4 = n2 * n2;
    double *aa = P.rdm1, *bb = P.rdm1 + n2;
    double *aaaa = P.rdm2, *bbbb = P.rdm2 + n4, *abab = P.rdm2 + 2 * n4;
    for (long r = blockIdx.x; r < P.nloc; r += gridDim.x) {
        const long row = P.row0 + r;
        __syncthreads();
        row_setup(rs, P, row, nspin);
        __syncthreads();
        const double ci = __ldg(P.coeffs + row);
        const double val1 = ci * ci;
        const int na = rs.nocc[0], nb = (KIND == PYCI_FULLCI) ? rs.nocc[1] : 0;
        // ---- diagonal terms: all ordered pairs of occupied orbitals, split over the threads
        if (KIND == PYCI_DOCI) {
            // rdm.cpp:41-50: d0[k,k] += c^2 ; d2[k,l] += c^2 for k != l both occupied
            for (int t = threadIdx.x; t < na * na; t += blockDim.x) {
                const int i = t / na, j = t - i * na;
                const long k = rs.occ[0][i], l = rs.occ[0][j];
                if (i == j)
                    atomicAdd(P.rdm1 + k * (n + 1), val1);
                else
                    atomicAdd(P.rdm2 + k * n + l, val1);
            }
        } else {
            const int ntot = na + nb;
            for (int t = threadIdx.x; t < ntot * ntot; t += blockDim.x) {
                const int i = t / ntot, j = t - i * ntot;
                if (j < i)
                    continue;
                const bool ib = i >= na, jb = j >= na;
                const long p = ib ? rs.occ[1][i - na] : rs.occ[0][i];
                const long q = jb ? rs.occ[1][j - na] : rs.occ[0][j];
                if (i == j) {
                    atomicAdd((ib ? bb : aa) + (n1 + 1) * p, val1); // rdm.cpp:306,431
                } else if (ib == jb) {
                    scatter4_diag(ib ? bbbb : aaaa, n, p, q, val1); // rdm.cpp:308-318,433-441
                } else {
                    atomicAdd(abab + p * n3 + q * n2 + p * n1 + q, val1); // rdm.cpp:320-323 (i alpha, j beta)
                }
            }
        }
        // ---- off-diagonal terms: connected pairs with jdet > idet
        for (u32 base = 0; base < P.ncand; base += RDM_UNROLL * blockDim.x) {
            int hit[RDM_UNROLL];
            u32 codes[RDM_UNROLL];
#pragma unroll
            for (int u = 0; u < RDM_UNROLL; ++u) {
                const u32 c = base + u * blockDim.x + threadIdx.x;
                hit[u] = -1;
                codes[u] = 0;
                if (c < P.ncand) {
                    u64 A, B;
                    decode<KIND>(P, rs, pairs, c, A, B, codes[u]);
                    hit[u] = index.find(A, B);
                }
            }
#pragma unroll
            for (int u = 0; u < RDM_UNROLL; ++u) {
                if ((long)hit[u] <= row)
                    continue;
                const u32 code = codes[u];
                const int type = code >> 24;
                const long i = (code >> 18) & 63, a = (code >> 12) & 63, k = (code >> 6) & 63, l = code & 63;
                const double cc = ci * __ldg(P.coeffs + hit[u]);
                switch (type) {
                case T_PAIR: // rdm.cpp:57-61
                    atomicAdd(P.rdm1 + n * i + a, cc);
                    atomicAdd(P.rdm1 + n * a + i, cc);
                    break;
                case T_AB: { // rdm.cpp:372-379
                    const double x = apply_sign(cc, parity_single(rs.det[0], (int)i, (int)a) ^
                                                        parity_single(rs.det[1], (int)k, (int)l));
                    atomicAdd(abab + i * n3 + k * n2 + a * n1 + l, x);
                    atomicAdd(abab + a * n3 + l * n2 + i * n1 + k, x);
                    break;
                }
                case T_AA: // rdm.cpp:393-414
                    scatter8(aaaa, n, i, k, a, l, apply_sign(cc, parity_double(rs.det[0], (int)i, (int)k, (int)a, (int)l)));
                    break;
                case T_BB: // rdm.cpp:495-517
                    scatter8(bbbb, n, i, k, a, l, apply_sign(cc, parity_double(rs.det[1], (int)i, (int)k, (int)a, (int)l)));
                    break;
                case T_SA: { // rdm.cpp:329-362
                    const double x = apply_sign(cc, parity_single(rs.det[0], (int)i, (int)a));
                    atomicAdd(aa + i * n1 + a, x);
                    atomicAdd(aa + a * n1 + i, x);
                    for (int q = 0; q < na; ++q) {
                        const long kk = rs.occ[0][q];
                        if (kk != i)
                            scatter8(aaaa, n, i, kk, a, kk, x);
                    }
                    for (int q = 0; q < nb; ++q) {
                        const long kk = rs.occ[1][q];
                        atomicAdd(abab + i * n3 + kk * n2 + a * n1 + kk, x);
                        atomicAdd(abab + a * n3 + kk * n2 + i * n1 + kk, x);
                    }
                    break;
                }
                case T_SB: { // rdm.cpp:446-483
                    const double x = apply_sign(cc, parity_single(rs.det[1], (int)i, (int)a));
                    atomicAdd(bb + i * n1 + a, x);
                    atomicAdd(bb + a * n1 + i, x);
                    for (int q = 0; q < na; ++q) {
                        const long kk = rs.occ[0][q];
                        atomicAdd(abab + kk * n3 + i * n2 + kk * n1 + a, x);
                        atomicAdd(abab + kk * n3 + a * n2 + kk * n1 + i, x);
                    }
                    for (int q = 0; q < nb; ++q) {
                        const long kk = rs.occ[1][q];
                        if (kk != i)
                            scatter8(bbbb, n, i, kk, a, kk, x);
                    }
                    break;
                }
                default:
                    break;
                }
            }
        }
    }
}

template<int KIND, int KM>
int run_rdm(pyci_ctx *ctx, const pyci_wfn *wfn, BuildParams &P) {
    const DetIndex<KM> ix = make_index<KM>(wfn);
    const size_t smem = pair_table_bytes(P);
    const long work = (long)P.ncand / 4;
    const int block = work <= 256 ? 32 : work <= 1024 ? 64 : work <= 4096 ? 128 : 256;
    int per_sm = 1;
    PYCI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rdm_kernel<KIND, KM>, block, smem));
    const long grid = std::min<long>(P.nloc, (long)ctx->sm_count * std::max(per_sm, 1));
    if (grid > 0) {
        rdm_kernel<KIND, KM><<<(unsigned)grid, block, smem, ctx->stream>>>(P, ix);
        ctx->launches++;
    }
    PYCI_CUDA(cudaGetLastError());
    return PYCI_OK;
}

template<int KIND>
int rdm_dispatch(pyci_ctx *ctx, const pyci_wfn *wfn, BuildParams &P) {
    switch (wfn->keymode) {
    case KEY32:
        return run_rdm<KIND, KEY32>(ctx, wfn, P);
    case KEY64:
        return run_rdm<KIND, KEY64>(ctx, wfn, P);
    default:
        if constexpr (KIND == PYCI_FULLCI)
            return run_rdm<KIND, KEY128>(ctx, wfn, P);
        else
            PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "one-spin wave functions use 32- or 64-bit keys");
    }
}

} // namespace

int rdms_impl(pyci_ctx *ctx, const pyci_wfn *wfn, const double *coeffs, double *rdm1, double *rdm2) {
    BuildParams P;
    PYCI_TRY(enum_params_init(P, wfn));
    const long n = wfn->nbasis, n2 = n * n, n4 = n2 * n2;
    const int kind = wfn->kind;
    const size_t s1 = (size_t)((kind == PYCI_FULLCI) ? 2 * n2 : n2);
    const size_t s2 = (size_t)((kind == PYCI_FULLCI) ? 3 * n4 : (kind == PYCI_DOCI) ? n2 : n4);
    // rows are split evenly over the ranks; every rank needs all coefficients
    const long R = ctx->nranks, ndet = wfn->ndet;
    const long per = (ndet + R - 1) / R;
    P.row0 = std::min(ndet, per * ctx->rank);
    P.nloc = std::min(ndet, per * (ctx->rank + 1)) - P.row0;
    P.ncol = ndet;
    double *dc = nullptr, *d12 = nullptr;
    PYCI_CUDA(dev_malloc(&dc, sizeof(double) * (size_t)std::max<long>(ndet, 1)));
    cudaError_t e = dev_malloc(&d12, sizeof(double) * (s1 + s2));
    if (e != cudaSuccess) {
        dev_free(dc);
        PYCI_CUDA(e);
    }
    int rc = PYCI_OK;
    auto body = [&]() -> int {
        PYCI_CUDA(cudaMemcpyAsync(dc, coeffs, sizeof(double) * ndet, cudaMemcpyHostToDevice, ctx->stream));
        PYCI_CUDA(cudaMemsetAsync(d12, 0, sizeof(double) * (s1 + s2), ctx->stream));
        P.coeffs = dc;
        P.rdm1 = d12;
        P.rdm2 = d12 + s1;
        if (kind == PYCI_DOCI)
            PYCI_TRY(rdm_dispatch<PYCI_DOCI>(ctx, wfn, P));
        else if (kind == PYCI_FULLCI)
            PYCI_TRY(rdm_dispatch<PYCI_FULLCI>(ctx, wfn, P));
        else
            PYCI_TRY(rdm_dispatch<PYCI_GENCI>(ctx, wfn, P));
        PYCI_TRY(comm_allreduce_sum_f64(ctx, d12, (long)(s1 + s2)));
        PYCI_CUDA(cudaMemcpyAsync(rdm1, d12, sizeof(double) * s1, cudaMemcpyDeviceToHost, ctx->stream));
        PYCI_CUDA(cudaMemcpyAsync(rdm2, d12 + s1, sizeof(double) * s2, cudaMemcpyDeviceToHost, ctx->stream));
        PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
        return PYCI_OK;
    };
    rc = body();
    dev_free(dc);
    dev_free(d12);
    return rc;
}
