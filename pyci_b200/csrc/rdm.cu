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
#include "join.cuh"

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
__global__ void __launch_bounds__(256) rdm_kernel(BuildParams P, DetIndex<KM> index, u32 pair_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uchar2 *pairs = reinterpret_cast<uchar2 *>(smem_raw);
    __shared__ RowShared rs;
    fill_pairs(pairs, P.npairs_dim);
    const int nspin = (KIND == PYCI_FULLCI) ? 2 : 1;
    // selected space (Bloom filter present): most candidates miss, so only the strings are formed up front (pair
    // masks) and the excitation code is decoded for the hits
    const bool lazy = KIND != PYCI_DOCI && index.bloom != nullptr;
    if (lazy)
        pair_masks_carve(rs, P, smem_raw + pair_bytes, nspin);
    else
        pair_masks_none(rs);
    const long n = P.n, n1 = n, n2 = n * n, n3 = n2 * n, n4 = n2 * n2;
    double *aa = P.rdm1, *bb = P.rdm1 + n2;
    double *aaaa = P.rdm2, *bbbb = P.rdm2 + n4, *abab = P.rdm2 + 2 * n4;
    for (long r = blockIdx.x; r < P.nloc; r += gridDim.x) {
        const long row = P.row0 + r;
        __syncthreads();
        row_setup(rs, P, row, nspin);
        __syncthreads();
        if (lazy) {
            pair_masks_build(rs, P, pairs, nspin);
            __syncthreads();
        }
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
            u64 A[RDM_UNROLL], B[RDM_UNROLL];
            bool want[RDM_UNROLL];
#pragma unroll
            for (int u = 0; u < RDM_UNROLL; ++u) {
                const u32 c = base + u * blockDim.x + threadIdx.x;
                want[u] = c < P.ncand;
                codes[u] = 0;
                A[u] = B[u] = 0ULL;
                if (want[u]) {
                    if (lazy)
                        decode_dets<KIND>(P, rs, pairs, c, A[u], B[u]);
                    else
                        decode<KIND>(P, rs, pairs, c, A[u], B[u], codes[u]);
                }
            }
            find_batch<RDM_UNROLL>(index, A, B, want, hit);
#pragma unroll
            for (int u = 0; u < RDM_UNROLL; ++u) {
                if ((long)hit[u] <= row)
                    continue;
                u32 code = codes[u];
                if (lazy) {
                    u64 a2, b2;
                    decode<KIND>(P, rs, pairs, base + u * blockDim.x + threadIdx.x, a2, b2, code);
                }
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
    DetIndex<KM> ix = make_index<KM>(wfn);
    const size_t pb = (pair_table_bytes(P) + 7) & ~(size_t)7;
    const size_t smem = pb + pair_mask_bytes(P, KIND);
    BuildParams Pk = P;
    // Selected two-body space: the connected pairs come from the segment-pair join (join.cuh) instead of one index
    // probe per candidate excitation; the enumeration kernel then only adds the diagonal terms.
    if constexpr (KIND != PYCI_DOCI) {
        const bool force = getenv("PYCI_B200_FORCE_JOIN") != nullptr;
        if (!wfn->complete && (P.ncand >= 8192 || force) && !getenv("PYCI_B200_NO_JOIN")) {
            int used = 0;
            const double budget = force ? 1.0e300 : 10.0 * (double)P.ncand * (double)P.nloc;
            PYCI_TRY((join_run<KIND, JOIN_RDM>(ctx, wfn, P, nullptr, 0, nullptr, budget, &used, nullptr)));
            if (used) {
                Pk.ncand = 0;       // no off-diagonal enumeration
                ix.bloom = nullptr; // ... and no per-row pair masks for it
            }
        }
    }
    const long work = (long)Pk.ncand / 4;
    const int block = work <= 256 ? 32 : work <= 1024 ? 64 : work <= 4096 ? 128 : 256;
    int per_sm = 1;
    PYCI_CUDA(cudaFuncSetAttribute(rdm_kernel<KIND, KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PYCI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rdm_kernel<KIND, KM>, block, smem));
    const long grid = std::min<long>(P.nloc, (long)ctx->sm_count * std::max(per_sm, 1));
    if (grid > 0) {
        rdm_kernel<KIND, KM><<<(unsigned)grid, block, smem, ctx->stream>>>(Pk, ix, (u32)pb);
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


// ---- transition density matrices: compute_transition_rdms (rdm.cpp:634-1009) ---------------------------
// Rows are the determinants of wfn1 (P.dets), the excited determinant is looked up in wfn2's index, and every
// connected ordered pair contributes c1_i c2_j sign in ONE direction (the four positions of rdm.cpp:817-824 instead
// of the eight of the symmetric routine).  T(wfn, wfn, c, c) = compute_rdms(wfn, c).

__device__ __forceinline__ void scatter4_dir(double *G, long n, long p, long q, long r, long s, double x) {
    const long n1 = n, n2 = n * n, n3 = n2 * n;
    atomicAdd(G + p * n3 + q * n2 + r * n1 + s, x);
    atomicAdd(G + p * n3 + q * n2 + s * n1 + r, -x);
    atomicAdd(G + q * n3 + p * n2 + r * n1 + s, -x);
    atomicAdd(G + q * n3 + p * n2 + s * n1 + r, x);
}

// P.coeffs = coefficients of wfn1, c2 = coefficients of wfn2
template<int KIND, int KM>
__global__ void __launch_bounds__(256) trdm_kernel(BuildParams P, DetIndex<KM> index2, const double *__restrict__ c2,
                                                   u32 pair_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uchar2 *pairs = reinterpret_cast<uchar2 *>(smem_raw);
    __shared__ RowShared rs;
    __shared__ int self_hit;
    fill_pairs(pairs, P.npairs_dim);
    const int nspin = (KIND == PYCI_FULLCI) ? 2 : 1;
    const bool lazy = KIND != PYCI_DOCI && index2.bloom != nullptr;
    if (lazy)
        pair_masks_carve(rs, P, smem_raw + pair_bytes, nspin);
    else
        pair_masks_none(rs);
    const long n = P.n, n1 = n, n2 = n * n, n3 = n2 * n, n4 = n2 * n2;
    double *aa = P.rdm1, *bb = P.rdm1 + n2;
    double *aaaa = P.rdm2, *bbbb = P.rdm2 + n4, *abab = P.rdm2 + 2 * n4;
    for (long r = blockIdx.x; r < P.nloc; r += gridDim.x) {
        const long row = P.row0 + r;
        __syncthreads();
        row_setup(rs, P, row, nspin);
        __syncthreads();
        if (threadIdx.x == 0)
            self_hit = index2.find(rs.det[0], rs.det[1]);
        if (lazy)
            pair_masks_build(rs, P, pairs, nspin);
        __syncthreads();
        const double ci = __ldg(P.coeffs + row);
        const double val1 = (self_hit >= 0) ? ci * __ldg(c2 + self_hit) : 0.0; // rdm.cpp:655-656,719-720
        const int na = rs.nocc[0], nb = (KIND == PYCI_FULLCI) ? rs.nocc[1] : 0;
        if (val1 != 0.0) { // 0-0 terms
            if (KIND == PYCI_DOCI) {
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
                    if (i == j)
                        atomicAdd((ib ? bb : aa) + (n1 + 1) * p, val1);
                    else if (ib == jb)
                        scatter4_diag(ib ? bbbb : aaaa, n, p, q, val1);
                    else
                        atomicAdd(abab + p * n3 + q * n2 + p * n1 + q, val1);
                }
            }
        }
        for (u32 base = 0; base < P.ncand; base += RDM_UNROLL * blockDim.x) {
            int hit[RDM_UNROLL];
            u32 codes[RDM_UNROLL];
            u64 A[RDM_UNROLL], B[RDM_UNROLL];
            bool want[RDM_UNROLL];
#pragma unroll
            for (int u = 0; u < RDM_UNROLL; ++u) {
                const u32 c = base + u * blockDim.x + threadIdx.x;
                want[u] = c < P.ncand;
                codes[u] = 0;
                A[u] = B[u] = 0ULL;
                if (want[u]) {
                    if (lazy)
                        decode_dets<KIND>(P, rs, pairs, c, A[u], B[u]);
                    else
                        decode<KIND>(P, rs, pairs, c, A[u], B[u], codes[u]);
                }
            }
            find_batch<RDM_UNROLL>(index2, A, B, want, hit);
#pragma unroll
            for (int u = 0; u < RDM_UNROLL; ++u) {
                if (hit[u] < 0)
                    continue;
                u32 code = codes[u];
                if (lazy) {
                    u64 a2, b2;
                    decode<KIND>(P, rs, pairs, base + u * blockDim.x + threadIdx.x, a2, b2, code);
                }
                const int type = code >> 24;
                const long i = (code >> 18) & 63, a = (code >> 12) & 63, k = (code >> 6) & 63, l = code & 63;
                const double cc = ci * __ldg(c2 + hit[u]);
                switch (type) {
                case T_PAIR: // rdm.cpp:668-676
                    atomicAdd(P.rdm1 + n * i + a, cc);
                    break;
                case T_AB: // :777-796
                    atomicAdd(abab + i * n3 + k * n2 + a * n1 + l,
                              apply_sign(cc, parity_single(rs.det[0], (int)i, (int)a) ^ parity_single(rs.det[1], (int)k, (int)l)));
                    break;
                case T_AA: // :798-825
                    scatter4_dir(aaaa, n, i, k, a, l, apply_sign(cc, parity_double(rs.det[0], (int)i, (int)k, (int)a, (int)l)));
                    break;
                case T_BB: // :875-898
                    scatter4_dir(bbbb, n, i, k, a, l, apply_sign(cc, parity_double(rs.det[1], (int)i, (int)k, (int)a, (int)l)));
                    break;
                case T_SA: { // :751-775
                    const double x = apply_sign(cc, parity_single(rs.det[0], (int)i, (int)a));
                    atomicAdd(aa + i * n1 + a, x);
                    for (int q = 0; q < na; ++q) {
                        const long kk = rs.occ[0][q];
                        if (kk != i) {
                            atomicAdd(aaaa + i * n3 + kk * n2 + a * n1 + kk, x);
                            atomicAdd(aaaa + i * n3 + kk * n2 + kk * n1 + a, -x);
                            atomicAdd(aaaa + kk * n3 + i * n2 + kk * n1 + a, x);
                            atomicAdd(aaaa + kk * n3 + i * n2 + a * n1 + kk, -x);
                        }
                    }
                    for (int q = 0; q < nb; ++q) {
                        const long kk = rs.occ[1][q];
                        atomicAdd(abab + i * n3 + kk * n2 + a * n1 + kk, x);
                    }
                    break;
                }
                case T_SB: { // :849-873
                    const double x = apply_sign(cc, parity_single(rs.det[1], (int)i, (int)a));
                    atomicAdd(bb + i * n1 + a, x);
                    for (int q = 0; q < na; ++q) {
                        const long kk = rs.occ[0][q];
                        atomicAdd(abab + kk * n3 + i * n2 + kk * n1 + a, x);
                    }
                    for (int q = 0; q < nb; ++q) {
                        const long kk = rs.occ[1][q];
                        if (kk != i) {
                            atomicAdd(bbbb + i * n3 + kk * n2 + a * n1 + kk, x);
                            atomicAdd(bbbb + i * n3 + kk * n2 + kk * n1 + a, -x);
                            atomicAdd(bbbb + kk * n3 + i * n2 + kk * n1 + a, x);
                            atomicAdd(bbbb + kk * n3 + i * n2 + a * n1 + kk, -x);
                        }
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

// compute_overlap (overlap.cpp:21-31): sum_i c1_i c2_{index2(det1_i)}
template<int KM>
__global__ void __launch_bounds__(256) overlap_kernel(DetIndex<KM> index2, const u64 *__restrict__ dets1, int nwords,
                                                      long row0, long nloc, const double *__restrict__ c1,
                                                      const double *__restrict__ c2, double *out) {
    __shared__ double ws[8];
    double acc = 0.0;
    for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < nloc; r += (long)gridDim.x * blockDim.x) {
        const long i = row0 + r;
        const int j = index2.find(dets1[i * nwords], nwords == 2 ? dets1[i * nwords + 1] : 0ULL);
        if (j >= 0)
            acc = fma(c1[i], c2[j], acc);
    }
    for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0)
        ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w)
            t += ws[w];
        atomicAdd(out, t);
    }
}

template<int KIND, int KM>
int run_trdm(pyci_ctx *ctx, const pyci_wfn *wfn2, BuildParams &P, const double *c2) {
    const DetIndex<KM> ix = make_index<KM>(wfn2);
    const size_t pb = (pair_table_bytes(P) + 7) & ~(size_t)7;
    const size_t smem = pb + pair_mask_bytes(P, KIND);
    const long work = (long)P.ncand / 4;
    const int block = work <= 256 ? 32 : work <= 1024 ? 64 : work <= 4096 ? 128 : 256;
    int per_sm = 1;
    PYCI_CUDA(cudaFuncSetAttribute(trdm_kernel<KIND, KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PYCI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trdm_kernel<KIND, KM>, block, smem));
    const long grid = std::min<long>(P.nloc, (long)ctx->sm_count * std::max(per_sm, 1));
    if (grid > 0) {
        trdm_kernel<KIND, KM><<<(unsigned)grid, block, smem, ctx->stream>>>(P, ix, c2, (u32)pb);
        ctx->launches++;
    }
    PYCI_CUDA(cudaGetLastError());
    return PYCI_OK;
}

template<int KIND>
int trdm_dispatch(pyci_ctx *ctx, const pyci_wfn *wfn2, BuildParams &P, const double *c2) {
    switch (wfn2->keymode) {
    case KEY32:
        return run_trdm<KIND, KEY32>(ctx, wfn2, P, c2);
    case KEY64:
        return run_trdm<KIND, KEY64>(ctx, wfn2, P, c2);
    default:
        if constexpr (KIND == PYCI_FULLCI)
            return run_trdm<KIND, KEY128>(ctx, wfn2, P, c2);
        else
            PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "one-spin wave functions use 32- or 64-bit keys");
    }
}

} // namespace

int rdms_impl(pyci_ctx *ctx, const pyci_wfn *wfn, const double *coeffs, double *rdm1, double *rdm2) {
    PYCI_NVTX("pyci:compute_rdms");
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

namespace {
int same_space(const pyci_wfn *a, const pyci_wfn *b) {
    if (a->kind != b->kind || a->nbasis != b->nbasis || a->nocc_up != b->nocc_up || a->nocc_dn != b->nocc_dn)
        PYCI_FAIL(PYCI_ERR_VALUE, "the two wave functions differ in kind, basis size or occupation");
    return PYCI_OK;
}
} // namespace

int trdms_impl(pyci_ctx *ctx, const pyci_wfn *wfn1, const pyci_wfn *wfn2, const double *coeffs1, const double *coeffs2,
               double *rdm1, double *rdm2) {
    PYCI_TRY(same_space(wfn1, wfn2));
    BuildParams P;
    PYCI_TRY(enum_params_init(P, wfn1));
    const long n = wfn1->nbasis, n2 = n * n, n4 = n2 * n2;
    const int kind = wfn1->kind;
    const size_t s1 = (size_t)((kind == PYCI_FULLCI) ? 2 * n2 : n2);
    const size_t s2 = (size_t)((kind == PYCI_FULLCI) ? 3 * n4 : (kind == PYCI_DOCI) ? n2 : n4);
    const long R = ctx->nranks, ndet = wfn1->ndet;
    const long per = (ndet + R - 1) / R;
    P.row0 = std::min(ndet, per * ctx->rank);
    P.nloc = std::min(ndet, per * (ctx->rank + 1)) - P.row0;
    P.ncol = wfn2->ndet;
    double *dc1 = nullptr, *dc2 = nullptr, *d12 = nullptr;
    auto body = [&]() -> int {
        PYCI_CUDA(dev_malloc(&dc1, sizeof(double) * (size_t)std::max<long>(ndet, 1)));
        PYCI_CUDA(dev_malloc(&dc2, sizeof(double) * (size_t)std::max<long>(wfn2->ndet, 1)));
        PYCI_CUDA(dev_malloc(&d12, sizeof(double) * (s1 + s2)));
        PYCI_CUDA(cudaMemcpyAsync(dc1, coeffs1, sizeof(double) * ndet, cudaMemcpyHostToDevice, ctx->stream));
        PYCI_CUDA(cudaMemcpyAsync(dc2, coeffs2, sizeof(double) * wfn2->ndet, cudaMemcpyHostToDevice, ctx->stream));
        PYCI_CUDA(cudaMemsetAsync(d12, 0, sizeof(double) * (s1 + s2), ctx->stream));
        P.coeffs = dc1;
        P.rdm1 = d12;
        P.rdm2 = d12 + s1;
        if (kind == PYCI_DOCI)
            PYCI_TRY(trdm_dispatch<PYCI_DOCI>(ctx, wfn2, P, dc2));
        else if (kind == PYCI_FULLCI)
            PYCI_TRY(trdm_dispatch<PYCI_FULLCI>(ctx, wfn2, P, dc2));
        else
            PYCI_TRY(trdm_dispatch<PYCI_GENCI>(ctx, wfn2, P, dc2));
        PYCI_TRY(comm_allreduce_sum_f64(ctx, d12, (long)(s1 + s2)));
        PYCI_CUDA(cudaMemcpyAsync(rdm1, d12, sizeof(double) * s1, cudaMemcpyDeviceToHost, ctx->stream));
        PYCI_CUDA(cudaMemcpyAsync(rdm2, d12 + s1, sizeof(double) * s2, cudaMemcpyDeviceToHost, ctx->stream));
        PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
        return PYCI_OK;
    };
    const int rc = body();
    dev_free(dc1);
    dev_free(dc2);
    dev_free(d12);
    return rc;
}

int overlap_impl(pyci_ctx *ctx, const pyci_wfn *wfn1, const pyci_wfn *wfn2, const double *coeffs1, const double *coeffs2,
                 double *out) {
    PYCI_TRY(same_space(wfn1, wfn2));
    const long R = ctx->nranks, ndet = wfn1->ndet;
    const long per = (ndet + R - 1) / R;
    const long row0 = std::min(ndet, per * ctx->rank), nloc = std::min(ndet, per * (ctx->rank + 1)) - row0;
    double *dc1 = nullptr, *dc2 = nullptr, *acc = nullptr;
    auto body = [&]() -> int {
        PYCI_CUDA(dev_malloc(&dc1, sizeof(double) * (size_t)std::max<long>(ndet, 1)));
        PYCI_CUDA(dev_malloc(&dc2, sizeof(double) * (size_t)std::max<long>(wfn2->ndet, 1)));
        PYCI_CUDA(dev_malloc(&acc, sizeof(double)));
        PYCI_CUDA(cudaMemcpyAsync(dc1, coeffs1, sizeof(double) * ndet, cudaMemcpyHostToDevice, ctx->stream));
        PYCI_CUDA(cudaMemcpyAsync(dc2, coeffs2, sizeof(double) * wfn2->ndet, cudaMemcpyHostToDevice, ctx->stream));
        PYCI_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), ctx->stream));
        if (nloc > 0) {
            const unsigned blocks = (unsigned)std::min<long>((nloc + 255) / 256, (long)ctx->sm_count * 8);
            switch (wfn2->keymode) {
            case KEY32:
                overlap_kernel<KEY32><<<blocks, 256, 0, ctx->stream>>>(make_index<KEY32>(wfn2), wfn1->dets, wfn1->nwords, row0, nloc, dc1, dc2, acc);
                break;
            case KEY64:
                overlap_kernel<KEY64><<<blocks, 256, 0, ctx->stream>>>(make_index<KEY64>(wfn2), wfn1->dets, wfn1->nwords, row0, nloc, dc1, dc2, acc);
                break;
            default:
                overlap_kernel<KEY128><<<blocks, 256, 0, ctx->stream>>>(make_index<KEY128>(wfn2), wfn1->dets, wfn1->nwords, row0, nloc, dc1, dc2, acc);
                break;
            }
            ctx->launches++;
            PYCI_CUDA(cudaGetLastError());
        }
        PYCI_TRY(comm_allreduce_sum_f64(ctx, acc, 1));
        PYCI_CUDA(cudaMemcpyAsync(out, acc, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
        return PYCI_OK;
    };
    const int rc = body();
    dev_free(dc1);
    dev_free(dc2);
    dev_free(acc);
    return rc;
}
