// Connections of a SELECTED determinant space without one hash probe per candidate excitation.
//
// The reference finds the stored entries of a row by enumerating every single and double excitation of its
// determinant and looking each one up (sparseop.cpp:294-358,374-416,451-490; rdm.cpp:325-414): in a selected space of
// 64 spin-orbitals / 20 electrons that is 180 620 look-ups for ~10^2 hits.  Here the question is turned around:
// which determinants OF THE WAVE FUNCTION differ from the row's in two or four bit positions?  The bit positions
// (spin-orbitals) are dealt into NSEG = 6 segments.  Four differing positions touch at most four segments, so two
// connected determinants agree on at least two whole segments: for each of the C(6,2) = 15 segment pairs the
// determinants are bucketed by the bits of that pair (one counting sort per pair, keys hashed into ~2 ndet bins) and
// a row only meets the determinants of its own bucket -- a popcount of an XOR per meeting, exact, no false
// positives.  A pair of determinants is reported by the first segment pair (in lexicographic order) on which they
// agree, so every connection is found exactly once.  Cost per row ~ 15 x (bucket size) XOR/popcount tests instead
// of ncand hash probes.  Segments are dealt so that their occupation entropies balance (a sample of the
// determinants), which keeps the buckets small for spaces concentrated around a reference determinant.
//
// Used by the construction count pass (records (candidate index, column) for the fill pass) and by compute_rdms.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "enumerate.cuh"

int scan_counts(pyci_ctx *ctx, const int *cnt, long n, long *indptr, int *maxcnt);

namespace {

constexpr int JOIN_NSEG = 6;
constexpr int JOIN_NCOMBO = JOIN_NSEG * (JOIN_NSEG - 1) / 2;

struct JoinCombo {
    u64 m[2];        // positions of the two segments (word 0, word 1)
    u64 s1[2], s2[2]; // each segment alone
    u64 low[JOIN_NSEG][2]; // segments that precede s2 and are not s1: a pair that also agrees on one of them belongs
    int nlow;              // to an earlier segment pair
    u32 binmask;
};

enum { JOIN_HITLIST = 0, JOIN_RDM = 1, JOIN_STAGE = 2 };

__device__ __forceinline__ u32 join_key(const JoinCombo &C, u64 a, u64 b) {
    u64 x = (a & C.m[0]) * 0x9E3779B97F4A7C15ULL + (b & C.m[1]) * 0xC2B2AE3D27D4EB4FULL;
    x ^= x >> 29;
    x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 32;
    return (u32)x & C.binmask;
}

// occupation counts of every bit position over a strided sample of the determinants
__global__ void join_occupancy_kernel(const u64 *__restrict__ dets, int nw, long ndet, long stride, unsigned *out) {
    __shared__ unsigned h[128];
    if (threadIdx.x < 128)
        h[threadIdx.x] = 0;
    __syncthreads();
    for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * stride; i < ndet; i += (long)gridDim.x * blockDim.x * stride)
        for (int w = 0; w < nw; ++w)
            for (u64 q = dets[i * nw + w]; q; q &= q - 1)
                atomicAdd(&h[64 * w + __ffsll((long long)q) - 1], 1u);
    __syncthreads();
    if (threadIdx.x < 128 && h[threadIdx.x])
        atomicAdd(out + threadIdx.x, h[threadIdx.x]);
}

template<int NW>
__global__ void join_hist_kernel(const u64 *__restrict__ dets, long ndet, JoinCombo C, int *__restrict__ cnt) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ndet)
        return;
    const u64 a = dets[i * NW], b = (NW == 2) ? dets[i * NW + 1] : 0ULL;
    atomicAdd(cnt + join_key(C, a, b), 1);
}

// sum of squared bucket sizes = XOR/popcount tests a full pass over all rows would make
__global__ void join_sumsq_kernel(const int *__restrict__ cnt, long nbins, double *out) {
    double acc = 0.0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nbins; i += (long)gridDim.x * blockDim.x) {
        const double c = (double)cnt[i];
        acc += c * c;
    }
    for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc != 0.0)
        atomicAdd(out, acc);
}

template<int NW>
__global__ void join_scatter_kernel(const u64 *__restrict__ dets, long ndet, JoinCombo C, unsigned long long *cursor,
                                    u64 *__restrict__ sdet, u32 *__restrict__ sidx) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ndet)
        return;
    const u64 a = dets[i * NW], b = (NW == 2) ? dets[i * NW + 1] : 0ULL;
    const unsigned long long p = atomicAdd(cursor + join_key(C, a, b), 1ULL);
    if (NW == 2)
        reinterpret_cast<ulonglong2 *>(sdet)[p] = make_ulonglong2(a, b);
    else
        sdet[p] = a;
    sidx[p] = (u32)i;
}

__global__ void join_init_rowcnt_kernel(int *rowcnt, long row0, long nloc, long ncol) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nloc)
        rowcnt[r] = (row0 + r < ncol) ? 1 : 0; // the diagonal, sparseop.cpp:252-255 / :421-424 / :496-499
}

// excitation (A, B) -> (A2, B2) as the enumerator's code (type, i, a, k, l): i < k leave, a < l enter
template<int KIND>
__device__ __forceinline__ u32 join_pair_code(u64 A, u64 B, u64 A2, u64 B2) {
    const u64 ha = A & ~A2, pa = A2 & ~A;
    if (KIND == PYCI_FULLCI) {
        const u64 hb = B & ~B2, pb = B2 & ~B;
        if (ha && hb)
            return pack_code(T_AB, __ffsll((long long)ha) - 1, __ffsll((long long)pa) - 1, __ffsll((long long)hb) - 1,
                             __ffsll((long long)pb) - 1);
        if (hb) {
            const int i = __ffsll((long long)hb) - 1, a = __ffsll((long long)pb) - 1;
            const u64 h2 = hb & (hb - 1), p2 = pb & (pb - 1);
            if (!h2)
                return pack_code(T_SB, i, a, 0, 0);
            return pack_code(T_BB, i, a, __ffsll((long long)h2) - 1, __ffsll((long long)p2) - 1);
        }
    }
    const int i = __ffsll((long long)ha) - 1, a = __ffsll((long long)pa) - 1;
    const u64 h2 = ha & (ha - 1), p2 = pa & (pa - 1);
    if (!h2)
        return pack_code(T_SA, i, a, 0, 0);
    return pack_code(T_AA, i, a, __ffsll((long long)h2) - 1, __ffsll((long long)p2) - 1);
}

__device__ __forceinline__ u32 join_pos_occ(u64 D, int i) { return (u32)__popcll(D & ((1ULL << i) - 1ULL)); }
__device__ __forceinline__ u32 join_pos_vir(u64 D, int a) { return (u32)a - (u32)__popcll(D & ((1ULL << a) - 1ULL)); }

// the enumerator's candidate index of an excitation code of the row (A, B): inverse of decode<KIND>() (enumerate.cuh)
template<int KIND>
__device__ __forceinline__ u32 join_candidate(const BuildParams &P, u64 A, u64 B, u32 code) {
    const int type = code >> 24;
    const int i = (code >> 18) & 63, a = (code >> 12) & 63, k = (code >> 6) & 63, l = code & 63;
    const u32 nab = (KIND == PYCI_FULLCI) ? P.nAB : 0u, ndb = (KIND == PYCI_FULLCI) ? P.nDb : 0u;
    switch (type) {
    case T_AB:
        return (join_pos_occ(A, i) * (u32)P.nvir_a + join_pos_vir(A, a)) * P.nSb + join_pos_occ(B, k) * (u32)P.nvir_b +
               join_pos_vir(B, l);
    case T_AA: {
        const u32 xo = join_pos_occ(A, i), yo = join_pos_occ(A, k), xv = join_pos_vir(A, a), yv = join_pos_vir(A, l);
        return nab + (yo * (yo - 1) / 2 + xo) * P.nPva + yv * (yv - 1) / 2 + xv;
    }
    case T_BB: {
        const u32 xo = join_pos_occ(B, i), yo = join_pos_occ(B, k), xv = join_pos_vir(B, a), yv = join_pos_vir(B, l);
        return nab + P.nDa + (yo * (yo - 1) / 2 + xo) * P.nPvb + yv * (yv - 1) / 2 + xv;
    }
    case T_SA:
        return nab + P.nDa + ndb + join_pos_occ(A, i) * (u32)P.nvir_a + join_pos_vir(A, a);
    default: // T_SB
        return nab + P.nDa + ndb + P.nSa + join_pos_occ(B, i) * (u32)P.nvir_b + join_pos_vir(B, a);
    }
}

// eight symmetry-related positions of a same-spin contribution G[p,q,r,s] (rdm.cpp:399-414)
__device__ __forceinline__ void join_scatter8(double *G, long n, long p, long q, long r, long s, double x) {
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

// off-diagonal RDM contribution of the connected pair (row (A, B), partner by `code`), cc = c_row c_partner:
// the switch of rdm_kernel (rdm.cu; rdm.cpp:325-517), occupied orbitals read off the strings
template<int KIND>
__device__ void join_rdm_scatter(const BuildParams &P, u64 A, u64 B, u32 code, double cc) {
    const long n = P.n, n1 = n, n2 = n * n, n3 = n2 * n, n4 = n2 * n2;
    double *aa = P.rdm1, *bb = P.rdm1 + n2;
    double *aaaa = P.rdm2, *bbbb = P.rdm2 + n4, *abab = P.rdm2 + 2 * n4;
    const int type = code >> 24;
    const long i = (code >> 18) & 63, a = (code >> 12) & 63, k = (code >> 6) & 63, l = code & 63;
    switch (type) {
    case T_AB: {
        const double x = apply_sign(cc, parity_single(A, (int)i, (int)a) ^ parity_single(B, (int)k, (int)l));
        atomicAdd(abab + i * n3 + k * n2 + a * n1 + l, x);
        atomicAdd(abab + a * n3 + l * n2 + i * n1 + k, x);
        break;
    }
    case T_AA:
        join_scatter8(aaaa, n, i, k, a, l, apply_sign(cc, parity_double(A, (int)i, (int)k, (int)a, (int)l)));
        break;
    case T_BB:
        join_scatter8(bbbb, n, i, k, a, l, apply_sign(cc, parity_double(B, (int)i, (int)k, (int)a, (int)l)));
        break;
    case T_SA: {
        const double x = apply_sign(cc, parity_single(A, (int)i, (int)a));
        atomicAdd(aa + i * n1 + a, x);
        atomicAdd(aa + a * n1 + i, x);
        for (u64 q = A; q; q &= q - 1) {
            const long kk = __ffsll((long long)q) - 1;
            if (kk != i)
                join_scatter8(aaaa, n, i, kk, a, kk, x);
        }
        if (KIND == PYCI_FULLCI)
            for (u64 q = B; q; q &= q - 1) {
                const long kk = __ffsll((long long)q) - 1;
                atomicAdd(abab + i * n3 + kk * n2 + a * n1 + kk, x);
                atomicAdd(abab + a * n3 + kk * n2 + i * n1 + kk, x);
            }
        break;
    }
    case T_SB: {
        const double x = apply_sign(cc, parity_single(B, (int)i, (int)a));
        atomicAdd(bb + i * n1 + a, x);
        atomicAdd(bb + a * n1 + i, x);
        for (u64 q = A; q; q &= q - 1) {
            const long kk = __ffsll((long long)q) - 1;
            atomicAdd(abab + kk * n3 + i * n2 + kk * n1 + a, x);
            atomicAdd(abab + kk * n3 + a * n2 + kk * n1 + i, x);
        }
        for (u64 q = B; q; q &= q - 1) {
            const long kk = __ffsll((long long)q) - 1;
            if (kk != i)
                join_scatter8(bbbb, n, i, kk, a, kk, x);
        }
        break;
    }
    default:
        break;
    }
}

// One warp per row; rows are handed out by an atomic counter.  The lanes stride over the row's bucket.
// JOIN_HITLIST: append (candidate index, column) to hitlist[r][*] (up to cap) and count in rowcnt[r].
// JOIN_RDM: scatter c_row c_col sign for col > row.
// JOIN_STAGE: second pass for the rows whose hits overflowed the list (more than cap; the row pointer is known by
// now): (column, candidate index) go straight into the row's place in the CSR arrays -- cols[] and, as a bit pattern,
// vals[] -- where the fill pass picks them up, evaluates, sorts and overwrites them; rowcnt[] is a zeroed cursor.
template<int KIND, int MODE>
__global__ void __launch_bounds__(256) join_rows_kernel(BuildParams P, JoinCombo C, const long *__restrict__ start,
                                                        const u64 *__restrict__ sdet, const u32 *__restrict__ sidx,
                                                        uint2 *__restrict__ hitlist, int cap, int *__restrict__ rowcnt,
                                                        unsigned long long *next_row) {
    constexpr int NW = (KIND == PYCI_FULLCI) ? 2 : 1;
    const int lane = threadIdx.x & 31;
    const u32 lt = (1u << lane) - 1u;
    constexpr unsigned long long ROWS_PER_FETCH = 8; // rows a warp takes per atomic (one atomic per row: 18 % of the
                                                     // warp samples sat on it, ncu r2d)
    unsigned long long r64 = 0, rlast = 0;
    for (;;) {
        if (r64 == rlast) {
            if (lane == 0)
                r64 = atomicAdd(next_row, ROWS_PER_FETCH);
            r64 = __shfl_sync(0xffffffffu, r64, 0);
            rlast = min(r64 + ROWS_PER_FETCH, (unsigned long long)P.nloc);
            if (r64 >= (unsigned long long)P.nloc)
                return;
        }
        const long r = (long)r64++, row = P.row0 + r;
        const u64 A = __ldg(P.dets + row * NW), B = (NW == 2) ? __ldg(P.dets + row * NW + 1) : 0ULL;
        const u32 key = join_key(C, A, B);
        const long b0 = start[key], b1 = start[key + 1];
        const int ndiag = (row < P.ncol) ? 1 : 0;
        long stage0 = 0;
        if (MODE == JOIN_STAGE) {
            stage0 = P.indptr[r] + ndiag;
            if (P.indptr[r + 1] - stage0 <= (long)cap)
                continue; // recorded in full by the first pass
        }
        int cnt = (MODE == JOIN_HITLIST || MODE == JOIN_STAGE) ? rowcnt[r] : 0;
        const double ci = (MODE == JOIN_RDM) ? __ldg(P.coeffs + row) : 0.0;
        for (long base = b0; base < b1; base += 32) {
            const long p = base + lane;
            bool hit = false;
            u64 A2 = 0ULL, B2 = 0ULL;
            u32 j = 0;
            if (p < b1) {
                if (NW == 2) {
                    const ulonglong2 d = __ldg(reinterpret_cast<const ulonglong2 *>(sdet) + p);
                    A2 = d.x;
                    B2 = d.y;
                } else {
                    A2 = __ldg(sdet + p);
                }
                const u64 xa = A ^ A2, xb = B ^ B2;
                const int pc = __popcll(xa) + ((NW == 2) ? __popcll(xb) : 0);
                if (pc == 2 || pc == 4) {
                    // agrees on both segments of this pair (buckets are hashed) and on no earlier one
                    bool own = (((xa & C.m[0]) | (xb & C.m[1])) == 0ULL);
                    for (int q = 0; q < C.nlow; ++q)
                        own = own && (((xa & C.low[q][0]) | (xb & C.low[q][1])) != 0ULL);
                    if (own) {
                        j = __ldg(sidx + p);
                        hit = (MODE == JOIN_RDM) ? ((long)j > row) : ((long)j < P.ncol);
                    }
                }
            }
            if (MODE == JOIN_HITLIST) {
                const u32 msk = __ballot_sync(0xffffffffu, hit);
                if (msk) {
                    if (hit) {
                        const int slot = cnt - ndiag + __popc(msk & lt);
                        if (slot < cap)
                            hitlist[(size_t)r * cap + slot] =
                                make_uint2(join_candidate<KIND>(P, A, B, join_pair_code<KIND>(A, B, A2, B2)), j);
                    }
                    cnt += __popc(msk);
                }
            } else if (MODE == JOIN_STAGE) {
                const u32 msk = __ballot_sync(0xffffffffu, hit);
                if (msk) {
                    if (hit) {
                        const long at = stage0 + cnt + __popc(msk & lt);
                        P.cols[at] = (int)j;
                        P.vals[at] = __longlong_as_double(
                            (long long)join_candidate<KIND>(P, A, B, join_pair_code<KIND>(A, B, A2, B2)));
                    }
                    cnt += __popc(msk);
                }
            } else if (hit) {
                join_rdm_scatter<KIND>(P, A, B, join_pair_code<KIND>(A, B, A2, B2), ci * __ldg(P.coeffs + j));
            }
        }
        if ((MODE == JOIN_HITLIST || MODE == JOIN_STAGE) && lane == 0)
            rowcnt[r] = cnt;
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------

struct JoinPlan {
    u64 seg[JOIN_NSEG][2];
};

// deal the bit positions into segments of balanced occupation entropy: positions sorted by entropy, every round of
// NSEG positions goes to the NSEG segments in a pseudo-random order (correlated neighbours end up apart)
inline JoinPlan join_make_plan(const std::vector<unsigned> &occ, long nsample, int nbasis, int nw) {
    struct Pos {
        double h;
        int w, b;
    };
    std::vector<Pos> pos;
    for (int w = 0; w < nw; ++w)
        for (int b = 0; b < nbasis; ++b) {
            const double p = nsample > 0 ? (double)occ[64 * w + b] / (double)nsample : 0.5;
            const double h = (p <= 0.0 || p >= 1.0) ? 0.0 : -(p * std::log2(p) + (1.0 - p) * std::log2(1.0 - p));
            pos.push_back({h, w, b});
        }
    std::stable_sort(pos.begin(), pos.end(), [](const Pos &x, const Pos &y) { return x.h > y.h; });
    JoinPlan plan;
    memset(&plan, 0, sizeof(plan));
    u64 rng = 0x853c49e6748fea9bULL;
    for (size_t base = 0; base < pos.size(); base += JOIN_NSEG) {
        int perm[JOIN_NSEG];
        for (int s = 0; s < JOIN_NSEG; ++s)
            perm[s] = s;
        for (int s = JOIN_NSEG - 1; s > 0; --s) {
            rng ^= rng << 13;
            rng ^= rng >> 7;
            rng ^= rng << 17;
            std::swap(perm[s], perm[(int)(rng % (u64)(s + 1))]);
        }
        for (size_t q = base; q < std::min(pos.size(), base + JOIN_NSEG); ++q)
            plan.seg[perm[q - base]][pos[q].w] |= 1ULL << pos[q].b;
    }
    return plan;
}

inline JoinCombo join_make_combo(const JoinPlan &plan, int s1, int s2, u32 binmask) {
    JoinCombo C;
    memset(&C, 0, sizeof(C));
    for (int w = 0; w < 2; ++w) {
        C.s1[w] = plan.seg[s1][w];
        C.s2[w] = plan.seg[s2][w];
        C.m[w] = C.s1[w] | C.s2[w];
    }
    for (int s = 0; s < s2; ++s)
        if (s != s1) {
            C.low[C.nlow][0] = plan.seg[s][0];
            C.low[C.nlow][1] = plan.seg[s][1];
            ++C.nlow;
        }
    C.binmask = binmask;
    return C;
}

// Scratch of one join (freed by the caller through release()).
struct JoinScratch {
    int *cnt = nullptr;
    long *start = nullptr;
    unsigned long long *cursor = nullptr, *next_row = nullptr;
    u64 *sdet = nullptr;
    u32 *sidx = nullptr;
    unsigned *occ = nullptr;
    double *sumsq = nullptr;
    void release() {
        dev_free(cnt);
        dev_free(start);
        dev_free(cursor);
        dev_free(next_row);
        dev_free(sdet);
        dev_free(sidx);
        dev_free(occ);
        dev_free(sumsq);
        *this = JoinScratch();
    }
};

// Runs the join over rows [P.row0, P.row0 + P.nloc) of wfn.  *used = 0 (and nothing written) when the predicted
// number of XOR tests exceeds `budget_tests` (the caller then enumerates and probes as before).
// JOIN_HITLIST: rowcnt[nloc] receives the entries per row (diagonal included), hitlist[nloc][cap] the hits.
template<int KIND, int MODE>
int join_run(pyci_ctx *ctx, const pyci_wfn *wfn, const BuildParams &P, uint2 *hitlist, int cap, int *rowcnt,
             double budget_tests, int *used, double *tests_out) {
    constexpr int NW = (KIND == PYCI_FULLCI) ? 2 : 1;
    cudaStream_t st = ctx->stream;
    *used = 0;
    PYCI_NVTX("pyci:join(segment pairs)");
    const long ndet = wfn->ndet;
    if (ndet <= 0 || P.nloc <= 0)
        return PYCI_OK;
    long nbins = 1024;
    while (nbins < 2 * ndet && nbins < (1L << 26))
        nbins <<= 1;
    JoinScratch S;
    auto body = [&]() -> int {
        const unsigned dblocks = (unsigned)((ndet + 255) / 256);
        // ---- plan: occupation entropies from a sample
        PYCI_CUDA(dev_malloc(&S.occ, sizeof(unsigned) * 128));
        PYCI_CUDA(cudaMemsetAsync(S.occ, 0, sizeof(unsigned) * 128, st));
        const long stride = std::max<long>(1, ndet >> 18);
        const long nsample = (ndet + stride - 1) / stride;
        join_occupancy_kernel<<<(unsigned)std::min<long>((nsample + 255) / 256, 4L * ctx->sm_count), 256, 0, st>>>(
            wfn->dets, NW, ndet, stride, S.occ);
        ctx->launches++;
        std::vector<unsigned> hocc(128, 0u);
        PYCI_CUDA(cudaMemcpyAsync(hocc.data(), S.occ, sizeof(unsigned) * 128, cudaMemcpyDeviceToHost, st));
        PYCI_CUDA(cudaStreamSynchronize(st));
        const JoinPlan plan = join_make_plan(hocc, nsample, (int)wfn->nbasis, NW);
        JoinCombo combos[JOIN_NCOMBO];
        int nc = 0;
        for (int s1 = 0; s1 < JOIN_NSEG; ++s1)
            for (int s2 = s1 + 1; s2 < JOIN_NSEG; ++s2)
                combos[nc++] = join_make_combo(plan, s1, s2, (u32)(nbins - 1));
        PYCI_CUDA(dev_malloc(&S.cnt, sizeof(int) * (size_t)nbins));
        // ---- predicted work: sum over segment pairs of the squared bucket sizes, scaled to this rank's rows
        // (a JOIN_STAGE pass repeats a join that was already chosen: nothing to predict)
        if (MODE != JOIN_STAGE) {
        PYCI_CUDA(dev_malloc(&S.sumsq, sizeof(double)));
        PYCI_CUDA(cudaMemsetAsync(S.sumsq, 0, sizeof(double), st));
        for (int c = 0; c < nc; ++c) {
            PYCI_CUDA(cudaMemsetAsync(S.cnt, 0, sizeof(int) * (size_t)nbins, st));
            join_hist_kernel<NW><<<dblocks, 256, 0, st>>>(wfn->dets, ndet, combos[c], S.cnt);
            join_sumsq_kernel<<<ctx->sm_count * 2, 256, 0, st>>>(S.cnt, nbins, S.sumsq);
            ctx->launches += 2;
        }
        double sumsq = 0.0;
        PYCI_CUDA(cudaMemcpyAsync(&sumsq, S.sumsq, sizeof(double), cudaMemcpyDeviceToHost, st));
        PYCI_CUDA(cudaStreamSynchronize(st));
        const double tests = sumsq * (double)P.nloc / (double)ndet;
        if (tests_out)
            *tests_out = tests;
        if (tests > budget_tests)
            return PYCI_OK;
        }
        // ---- per segment pair: counting sort of the determinants by bucket, then every row meets its bucket
        PYCI_CUDA(dev_malloc(&S.start, sizeof(long) * (size_t)(nbins + 1)));
        PYCI_CUDA(dev_malloc(&S.cursor, sizeof(unsigned long long) * (size_t)nbins));
        PYCI_CUDA(dev_malloc(&S.next_row, sizeof(unsigned long long)));
        PYCI_CUDA(dev_malloc(&S.sdet, sizeof(u64) * (size_t)ndet * NW));
        PYCI_CUDA(dev_malloc(&S.sidx, sizeof(u32) * (size_t)ndet));
        if (MODE == JOIN_HITLIST) {
            join_init_rowcnt_kernel<<<(unsigned)((P.nloc + 255) / 256), 256, 0, st>>>(rowcnt, P.row0, P.nloc, P.ncol);
            ctx->launches++;
        }
        int per_sm = 1;
        PYCI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, join_rows_kernel<KIND, MODE>, 256, 0));
        const long grid = std::min<long>((P.nloc + 7) / 8, (long)ctx->sm_count * std::max(per_sm, 1));
        for (int c = 0; c < nc; ++c) {
            PYCI_CUDA(cudaMemsetAsync(S.cnt, 0, sizeof(int) * (size_t)nbins, st));
            join_hist_kernel<NW><<<dblocks, 256, 0, st>>>(wfn->dets, ndet, combos[c], S.cnt);
            ctx->launches++;
            PYCI_TRY(scan_counts(ctx, S.cnt, nbins, S.start, nullptr));
            PYCI_CUDA(cudaMemcpyAsync(S.cursor, S.start, sizeof(long) * (size_t)nbins, cudaMemcpyDeviceToDevice, st));
            join_scatter_kernel<NW><<<dblocks, 256, 0, st>>>(wfn->dets, ndet, combos[c], S.cursor, S.sdet, S.sidx);
            PYCI_CUDA(cudaMemsetAsync(S.next_row, 0, sizeof(unsigned long long), st));
            join_rows_kernel<KIND, MODE><<<(unsigned)grid, 256, 0, st>>>(P, combos[c], S.start, S.sdet, S.sidx, hitlist, cap,
                                                                       rowcnt, S.next_row);
            ctx->launches += 2;
        }
        PYCI_CUDA(cudaGetLastError());
        *used = 1;
        return PYCI_OK;
    };
    const int rc = body();
    S.release();
    return rc;
}

} // namespace
