// Fill pass for two-spin (FullCI) wave functions whose determinant array is sorted by (alpha string,
// beta string) as integers -- the order add_all_dets produces (twospinwfn.cpp:195-218: colex rank of alpha
// times C(n, nb) plus colex rank of beta; colex order of a combination IS the numeric order of its bit
// mask).  Column order inside a row is then the order of the excited pair (A', B'), and that order can be
// read off two short per-row lists instead of sorting the ~10^3 row entries (sort_row, sparseop.cpp:214-218):
//
//   alpha list: A itself, its nSa single and nDa double excitations   (La entries)
//   beta  list: B itself, its nSb single and nDb double excitations   (Lb entries)
//
// Every candidate of the row is a pair (g, h) of list entries: the row's entries with A' = A span the whole
// beta list, with A' a single excitation only B and its singles, with A' a double only B.  Ranking each
// list once per row (a bitmap over the C(n, nocc) colex ranks + prefix popcounts: rank = number of set
// bits below) gives every candidate its slot in sorted order in O(1):
//
//   slot(g, h) = offset_of_group[rank_alpha(g)] + rank of h inside the group's beta candidates.
//
// Column indices still come from the hash index, so the wave function does not have to be complete: entries
// are staged by slot in shared memory and compacted over the hit bitmap.  When it IS complete (every
// candidate hits) the slot is the final position and entries go straight to HBM (DIRECT).
#pragma once
#include "enumerate.cuh"

namespace {

struct SortedParams {
    u32 La, Lb;      // list lengths: 1 + nSa + nDa, 1 + nSb + nDb
    u32 Wa, Wb;      // bitmap words: ceil(C(n, nocc_a) / 32), ceil(C(n, nocc_b) / 32)
    u32 K1;          // binomial table row length (max(nocc_a, nocc_b) + 1)
    u32 M;           // candidate slots per row = ncand + 1
    u32 Nb;          // C(n, nocc_b): column = colex(A') * Nb + colex(B') in a complete sorted space
    const u32 *binom; // device: C(p, j), p < n, j < K1
};

__device__ __forceinline__ u32 colex_rank(u64 s, const u32 *__restrict__ binom, u32 K1) {
    u32 r = 0;
    int j = 1;
    for (u64 w = s; w; w &= w - 1, ++j)
        r += binom[(u32)(__ffsll((long long)w) - 1) * K1 + j];
    return r;
}

// out[w] = number of set bits in bm[0..w) ; one warp, W words
__device__ __forceinline__ void warp_popc_prefix(const u32 *bm, u32 *out, int W, int lane) {
    const int per = (W + 31) >> 5;
    const int lo = min(W, lane * per), hi = min(W, lo + per);
    u32 local = 0;
    for (int q = lo; q < hi; ++q)
        local += __popc(bm[q]);
    u32 incl = local;
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += t;
    }
    u32 run = incl - local;
    for (int q = lo; q < hi; ++q) {
        out[q] = run;
        run += __popc(bm[q]);
    }
}

// in-place exclusive scan of a[0..n) by one warp
__device__ __forceinline__ void warp_excl_scan(u32 *a, int n, int lane) {
    const int per = (n + 31) >> 5;
    const int lo = min(n, lane * per), hi = min(n, lo + per);
    u32 local = 0;
    for (int q = lo; q < hi; ++q)
        local += a[q];
    u32 incl = local;
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += t;
    }
    u32 run = incl - local;
    for (int q = lo; q < hi; ++q) {
        const u32 t = a[q];
        a[q] = run;
        run += t;
    }
}

__device__ __forceinline__ u32 bitmap_rank(const u32 *bm, const u32 *pf, u32 r) {
    return pf[r >> 5] + __popc(bm[r >> 5] & ((1u << (r & 31)) - 1u));
}

__host__ __device__ inline size_t sorted_smem_bytes(const SortedParams &S, u32 nSa, u32 nSb, u32 n, bool direct,
                                                    size_t pair_bytes) {
    size_t words = 2 * (size_t)S.Wa + 4 * (size_t)S.Wb + 2 * (size_t)S.La + (size_t)S.Lb + (nSb + 1) + (size_t)n * S.K1;
    if (!direct)
        words += S.M + 2 * (size_t)((S.M + 31) / 32);
    else
        words += (size_t)S.La + S.Lb;
    words = (words + 1) & ~(size_t)1;
    size_t bytes = 4 * words + 24 * (size_t)((nSa + 1) & ~1u) + 24 * (size_t)((nSb + 1) & ~1u) + pair_bytes;
    bytes = (bytes + 7) & ~(size_t)7;
    if (!direct)
        bytes += 8 * (size_t)S.M;
    return bytes;
}

template<int KM, bool DIRECT>
__global__ void __launch_bounds__(256) fill_sorted_kernel(BuildParams P, DetIndex<KM> index, SortedParams S, u32 nSa, u32 nSb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // ---- carve shared memory (all u32 arrays first, then 8-byte aligned tables)
    u32 *bmA = reinterpret_cast<u32 *>(smem_raw);
    u32 *pfA = bmA + S.Wa;
    u32 *bmB = pfA + S.Wa;
    u32 *pfB = bmB + S.Wb;
    u32 *bmB1 = pfB + S.Wb;
    u32 *pfB1 = bmB1 + S.Wb;
    u32 *slotA = pfB1 + S.Wb;  // [La]  colex rank, then sorted rank, then first slot of the entry's group
    u32 *gsz = slotA + S.La;   // [La]  group sizes in sorted alpha order -> group offsets
    u32 *slotB0 = gsz + S.La;  // [Lb]  colex rank, then slot inside the A' = A group
    u32 *r1 = slotB0 + S.Lb;   // [1+nSb] rank inside {B} u singles(B)
    u32 *binom = r1 + (nSb + 1);
    u32 *wend = binom + (u32)P.n * S.K1;
    u32 *scol = wend, *hitmap = wend, *hpref = wend;
    const u32 HW = (S.M + 31) / 32;
    u32 *crA = wend, *crB = wend; // DIRECT: colex ranks of the list entries (column = crA * Nb + crB)
    if (!DIRECT) {
        hitmap = scol + S.M;
        hpref = hitmap + HW;
        wend = hpref + HW;
    } else {
        crB = crA + S.La;
        wend = crB + S.Lb;
    }
    size_t off = ((size_t)(wend - bmA) + 1) & ~(size_t)1;
    unsigned char *tbase = smem_raw + 4 * off;
    const RowTables T = carve_tables(tbase, nSa, nSb);
    unsigned char *pbase = tbase + tables_bytes(nSa, nSb);
    uchar2 *pairs = reinterpret_cast<uchar2 *>(pbase);
    size_t pend = (size_t)(pbase - smem_raw) + sizeof(uchar2) * (size_t)(P.npairs_dim * (P.npairs_dim - 1) / 2 + 1);
    pend = (pend + 7) & ~(size_t)7;
    double *sval = reinterpret_cast<double *>(smem_raw + pend);

    __shared__ RowShared rs;
    __shared__ int low_count;
    fill_pairs(pairs, P.npairs_dim);
    for (u32 t = threadIdx.x; t < (u32)P.n * S.K1; t += blockDim.x)
        binom[t] = S.binom[t];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long n1 = P.n, n2 = n1 * n1, n3 = n2 * n1;
    const u32 nDa = P.nDa, nDb = P.nDb;

    for (long r = blockIdx.x; r < P.nloc; r += gridDim.x) {
        const long row = P.row0 + r;
        __syncthreads();
        row_setup(rs, P, row, 2);
        for (u32 t = threadIdx.x; t < 2 * S.Wa + 4 * S.Wb; t += blockDim.x)
            bmA[t] = 0;
        if (!DIRECT)
            for (u32 t = threadIdx.x; t < HW; t += blockDim.x)
                hitmap[t] = 0;
        __syncthreads();
        // ---- tables of single excitations (strings, values) and colex ranks of every list entry
        build_tables<PYCI_FULLCI, true>(P, rs, T, nSa, nSb);
        for (u32 t = threadIdx.x; t < nDa + nDb + 2; t += blockDim.x) {
            if (t < nDa) { // alpha double c' -> string
                const u32 po = fdiv(t, P.dPva), pv = t - po * P.nPva;
                const uchar2 o = pairs[po], v = pairs[pv];
                const u64 s = rs.det[0] ^ (1ULL << rs.occ[0][o.x]) ^ (1ULL << rs.occ[0][o.y]) ^ (1ULL << rs.vir[0][v.x]) ^
                              (1ULL << rs.vir[0][v.y]);
                const u32 cr = colex_rank(s, binom, S.K1);
                slotA[1 + nSa + t] = cr;
                atomicOr(&bmA[cr >> 5], 1u << (cr & 31));
            } else if (t < nDa + nDb) {
                const u32 tb = t - nDa;
                const u32 po = fdiv(tb, P.dPvb), pv = tb - po * P.nPvb;
                const uchar2 o = pairs[po], v = pairs[pv];
                const u64 s = rs.det[1] ^ (1ULL << rs.occ[1][o.x]) ^ (1ULL << rs.occ[1][o.y]) ^ (1ULL << rs.vir[1][v.x]) ^
                              (1ULL << rs.vir[1][v.y]);
                const u32 cr = colex_rank(s, binom, S.K1);
                slotB0[1 + nSb + tb] = cr;
                atomicOr(&bmB[cr >> 5], 1u << (cr & 31));
            } else if (t == nDa + nDb) {
                const u32 cr = colex_rank(rs.det[0], binom, S.K1);
                slotA[0] = cr;
                atomicOr(&bmA[cr >> 5], 1u << (cr & 31));
            } else {
                const u32 cr = colex_rank(rs.det[1], binom, S.K1);
                slotB0[0] = cr;
                atomicOr(&bmB[cr >> 5], 1u << (cr & 31));
                atomicOr(&bmB1[cr >> 5], 1u << (cr & 31));
            }
        }
        __syncthreads(); // single-excitation strings are in the tables now
        for (u32 t = threadIdx.x; t < nSa + nSb; t += blockDim.x) {
            if (t < nSa) {
                const u32 cr = colex_rank(T.sa_str[t], binom, S.K1);
                slotA[1 + t] = cr;
                atomicOr(&bmA[cr >> 5], 1u << (cr & 31));
            } else {
                const u32 tb = t - nSa;
                const u32 cr = colex_rank(T.sb_str[tb], binom, S.K1);
                slotB0[1 + tb] = cr;
                atomicOr(&bmB[cr >> 5], 1u << (cr & 31));
                atomicOr(&bmB1[cr >> 5], 1u << (cr & 31));
            }
        }
        __syncthreads();
        // ---- prefix popcounts of the three bitmaps (one warp each)
        for (int j = warp; j < 3; j += (int)(blockDim.x >> 5)) {
            if (j == 0)
                warp_popc_prefix(bmA, pfA, (int)S.Wa, lane);
            else if (j == 1)
                warp_popc_prefix(bmB, pfB, (int)S.Wb, lane);
            else
                warp_popc_prefix(bmB1, pfB1, (int)S.Wb, lane);
        }
        __syncthreads();
        // ---- sorted ranks; group sizes in sorted alpha order
        for (u32 t = threadIdx.x; t < S.La + S.Lb; t += blockDim.x) {
            if (t < S.La) {
                if (DIRECT)
                    crA[t] = slotA[t] * S.Nb;
                const u32 rk = bitmap_rank(bmA, pfA, slotA[t]);
                slotA[t] = rk;
                gsz[rk] = (t == 0) ? S.Lb : (t <= nSa) ? (1u + nSb) : 1u;
            } else {
                const u32 h = t - S.La;
                const u32 cr = slotB0[h];
                if (DIRECT)
                    crB[h] = cr;
                if (h <= nSb)
                    r1[h] = bitmap_rank(bmB1, pfB1, cr);
                slotB0[h] = bitmap_rank(bmB, pfB, cr);
            }
        }
        __syncthreads();
        if (warp == 0)
            warp_excl_scan(gsz, (int)S.La, lane);
        __syncthreads();
        const u32 base0 = gsz[slotA[0]]; // first slot of the A' = A group
        __syncthreads();
        for (u32 t = threadIdx.x; t < S.La + S.Lb; t += blockDim.x) {
            if (t < S.La)
                slotA[t] = gsz[slotA[t]];
            else
                slotB0[t - S.La] += base0;
        }
        if (threadIdx.x == 0)
            low_count = 0;
        __syncthreads();
        const u32 sdiag = slotB0[0];
        const long out0 = P.indptr[r];
        // ---- diagonal (sparseop.cpp:421-424)
        if (threadIdx.x == 0 && row < P.ncol) {
            if (DIRECT) {
                P.cols[out0 + sdiag] = (int)row;
                P.vals[out0 + sdiag] = P.diag[r];
            } else {
                scol[sdiag] = (u32)row;
                sval[sdiag] = P.diag[r];
                atomicOr(&hitmap[sdiag >> 5], 1u << (sdiag & 31));
            }
        }
        // ---- candidates
        for (u32 base = 0; base < P.ncand; base += UNROLL * blockDim.x) {
            int hit[UNROLL];
            double val[UNROLL];
            u32 slot[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                u32 c = base + u * blockDim.x + threadIdx.x;
                u32 ga = 0, hb = 0; // list entries (alpha, beta) of the candidate
                hit[u] = -1;
                val[u] = 0.0;
                slot[u] = 0;
                if (c < P.ncand) {
                    u64 A = rs.det[0], B = rs.det[1];
                    if (c < P.nAB) { // sparseop.cpp:318-337
                        const u32 sa = fdiv(c, P.dSb), sb = c - sa * P.nSb;
                        A = T.sa_str[sa];
                        B = T.sb_str[sb];
                        const int par = ((T.sa_meta[sa] ^ T.sb_meta[sb]) >> 16) & 1;
                        val[u] = apply_sign(__ldg(P.two_mo + (T.sa_off[sa] + T.sb_off[sb])), par);
                        slot[u] = slotA[1 + sa] + r1[1 + sb];
                        ga = 1 + sa;
                        hb = 1 + sb;
                    } else if ((c -= P.nAB) < nDa) { // sparseop.cpp:339-358
                        const u32 po = fdiv(c, P.dPva), pv = c - po * P.nPva;
                        const uchar2 o = pairs[po], v = pairs[pv];
                        const long i = rs.occ[0][o.x], k = rs.occ[0][o.y], a = rs.vir[0][v.x], l = rs.vir[0][v.y];
                        A ^= (1ULL << i) | (1ULL << k) | (1ULL << a) | (1ULL << l);
                        const long koff = n3 * i + n2 * k;
                        const double x = __ldg(P.two_mo + koff + n1 * a + l) - __ldg(P.two_mo + koff + n1 * l + a);
                        val[u] = apply_sign(x, parity_double(rs.det[0], (int)i, (int)k, (int)a, (int)l));
                        slot[u] = slotA[1 + nSa + c];
                        ga = 1 + nSa + c;
                    } else if ((c -= nDa) < nDb) { // sparseop.cpp:397-416
                        const u32 po = fdiv(c, P.dPvb), pv = c - po * P.nPvb;
                        const uchar2 o = pairs[po], v = pairs[pv];
                        const long i = rs.occ[1][o.x], k = rs.occ[1][o.y], a = rs.vir[1][v.x], l = rs.vir[1][v.y];
                        B ^= (1ULL << i) | (1ULL << k) | (1ULL << a) | (1ULL << l);
                        const long koff = n3 * i + n2 * k;
                        const double x = __ldg(P.two_mo + koff + n1 * a + l) - __ldg(P.two_mo + koff + n1 * l + a);
                        val[u] = apply_sign(x, parity_double(rs.det[1], (int)i, (int)k, (int)a, (int)l));
                        slot[u] = slotB0[1 + nSb + c];
                        hb = 1 + nSb + c;
                    } else if ((c -= nDb) < nSa) {
                        A = T.sa_str[c];
                        val[u] = T.sa_val[c];
                        slot[u] = slotA[1 + c] + r1[0];
                        ga = 1 + c;
                    } else {
                        c -= nSa;
                        B = T.sb_str[c];
                        val[u] = T.sb_val[c];
                        slot[u] = slotB0[1 + c];
                        hb = 1 + c;
                    }
                    // complete sorted space: the column IS the colex rank pair (twospinwfn.cpp:195-218), no probe
                    hit[u] = DIRECT ? (int)(crA[ga] + crB[hb]) : index.find(A, B);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (hit[u] >= 0 && hit[u] < P.ncol) {
                    if (DIRECT) {
                        P.cols[out0 + slot[u]] = hit[u];
                        P.vals[out0 + slot[u]] = val[u];
                    } else {
                        scol[slot[u]] = (u32)hit[u];
                        sval[slot[u]] = val[u];
                        atomicOr(&hitmap[slot[u] >> 5], 1u << (slot[u] & 31));
                    }
                }
            }
        }
        if (DIRECT) {
            if (threadIdx.x == 0)
                P.lowcnt[r] = (int)sdiag + 1; // every slot is filled: the diagonal's slot counts the entries before it
        } else {
            __syncthreads();
            if (warp == 0)
                warp_popc_prefix(hitmap, hpref, (int)HW, lane);
            __syncthreads();
            for (u32 s = threadIdx.x; s < S.M; s += blockDim.x) {
                const u32 wbits = hitmap[s >> 5];
                if ((wbits >> (s & 31)) & 1u) {
                    const u32 pos = hpref[s >> 5] + __popc(wbits & ((1u << (s & 31)) - 1u));
                    P.cols[out0 + pos] = (int)scol[s];
                    P.vals[out0 + pos] = sval[s];
                }
            }
            if (threadIdx.x == 0) {
                // entries with col <= row: hits among the slots up to and including the diagonal's
                const u32 wbits = hitmap[sdiag >> 5];
                const u32 upto = (sdiag & 31) == 31 ? 0xffffffffu : ((2u << (sdiag & 31)) - 1u);
                P.lowcnt[r] = (int)(hpref[sdiag >> 5] + __popc(wbits & upto));
            }
        }
    }
}

} // namespace
