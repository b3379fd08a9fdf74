// Fill pass for COMPLETE two-spin (FullCI) spaces in add_all_dets order (twospinwfn.cpp:195-218:
// row = colex(A) * C(n, nb) + colex(B)), square operator.  Every excitation of every row is present, the
// column of (A', B') is colex(A') * Nb + colex(B'), and the sorted row (sort_row, sparseop.cpp:214-218)
// factorises into data that depends on ONE string only:
//
//   sorted row of (A, B) = for A' in sorted {A, singles(A), doubles(A)}:
//        A' = A        -> every B' in sorted {B, singles(B), doubles(B)}     (Lb entries)
//        A' a single   -> every B' in sorted {B, singles(B)}                 (1 + nSb entries)
//        A' a double   -> B' = B                                             (1 entry)
//
// so a pre-pass builds one table per distinct string (C(n, nocc) of them, not ndet): its excitations sorted
// by colex rank with everything of the Slater-Condon element that depends on that string alone (same-spin
// double elements, parities, integral offsets, the per-string partial sums of the single-excitation
// elements in the reference's summation order) and the first output slot of every group.  The fill kernel
// then only combines one alpha-side entry (shared memory) with one beta-side entry (coalesced read of the
// beta string's table) per matrix element: one integral load per alpha-beta element (sparseop.cpp:318-337),
// consecutive lanes writing consecutive slots, and no sort, hash probe or per-row table construction.
#pragma once
#include "build_sorted.cuh"

namespace {

// Tables of one spin: N strings; per string L = 1 + nS + nD entries (the string, its single and its same-spin
// double excitations) sorted by colex rank, and the L1 = 1 + nS entries of the {string, singles} sub-list
// sorted likewise.  The same string serves as the alpha string of Nb rows and as the beta string of Na rows;
// both uses are tabulated.
struct StringTables {
    u32 N, L, L1, nocc, nS, nD;
    // ---- beta-side use: the A' = A group walks the whole list, an A' = single group walks the sub-list
    u32 *cr;       // [N][L]   colex rank | (1 << 31 if the entry is a double excitation)
    double *dval;  // [N][L]   double: signed same-spin element (sparseop.cpp:397-416); else 0
    uint2 *sub;    // [N][L1]  x: colex rank; y: parity << 31 | (n i + a) << 18 | (n^2 i + a), the beta half of
                   //          the two_mo offset of i -> a (and its offset inside a slice two_mo[i', :, a', :])
    u32 *pos1;     // [N][L1]  position in the full list | (n i + a) << 16
    double *terms; // [N][L1][nocc]  own-spin (J - K) terms of the single, one per occupied orbital (ascending)
    u32 *j1self;   // [N]      position of the string itself in the sub-list
    u32 *selfj;    // [N]      ... and in the full list
    // ---- alpha-side use: first output slot of the group each entry heads, by kind, in sorted order
    u32 *s_off, *s_aux, *s_cr; // [N][nS]  slot | parity << 31 | (n^3 i + n a) | colex rank
    double *s_pre;             // [N][nS]  one_mo[i,a] + own-spin (J - K) sum (sparseop.cpp:303-311)
    u32 *d_off, *d_cr;         // [N][nD]
    double *d_val;             // [N][nD]  signed same-spin element (sparseop.cpp:339-358)
    u32 *self_off;             // [N]
};

struct CompleteParams {
    StringTables A, B;
    u32 M;  // entries per row = ncand + 1
    u32 Nb; // C(n, nocc_b)
    FastDiv dSb;
    // launch-time constants of the fill kernel (kept out of registers: the compiler re-derives them per use)
    u32 nn, GP, GPnn; // n^2; walkers side by side in the alpha-beta segment = max(1, 256 / L1b); GP * nsl
    u32 GPw, GPwnn;   // the same for the warp-specialised kernel (192 alpha-beta threads per group)
    // slice of one alpha single i -> a in shared memory: two_mo[i, k, a, l] over (k, l).  When the integrals satisfy
    // <ik|al> = <il|ak> bit for bit (checked at upload; real orbitals do) only k <= l is kept, n (n + 1) / 2 doubles
    // at l (l + 1) / 2 + k instead of n^2: 51 KB instead of 98 KB at n = 16, which is a fourth row buffer per SM
    u32 nsl, packed;
    u32 l2hint;       // bulk stores of the rows carry an L2 evict-first policy (PYCI_B200_FILL_L2HINT=0 turns it off)
    FastDiv dL1b;     // division of the thread index by L1b
};

inline size_t string_table_smem(u32 W, u32 L, u32 n, u32 K1, size_t pair_bytes) {
    return 4 * (size_t)(4 * W + 3 * L + n * K1) + pair_bytes + 16;
}

// One CTA per string (CTA `bid` of `nb`, 128 threads).  spin = 0 reads the alpha string of row rank * stride,
// spin = 1 the beta string of row rank.  other_L / other_L1: sizes of the A' = A and A' = single groups when this
// string is the alpha side.
__device__ __forceinline__ void string_table_body(const BuildParams &P, const StringTables &T, int spin, long stride, u32 W,
                                                  u32 K1, const u32 *gbinom, u32 other_L, u32 other_L1, int packed, u32 bid,
                                                  u32 nb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u32 *bm = reinterpret_cast<u32 *>(smem_raw);
    u32 *pf = bm + W;
    u32 *bm1 = pf + W;
    u32 *pf1 = bm1 + W;
    u32 *crt = pf1 + W;    // [L] colex rank by enumeration index
    u32 *jp = crt + T.L;   // [L] sorted position by enumeration index
    u32 *gs = jp + T.L;    // [L] group size by sorted position -> first slot
    u32 *binom = gs + T.L; // [n][K1]
    uchar2 *pairs = reinterpret_cast<uchar2 *>(binom + (u32)P.n * K1);
    __shared__ unsigned char occ[64], vir[64];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long n1 = P.n, n2 = n1 * n1, n3 = n2 * n1;
    const u32 nocc = T.nocc, nvir = (u32)P.n - nocc;
    const u32 nS = T.nS, nD = T.nD;
    const u32 nPv = nvir * (nvir - 1) / 2 ? nvir * (nvir - 1) / 2 : 1;
    const double *one_mo = P.one_mo, *two_mo = P.two_mo;

    fill_pairs(pairs, P.npairs_dim);
    for (u32 t = threadIdx.x; t < (u32)P.n * K1; t += blockDim.x)
        binom[t] = gbinom[t];
    for (u32 s = bid; s < T.N; s += nb) {
        __syncthreads();
        const u64 S = (spin == 0) ? P.dets[2 * ((long)s * stride)] : P.dets[2 * (long)s + 1];
        if (threadIdx.x == 0) {
            int no = 0, nv = 0;
            for (int p = 0; p < P.n; ++p) {
                if ((S >> p) & 1ULL)
                    occ[no++] = (unsigned char)p;
                else
                    vir[nv++] = (unsigned char)p;
            }
        }
        for (u32 t = threadIdx.x; t < 4 * W; t += blockDim.x)
            bm[t] = 0;
        __syncthreads();
        // ---- colex ranks of every entry (enumeration order: self | singles | doubles)
        for (u32 t = threadIdx.x; t < T.L; t += blockDim.x) {
            u64 X = S;
            if (t >= 1 && t <= nS) {
                const u32 q = t - 1, io = q / nvir, ia = q - io * nvir;
                X ^= (1ULL << occ[io]) ^ (1ULL << vir[ia]);
            } else if (t > nS) {
                const u32 q = t - 1 - nS, po = q / nPv, pv = q - po * nPv;
                const uchar2 o = pairs[po], v = pairs[pv];
                X ^= (1ULL << occ[o.x]) ^ (1ULL << occ[o.y]) ^ (1ULL << vir[v.x]) ^ (1ULL << vir[v.y]);
            }
            const u32 cr = colex_rank(X, binom, K1);
            crt[t] = cr;
            atomicOr(&bm[cr >> 5], 1u << (cr & 31));
            if (t <= nS)
                atomicOr(&bm1[cr >> 5], 1u << (cr & 31));
        }
        __syncthreads();
        if (warp == 0)
            warp_popc_prefix(bm, pf, (int)W, lane);
        else if (warp == 1)
            warp_popc_prefix(bm1, pf1, (int)W, lane);
        __syncthreads();
        // ---- beta-side tables: entries to their sorted positions
        const size_t base = (size_t)s * T.L, base1 = (size_t)s * T.L1;
        const u32 j1s = bitmap_rank(bm1, pf1, crt[0]);
        for (u32 t = threadIdx.x; t < T.L; t += blockDim.x) {
            const u32 cr = crt[t];
            const u32 j = bitmap_rank(bm, pf, cr);
            const u32 j1 = bitmap_rank(bm1, pf1, cr); // sub-list entries below this one
            jp[t] = j;
            if (t == 0) {
                T.cr[base + j] = cr;
                T.dval[base + j] = 0.0;
                T.sub[base1 + j1] = make_uint2(cr, 0u);
                T.pos1[base1 + j1] = j;
                for (u32 q = 0; q < nocc; ++q)
                    T.terms[(base1 + j1) * nocc + q] = 0.0;
                T.selfj[s] = j;
                T.j1self[s] = j1;
                gs[j] = other_L;
            } else if (t <= nS) {
                const u32 q = t - 1, io = q / nvir, iv = q - io * nvir;
                const long i = occ[io], a = vir[iv];
                const u32 par = (u32)parity_single(S, (int)i, (int)a);
                // own-spin part of the single-excitation element (sparseop.cpp:303-311 / :389-394)
                double val1 = one_mo[n1 * i + a];
                const long ioff = n3 * i;
                for (u32 z = 0; z < nocc; ++z) {
                    const long kk = occ[z], koff = ioff + n2 * kk;
                    const double tz = two_mo[koff + n1 * a + kk] - two_mo[koff + n1 * kk + a];
                    T.terms[(base1 + j1) * nocc + z] = tz;
                    val1 += tz;
                }
                T.cr[base + j] = cr;
                T.dval[base + j] = 0.0;
                // (bits 18-29: offset of (i, a) inside a slice two_mo[i', :, a', :] -- n i + a, or the packed index
                // of the unordered pair when the slices keep k <= l only)
                const u32 hi = (u32)max(i, a), lo = (u32)min(i, a);
                const u32 in_slice = packed ? hi * (hi + 1) / 2 + lo : (u32)(n1 * i + a);
                T.sub[base1 + j1] = make_uint2(cr, (par << 31) | (in_slice << 18) | (u32)(n2 * i + a));
                T.pos1[base1 + j1] = j | ((u32)(n1 * i + a) << 16);
                const size_t g = (size_t)s * nS + (j1 - (j1 > j1s ? 1u : 0u));
                T.s_aux[g] = (par << 31) | (u32)(n3 * i + n1 * a);
                T.s_cr[g] = cr;
                T.s_pre[g] = val1;
                gs[j] = other_L1;
            } else { // same-spin double (sparseop.cpp:339-358 / :397-416)
                const u32 q = t - 1 - nS, po = q / nPv, pv = q - po * nPv;
                const uchar2 o = pairs[po], v = pairs[pv];
                const long i = occ[o.x], k = occ[o.y], a = vir[v.x], l = vir[v.y];
                const long koff = n3 * i + n2 * k;
                const double x = two_mo[koff + n1 * a + l] - two_mo[koff + n1 * l + a];
                const double val = apply_sign(x, parity_double(S, (int)i, (int)k, (int)a, (int)l));
                T.cr[base + j] = cr | (1u << 31);
                T.dval[base + j] = val;
                const size_t d = (size_t)s * nD + (j - j1);
                T.d_cr[d] = cr;
                T.d_val[d] = val;
                gs[j] = 1u;
            }
        }
        // ---- alpha-side: first slot of every group
        __syncthreads();
        if (warp == 0)
            warp_excl_scan(gs, (int)T.L, lane);
        __syncthreads();
        for (u32 t = threadIdx.x; t < T.L; t += blockDim.x) {
            const u32 j = jp[t], off = gs[j];
            const u32 j1 = bitmap_rank(bm1, pf1, crt[t]);
            if (t == 0)
                T.self_off[s] = off;
            else if (t <= nS)
                T.s_off[(size_t)s * nS + (j1 - (j1 > j1s ? 1u : 0u))] = off;
            else
                T.d_off[(size_t)s * nD + (j - j1)] = off;
        }
    }
}

// Everything the fill kernel reads that is not the determinant list, in one launch of independent CTAs: the string
// tables of both spins (CTAs [0, ga) and [ga, ga + gb)) and the diagonal H_ii of this rank's rows (the rest, 128 rows
// per CTA) -- three launches that each under-fill the GPU for ~20-50 us became one.
struct PrepParams {
    u32 ga, gb, Wa, Wb, K1, La, Lb, L1a, L1b;
    long stride_a;
    const u32 *binom;
    int packed;
};
__global__ void __launch_bounds__(128) complete_prep_kernel(BuildParams P, StringTables A, StringTables B, PrepParams Q) {
    if (blockIdx.x < Q.ga) {
        string_table_body(P, A, 0, Q.stride_a, Q.Wa, Q.K1, Q.binom, Q.Lb, Q.L1b, Q.packed, blockIdx.x, Q.ga);
    } else if (blockIdx.x < Q.ga + Q.gb) {
        string_table_body(P, B, 1, 1L, Q.Wb, Q.K1, Q.binom, Q.La, Q.L1a, Q.packed, blockIdx.x - Q.ga, Q.gb);
    } else {
        const long r = (long)(blockIdx.x - Q.ga - Q.gb) * 128 + threadIdx.x;
        if (r < P.nloc) {
            const long row = P.row0 + r;
            P.diag[r] = diag_twobody(P, P.dets[2 * row], P.dets[2 * row + 1]);
        }
    }
}

// shared memory of the fill kernel: alpha-side tables (+ the two_mo slice) once per CTA, one row buffer per group
struct CompleteSmem {
    size_t tables, slice, rowbuf, total;
    u32 MP;
};
__host__ __device__ inline CompleteSmem complete_smem(u32 nSa, u32 nDa, u32 n, u32 M, int groups, bool with_slice,
                                                      u32 nsl) {
    CompleteSmem L;
    L.tables = (16 * (size_t)nSa + 8 * (size_t)nDa + 8 * (size_t)(nSa + nDa + n * n) + 15) & ~(size_t)15;
    L.slice = with_slice ? (8 * (size_t)nSa * nsl + 15) & ~(size_t)15 : 0; // the row buffers behind it stay 16-byte aligned
    L.MP = (M + 8) & ~3u; // room for the alignment shift (<= 3 entries), multiple of 4 entries
    L.rowbuf = 12 * (size_t)L.MP;
    L.total = L.tables + L.slice + (size_t)groups * L.rowbuf;
    return L;
}

// shared memory -> global bulk copy of a finished row (one per array).  hint != 0: with an L2 evict-first policy -- the
// CSR is written once and not read by this kernel, so its lines should leave L2 first; measured on the bare store
// pattern (tools/write_bw.cu, 6 row buffers per SM): 5567 -> 6185 GB/s.
__device__ __forceinline__ void bulk_store_row(const void *gdst, u32 ssrc, u32 bytes, u32 hint) {
    if (hint) {
        u64 pol; // (made where it is used, by the one issuing thread: no register held across the row)
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(ssrc),
                     "r"(bytes), "l"(pol)
                     : "memory");
    } else {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
    }
}

__device__ __forceinline__ double flip_sign(double x, u32 signbit31) { // signbit31: 0 or 1 << 31
    return __hiloint2double(__double2hiint(x) ^ (int)signbit31, __double2loint(x));
}

// Two named barriers per group of 256 threads.  FREE: the row buffer may be overwritten (all eight warps wait;
// warp 0 gets there after the bulk stores of the previous row have read the buffer).  FULL: the row is
// complete -- warps 1-7 only arrive and move on to the next row's loads, warp 0 waits and launches the TMA.
template<int GT = 256>
__device__ __forceinline__ void bar_free_sync(int group) {
    asm volatile("bar.sync %0, %1;" ::"r"(2 * group + 1), "n"(GT) : "memory");
}
template<int GT = 256>
__device__ __forceinline__ void bar_full_sync(int group) {
    asm volatile("bar.sync %0, %1;" ::"r"(2 * group + 2), "n"(GT) : "memory");
}
template<int GT = 256>
__device__ __forceinline__ void bar_full_arrive(int group) {
    asm volatile("bar.arrive %0, %1;" ::"r"(2 * group + 2), "n"(GT) : "memory");
}

// One CTA per SM; rows [row0, row0 + nloc) are split into one contiguous range per CTA, so a CTA changes alpha
// string only every Nb rows.  The alpha-side tables of the current string -- and, when it fits (SLICE), the
// slice two_mo[i, :, a, :] of every alpha single i -> a, i.e. every integral the alpha-beta elements of these
// rows can touch -- are staged in shared memory once per alpha string.  The CTA is G groups of 256 threads;
// each group builds one row at a time in its shared-memory row buffer, segment by segment (alpha-beta doubles
// | the A' = A group | alpha-alpha doubles | alpha singles | beta singles + diagonal) so that the lanes of a
// warp run the same code, then streams the finished row to HBM with aligned 16-byte stores (scalar 4- and
// 8-byte stores to rows that start at arbitrary offsets reach only a third of the HBM write bandwidth).
// (Measured and dropped, profiles/r2e: letting all 256 threads of the group copy the finished row out with 16-byte
// st.global -- the faster of the two in the bare store probe tools/write_bw.cu, 6.19 against 5.85 TB/s -- costs a
// second full barrier per row and the copy on every warp's critical path: 7.29 ms against 5.09 ms with the bulk copies.
// Likewise measured and dropped, profiles/r2h: cp.async prefetch of the NEXT row's beta-side data (~5 KB) into two
// shared-memory buffers per group while the current row is built, so that no row waits for an L2 round trip: 5.89 ms
// against 5.04 ms -- the exposed load latency is not what sets the pace; the extra ~800 copy instructions per row are.)
// GT: threads per group (256 or 128).  The pace of a group is set by the dependency chain of its row, not by its
// instruction count, so when shared memory has room for more row buffers than 1024 / 256 groups can use, smaller groups
// -- more rows in flight per SM -- are tried (run_complete picks; PYCI_B200_FILL_GT overrides).
template<bool SLICE, int GT>
__global__ void __launch_bounds__(GT == 128 ? 896 : 1024, 1) fill_complete_kernel(BuildParams P, CompleteParams C, int G) {
    constexpr u32 NW = GT / 32; // warps per group (a power of two)
    constexpr int NQ = 512 / GT; // entries of the beta list a thread loads before the row barrier
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 nSa = C.A.nS, nDa = C.A.nD, Lb = C.B.L, L1b = C.B.L1, nb = C.B.nocc, M = C.M, Nb = C.Nb;
    const u32 nn = C.nn;
    const CompleteSmem SL = complete_smem(nSa, nDa, (u32)P.n, M, G, SLICE, C.nsl);
    const u32 nsl = C.nsl;
    uint4 *s_pack = reinterpret_cast<uint4 *>(smem_raw); // [nSa] first slot, colex(A') * Nb, n^3 i + n a, parity << 31
    uint2 *d_pack = reinterpret_cast<uint2 *>(s_pack + nSa); // [nDa] slot | colex(A') * Nb
    double *s_pre = reinterpret_cast<double *>(d_pack + nDa);
    double *d_val = s_pre + nSa;
    double *JA = d_val + nDa; // [n][n] one_mo[i,a] + sum_{k in A} <ik|ak>
    const double *slice = reinterpret_cast<const double *>(smem_raw + SL.tables); // [nSa][n][n]
    const int group = threadIdx.x / GT;
    const u32 t = threadIdx.x % GT, wq = (NW - 1u) - (t >> 5), lane = t & 31u;
    double *sval = reinterpret_cast<double *>(smem_raw + SL.tables + SL.slice + (size_t)group * SL.rowbuf);
    int *scol = reinterpret_cast<int *>(sval + SL.MP);

    const long n1 = P.n, n2 = n1 * n1, n3 = n2 * n1;
    const double *__restrict__ two_mo = P.two_mo;
    const long per = (P.nloc + gridDim.x - 1) / gridDim.x;
    const long rbeg = (long)blockIdx.x * per, rend = min(P.nloc, rbeg + per);
    if (rbeg >= rend)
        return;
    // alpha-beta doubles: a thread keeps ONE entry of the beta sub-list (L1b <= 256) and walks the alpha singles
    // g = gq, gq + GP, ...; 256 / L1b such walkers side by side
    const u32 GP = C.GP;
    const u32 gq = (L1b <= (u32)GT) ? fdiv(t, C.dL1b) : 0u;
    const u32 w0 = (L1b <= (u32)GT) ? t - gq * L1b : t;
    const bool ab_active = (L1b > (u32)GT) || gq < GP;
    const u32 Uda = (nDa + 31) >> 5, Usa = (nSa + 31) >> 5;
    // Work balance inside a group (all eight warps meet at the FULL barrier of every row, so the slowest warp sets
    // the pace): the alpha-beta walk loads every warp alike; the short segments are dealt to different warps through
    // rotated thread indices -- the second trip of the A' = A list to warps 3-4 (tA), the beta singles to warps 1-2
    // (tB), the alpha-side units from warp 7 downwards (wq) -- and warp 0 keeps the bulk stores.
    const u32 tA = (t + 5u * GT / 8u) % GT, tB = (t + 7u * GT / 8u) % GT;
    const u32 ra_first = (u32)((P.row0 + rbeg) / Nb), ra_last = (u32)((P.row0 + rend - 1) / Nb);
    for (u32 ra = ra_first; ra <= ra_last; ++ra) {
        // ---- stage the alpha string's tables
        __syncthreads();
        {
            const u64 Adet = P.dets[2 * ((long)ra * Nb)];
            const size_t bS = (size_t)ra * nSa, bD = (size_t)ra * nDa;
            for (u32 g = threadIdx.x; g < nSa; g += blockDim.x) {
                const u32 aux = C.A.s_aux[bS + g];
                s_pack[g] = make_uint4(C.A.s_off[bS + g], C.A.s_cr[bS + g] * Nb, aux & 0x7fffffffu, aux & 0x80000000u);
                s_pre[g] = C.A.s_pre[bS + g];
            }
            for (u32 d = threadIdx.x; d < nDa; d += blockDim.x) {
                d_pack[d] = make_uint2(C.A.d_off[bD + d], C.A.d_cr[bD + d] * Nb);
                d_val[d] = C.A.d_val[bD + d];
            }
            for (u32 q = threadIdx.x; q < nn; q += blockDim.x) { // sparseop.cpp:382-388
                const long i = q / (u32)n1, a = q - i * n1;
                double v = P.one_mo[n1 * i + a];
                for (u64 w = Adet; w; w &= w - 1) {
                    const long kk = __ffsll((long long)w) - 1;
                    v += two_mo[n3 * i + n2 * kk + n1 * a + kk];
                }
                JA[q] = v;
            }
            if (SLICE) {
                double *wslice = const_cast<double *>(slice);
                for (u32 q = threadIdx.x; q < nSa * nsl; q += blockDim.x) {
                    const u32 g = q / nsl, kl = q - g * nsl;
                    u32 k, l;
                    if (C.packed) { // kl = l (l + 1) / 2 + k, k <= l
                        l = (u32)((sqrtf(8.0f * (float)kl + 1.0f) - 1.0f) * 0.5f);
                        while ((l + 1) * (l + 2) / 2 <= kl)
                            ++l;
                        while (l * (l + 1) / 2 > kl)
                            --l;
                        k = kl - l * (l + 1) / 2;
                    } else {
                        k = kl / (u32)n1;
                        l = kl - k * (u32)n1;
                    }
                    wslice[q] = two_mo[(C.A.s_aux[bS + g] & 0x7fffffffu) + n2 * k + l];
                }
            }
        }
        const u32 self_off = C.A.self_off[ra], self_colbase = ra * Nb;
        __syncthreads();
        const long blo = max(rbeg, (long)ra * Nb - P.row0), bhi = min(rend, (long)(ra + 1) * Nb - P.row0);
        u32 rb = (u32)(P.row0 + blo + group - (long)ra * Nb);
        for (long r = blo + group; r < bhi; r += G, rb += (u32)G) {
            // ---- this row's beta-side data: every global load is issued before anything waits
            const u32 *__restrict__ crB = C.B.cr + rb * Lb;
            const double *__restrict__ dvalB = C.B.dval + rb * Lb;
            const uint2 *__restrict__ subB = C.B.sub + rb * L1b;
            const u64 Bdet = __ldg(P.dets + 2 * (P.row0 + r) + 1);
            const u32 j1s = __ldg(C.B.j1self + rb);
            uint2 eb = make_uint2(0u, 0u);
            if (ab_active && w0 < L1b)
                eb = __ldg(subB + w0);
            u32 cb[NQ];
            double dv[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) { // the first 512 entries of the beta list (the rest, if any, below)
                cb[q] = 0u;
                dv[q] = 0.0;
                if (tA + (u32)GT * q < Lb) {
                    cb[q] = __ldg(crB + tA + (u32)GT * q);
                    dv[q] = __ldg(dvalB + tA + (u32)GT * q);
                }
            }
            u32 ps = 0u, sgn1 = 0u;
            double tq[4] = {0.0, 0.0, 0.0, 0.0}, diag_r = 0.0;
            if (tB < L1b) { // beta single tB of the sub-list (:382-394): position, parity, its first own-spin terms
                ps = __ldg(C.B.pos1 + rb * L1b + tB);
                sgn1 = __ldg(subB + tB).y & 0x80000000u;
                const double *tb = C.B.terms + (size_t)(rb * L1b + tB) * nb;
#pragma unroll
                for (u32 q = 0; q < 4; ++q)
                    if (q < nb)
                        tq[q] = __ldg(tb + q);
                diag_r = __ldg(P.diag + r);
            }
            const long out0 = r * (long)M; // complete space: every row holds M entries
            const u32 ov = (u32)out0 & 1u, oc = (u32)out0 & 3u;
            // row buffer shifted so that shared and global addresses share their 16-byte phase
            double *bval = sval + ov;
            int *bcol = scol + oc;
            // the bulk stores of the previous row must have read the buffer before it is overwritten
            if (t < 32) {
                if (t == 0)
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
            bar_free_sync<GT>(group);
            // (the prefetched beta-side values are consumed first so that their registers are free in the walk below)
            // ---- A' = A: columns of the whole beta list, values of its doubles (:397-416)
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const u32 w = tA + (u32)GT * q;
                if (w < Lb) {
                    bcol[self_off + w] = (int)(self_colbase + (cb[q] & 0x7fffffffu));
                    if (cb[q] >> 31)
                        bval[self_off + w] = dv[q];
                }
            }
            for (u32 w = tA + 512u; w < Lb; w += GT) {
                const u32 c2 = __ldg(crB + w);
                bcol[self_off + w] = (int)(self_colbase + (c2 & 0x7fffffffu));
                if (c2 >> 31)
                    bval[self_off + w] = __ldg(dvalB + w);
            }
            // ---- values of the beta singles (:382-394) and of the diagonal (:421-424)
            for (u32 j1 = tB; j1 < L1b; j1 += GT) {
                if (j1 != tB) {
                    ps = __ldg(C.B.pos1 + rb * L1b + j1);
                    sgn1 = __ldg(subB + j1).y & 0x80000000u;
                }
                const u32 slot = self_off + (ps & 0xffffu);
                if (j1 == j1s) {
                    bval[slot] = (j1 == tB) ? diag_r : P.diag[r];
                } else {
                    const double *tb = C.B.terms + (size_t)(rb * L1b + j1) * nb;
                    double v = JA[ps >> 16];
                    if (j1 == tB) {
#pragma unroll
                        for (u32 q = 0; q < 4; ++q)
                            if (q < nb)
                                v += tq[q];
                        for (u32 q = 4; q < nb; ++q)
                            v += __ldg(tb + q);
                    } else {
                        for (u32 q = 0; q < nb; ++q)
                            v += __ldg(tb + q);
                    }
                    bval[slot] = flip_sign(v, sgn1);
                }
            }
            // ---- alpha-beta doubles (sparseop.cpp:318-337): per element two shared loads (alpha entry, integral),
            // one add (column), one xor (sign) and two shared stores through pointers that already hold the beta
            // entry's position w -- everything that depends on w alone is hoisted out of the walk over g
            if (ab_active) {
                for (u32 w = w0; w < L1b; w += GT) {
                    if (w != w0)
                        eb = __ldg(subB + w); // colex rank | parity << 31, n i + a << 18, n^2 i + a
                    if (w == j1s)
                        continue; // B' = B: the alpha single below
                    const u32 kl = SLICE ? ((eb.y >> 18) & 0xfffu) : (eb.y & 0x3ffffu);
                    const u32 sgn_b = eb.y & 0x80000000u, cr_b = eb.x;
                    int *pc = bcol + w;
                    double *pv = bval + w;
                    const uint4 *pa = s_pack + gq;
                    const double *psl = slice + gq * nsl + kl;
#pragma unroll 4
                    for (u32 g = gq; g < nSa; g += GP, pa += GP, psl += C.GPnn) {
                        const uint4 a = *pa;
                        const double v = SLICE ? *psl : __ldg(two_mo + (a.z + kl));
                        pc[a.x] = (int)(a.y + cr_b);
                        pv[a.x] = flip_sign(v, a.w ^ sgn_b);
                    }
                }
            }
            // The alpha-side segments in units of 32 entries, dealt round-robin to the warps from the last one
            // down (it holds the idle lanes of the walk above): warp-uniform control flow.
            u32 ubase = 0;
            // ---- alpha-alpha doubles (:339-358)
            for (u32 u = (wq - ubase) & (NW - 1u); u < Uda; u += NW) {
                const u32 d = u * 32 + lane;
                if (d < nDa) {
                    const uint2 dp = d_pack[d];
                    bcol[dp.x] = (int)(dp.y + rb);
                    bval[dp.x] = d_val[d];
                }
            }
            ubase += Uda;
            // ---- alpha singles (:303-315)
            for (u32 u = (wq - ubase) & (NW - 1u); u < Usa; u += NW) {
                const u32 g = u * 32 + lane;
                if (g < nSa) {
                    const uint4 a = s_pack[g];
                    double v = s_pre[g];
                    for (u64 q = Bdet; q; q &= q - 1) {
                        const u32 kk = (u32)__ffsll((long long)q) - 1u;
                        v += SLICE ? slice[g * nsl + (C.packed ? kk * (kk + 1) / 2 + kk : kk * (u32)n1 + kk)]
                                   : __ldg(two_mo + a.z + (u32)n2 * kk + kk);
                    }
                    const u32 slot = a.x + j1s;
                    bcol[slot] = (int)(a.y + rb);
                    bval[slot] = flip_sign(v, a.w);
                }
            }
            // ---- the finished row goes out as two bulk copies (TMA): 16-byte aligned bodies of the value and
            // column streams; the few entries before / after the aligned bodies by scalar stores
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (t >= 32) {
                bar_full_arrive<GT>(group);
            } else {
                bar_full_sync<GT>(group);
                const u32 hv = ov, nv = (M - hv) >> 1;                    // values: pairs
                const u32 hc = min(M, (4u - oc) & 3u), nc = (M - hc) >> 2; // columns: quads
                if (t == 0) {
                    if (nv)
                        bulk_store_row(P.vals + out0 + hv, (u32)__cvta_generic_to_shared(bval + hv), nv * 16u, C.l2hint);
                    if (nc)
                        bulk_store_row(P.cols + out0 + hc, (u32)__cvta_generic_to_shared(bcol + hc), nc * 16u, C.l2hint);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                // the few entries before / after the aligned bodies (read before warp 0 frees the buffer again)
                if (t >= 4 && t < 4 + hv)
                    P.vals[out0 + t - 4] = bval[t - 4];
                if (t >= 8 && t - 8 + hv + 2 * nv < M)
                    P.vals[out0 + hv + 2 * nv + t - 8] = bval[hv + 2 * nv + t - 8];
                if (t >= 12 && t - 12 < hc)
                    P.cols[out0 + t - 12] = bcol[t - 12];
                if (t >= 16 && t < 20 && t - 16 + hc + 4 * nc < M)
                    P.cols[out0 + hc + 4 * nc + t - 16] = bcol[hc + 4 * nc + t - 16];
                if (t == 20)
                    P.lowcnt[r] = (int)(self_off + __ldg(C.B.selfj + rb)) + 1; // slots up to and including the diagonal
            }
        }
    }
    if (threadIdx.x % GT == 0)
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // all rows written before the CTA retires
}


// ---- warp-specialised form ------------------------------------------------------------------------------------------
// Same row layout, tables and shared-memory staging as fill_complete_kernel, but the eight warps of a group keep
// FIXED ROLES for every row instead of all passing through every segment's code with most lanes idle:
//   warps 0-5  the A' = A group (columns of the whole beta list, values of its doubles; its global loads are issued
//              before the row barrier, two entries per thread) and the alpha-beta doubles (72-76 % of a row):
//              192 / L1b walkers side by side
//   warp  6    the beta singles and the diagonal
//   warp  7    alpha-alpha doubles, alpha singles, then -- after the FULL barrier -- the two bulk stores of the row
// A role's per-row set-up is only what that role needs, so the straight-line overhead that every warp used to pay
// for every segment (three quarters of the instructions of fill_complete_kernel, ncu r1u) is paid once per row.
template<bool SLICE>
__global__ void __launch_bounds__(1024, 1) fill_complete_ws_kernel(BuildParams P, CompleteParams C, int G) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 nSa = C.A.nS, nDa = C.A.nD, Lb = C.B.L, L1b = C.B.L1, nb = C.B.nocc, M = C.M, Nb = C.Nb;
    const u32 nn = C.nn;
    const CompleteSmem SL = complete_smem(nSa, nDa, (u32)P.n, M, G, SLICE, C.nsl);
    const u32 nsl = C.nsl;
    uint4 *s_pack = reinterpret_cast<uint4 *>(smem_raw); // [nSa] first slot, colex(A') * Nb, n^3 i + n a, parity << 31
    uint2 *d_pack = reinterpret_cast<uint2 *>(s_pack + nSa); // [nDa] slot | colex(A') * Nb
    double *s_pre = reinterpret_cast<double *>(d_pack + nDa);
    double *d_val = s_pre + nSa;
    double *JA = d_val + nDa; // [n][n] one_mo[i,a] + sum_{k in A} <ik|ak>
    const double *slice = reinterpret_cast<const double *>(smem_raw + SL.tables); // [nSa][n][n]
    const int group = threadIdx.x >> 8;
    const u32 t = threadIdx.x & 255u, warp = t >> 5, lane = t & 31u;
    double *sval = reinterpret_cast<double *>(smem_raw + SL.tables + SL.slice + (size_t)group * SL.rowbuf);
    int *scol = reinterpret_cast<int *>(sval + SL.MP);

    const long n1 = P.n, n2 = n1 * n1, n3 = n2 * n1;
    const double *__restrict__ two_mo = P.two_mo;
    const long per = (P.nloc + gridDim.x - 1) / gridDim.x;
    const long rbeg = (long)blockIdx.x * per, rend = min(P.nloc, rbeg + per);
    if (rbeg >= rend)
        return;
    // alpha-beta role: a thread keeps ONE entry of the beta sub-list and walks the alpha singles g = gq, gq + GP, ...
    const u32 GP = C.GPw;
    const u32 gq = (L1b <= 192u) ? fdiv(t, C.dL1b) : 0u;
    const u32 w0 = (L1b <= 192u) ? t - gq * L1b : t;
    const bool ab_active = (L1b > 192u) || gq < GP;
    const u32 ra_first = (u32)((P.row0 + rbeg) / Nb), ra_last = (u32)((P.row0 + rend - 1) / Nb);
    for (u32 ra = ra_first; ra <= ra_last; ++ra) {
        // ---- stage the alpha string's tables (all threads of the CTA)
        __syncthreads();
        {
            const u64 Adet = P.dets[2 * ((long)ra * Nb)];
            const size_t bS = (size_t)ra * nSa, bD = (size_t)ra * nDa;
            for (u32 g = threadIdx.x; g < nSa; g += blockDim.x) {
                const u32 aux = C.A.s_aux[bS + g];
                s_pack[g] = make_uint4(C.A.s_off[bS + g], C.A.s_cr[bS + g] * Nb, aux & 0x7fffffffu, aux & 0x80000000u);
                s_pre[g] = C.A.s_pre[bS + g];
            }
            for (u32 d = threadIdx.x; d < nDa; d += blockDim.x) {
                d_pack[d] = make_uint2(C.A.d_off[bD + d], C.A.d_cr[bD + d] * Nb);
                d_val[d] = C.A.d_val[bD + d];
            }
            for (u32 q = threadIdx.x; q < nn; q += blockDim.x) { // sparseop.cpp:382-388
                const long i = q / (u32)n1, a = q - i * n1;
                double v = P.one_mo[n1 * i + a];
                for (u64 w = Adet; w; w &= w - 1) {
                    const long kk = __ffsll((long long)w) - 1;
                    v += two_mo[n3 * i + n2 * kk + n1 * a + kk];
                }
                JA[q] = v;
            }
            if (SLICE) {
                double *wslice = const_cast<double *>(slice);
                for (u32 q = threadIdx.x; q < nSa * nsl; q += blockDim.x) {
                    const u32 g = q / nsl, kl = q - g * nsl;
                    u32 k, l;
                    if (C.packed) { // kl = l (l + 1) / 2 + k, k <= l
                        l = (u32)((sqrtf(8.0f * (float)kl + 1.0f) - 1.0f) * 0.5f);
                        while ((l + 1) * (l + 2) / 2 <= kl)
                            ++l;
                        while (l * (l + 1) / 2 > kl)
                            --l;
                        k = kl - l * (l + 1) / 2;
                    } else {
                        k = kl / (u32)n1;
                        l = kl - k * (u32)n1;
                    }
                    wslice[q] = two_mo[(C.A.s_aux[bS + g] & 0x7fffffffu) + n2 * k + l];
                }
            }
        }
        const u32 self_off = C.A.self_off[ra], self_colbase = ra * Nb;
        __syncthreads();
        const long blo = max(rbeg, (long)ra * Nb - P.row0), bhi = min(rend, (long)(ra + 1) * Nb - P.row0);
        u32 rb = (u32)(P.row0 + blo + group - (long)ra * Nb);
        for (long r = blo + group; r < bhi; r += G, rb += (u32)G) {
            const long out0 = r * (long)M; // complete space: every row holds M entries
            const u32 ov = (u32)out0 & 1u, oc = (u32)out0 & 3u;
            // row buffer shifted so that shared and global addresses share their 16-byte phase
            double *bval = sval + ov;
            int *bcol = scol + oc;
            if (warp < 6u) {
                // ================= warps 0-5: the A' = A group (columns of the whole beta list, values of its
                // doubles, sparseop.cpp:397-416), then the alpha-beta doubles (:318-337).  Every global load of the
                // row is issued before the barrier.
                const u32 *__restrict__ crB = C.B.cr + rb * Lb;
                const double *__restrict__ dvalB = C.B.dval + rb * Lb;
                const uint2 *__restrict__ subB = C.B.sub + rb * L1b;
                const u32 j1s = __ldg(C.B.j1self + rb);
                uint2 eb = make_uint2(0u, 0u);
                if (ab_active && w0 < L1b)
                    eb = __ldg(subB + w0);
                u32 cb[2] = {0u, 0u};
                double dv[2] = {0.0, 0.0};
#pragma unroll
                for (int q = 0; q < 2; ++q) // the first 384 entries of the beta list (the rest, if any, below)
                    if (t + 192u * q < Lb) {
                        cb[q] = __ldg(crB + t + 192u * q);
                        dv[q] = __ldg(dvalB + t + 192u * q);
                    }
                bar_free_sync(group);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const u32 w = t + 192u * q;
                    if (w < Lb) {
                        bcol[self_off + w] = (int)(self_colbase + (cb[q] & 0x7fffffffu));
                        if (cb[q] >> 31)
                            bval[self_off + w] = dv[q];
                    }
                }
                for (u32 w = t + 384u; w < Lb; w += 192u) {
                    const u32 c2 = __ldg(crB + w);
                    bcol[self_off + w] = (int)(self_colbase + (c2 & 0x7fffffffu));
                    if (c2 >> 31)
                        bval[self_off + w] = __ldg(dvalB + w);
                }
                if (ab_active) {
                    for (u32 w = w0; w < L1b; w += 192) {
                        if (w != w0)
                            eb = __ldg(subB + w); // colex rank | parity << 31, n i + a << 18, n^2 i + a
                        if (w == j1s)
                            continue; // B' = B: the alpha single of warp 7
                        const u32 kl = SLICE ? ((eb.y >> 18) & 0xfffu) : (eb.y & 0x3ffffu);
                        const u32 sgn_b = eb.y & 0x80000000u, cr_b = eb.x;
                        int *pc = bcol + w;
                        double *pv = bval + w;
                        const uint4 *pa = s_pack + gq;
                        const double *psl = slice + gq * nsl + kl;
#pragma unroll 4
                        for (u32 g = gq; g < nSa; g += GP, pa += GP, psl += C.GPwnn) {
                            const uint4 a = *pa;
                            const double v = SLICE ? *psl : __ldg(two_mo + (a.z + kl));
                            pc[a.x] = (int)(a.y + cr_b);
                            pv[a.x] = flip_sign(v, a.w ^ sgn_b);
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bar_full_arrive(group);
            } else if (warp == 6u) {
                // ================= warp 6: values of the beta singles (:382-394) and of the diagonal (:421-424);
                // two sub-list entries per lane are loaded before the barrier
                const uint2 *__restrict__ subB = C.B.sub + rb * L1b;
                const u32 j1s = __ldg(C.B.j1self + rb);
                const double diag_r = __ldg(P.diag + r);
                u32 ps[2] = {0u, 0u}, sg[2] = {0u, 0u};
                double tq[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
#pragma unroll
                for (int z = 0; z < 2; ++z) {
                    const u32 j1 = lane + 32u * z;
                    if (j1 < L1b) {
                        ps[z] = __ldg(C.B.pos1 + rb * L1b + j1);
                        sg[z] = __ldg(subB + j1).y & 0x80000000u;
                        const double *tb = C.B.terms + (size_t)(rb * L1b + j1) * nb;
#pragma unroll
                        for (u32 q = 0; q < 4; ++q)
                            if (q < nb)
                                tq[z][q] = __ldg(tb + q);
                    }
                }
                bar_free_sync(group);
#pragma unroll
                for (int z = 0; z < 2; ++z) {
                    const u32 j1 = lane + 32u * z;
                    if (j1 < L1b) {
                        const u32 slot = self_off + (ps[z] & 0xffffu);
                        if (j1 == j1s) {
                            bval[slot] = diag_r;
                        } else {
                            const double *tb = C.B.terms + (size_t)(rb * L1b + j1) * nb;
                            double v = JA[ps[z] >> 16];
#pragma unroll
                            for (u32 q = 0; q < 4; ++q)
                                if (q < nb)
                                    v += tq[z][q];
                            for (u32 q = 4; q < nb; ++q)
                                v += __ldg(tb + q);
                            bval[slot] = flip_sign(v, sg[z]);
                        }
                    }
                }
                for (u32 j1 = lane + 64u; j1 < L1b; j1 += 32) {
                    const u32 p1 = __ldg(C.B.pos1 + rb * L1b + j1);
                    const u32 slot = self_off + (p1 & 0xffffu);
                    if (j1 == j1s) {
                        bval[slot] = diag_r;
                    } else {
                        const u32 sgn1 = __ldg(subB + j1).y & 0x80000000u;
                        const double *tb = C.B.terms + (size_t)(rb * L1b + j1) * nb;
                        double v = JA[p1 >> 16];
                        for (u32 q = 0; q < nb; ++q)
                            v += __ldg(tb + q);
                        bval[slot] = flip_sign(v, sgn1);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bar_full_arrive(group);
            } else {
                // ================= warp 7: alpha-alpha doubles (:339-358), alpha singles (:303-315), bulk stores
                const u64 Bdet = __ldg(P.dets + 2 * (P.row0 + r) + 1);
                const u32 j1s = __ldg(C.B.j1self + rb);
                // the bulk stores of the previous row must have read the buffer before it is overwritten
                if (lane == 0)
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                bar_free_sync(group);
                for (u32 d = lane; d < nDa; d += 32) {
                    const uint2 dp = d_pack[d];
                    bcol[dp.x] = (int)(dp.y + rb);
                    bval[dp.x] = d_val[d];
                }
                for (u32 g = lane; g < nSa; g += 32) {
                    const uint4 a = s_pack[g];
                    double v = s_pre[g];
                    for (u64 q = Bdet; q; q &= q - 1) {
                        const u32 kk = (u32)__ffsll((long long)q) - 1u;
                        v += SLICE ? slice[g * nsl + (C.packed ? kk * (kk + 1) / 2 + kk : kk * (u32)n1 + kk)]
                                   : __ldg(two_mo + a.z + (u32)n2 * kk + kk);
                    }
                    const u32 slot = a.x + j1s;
                    bcol[slot] = (int)(a.y + rb);
                    bval[slot] = flip_sign(v, a.w);
                }
                // ---- the finished row goes out as two bulk copies (TMA): 16-byte aligned bodies of the value and
                // column streams; the few entries before / after the aligned bodies by scalar stores
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bar_full_sync(group);
                const u32 hv = ov, nv = (M - hv) >> 1;                    // values: pairs
                const u32 hc = min(M, (4u - oc) & 3u), nc = (M - hc) >> 2; // columns: quads
                if (lane == 0) {
                    if (nv)
                        bulk_store_row(P.vals + out0 + hv, (u32)__cvta_generic_to_shared(bval + hv), nv * 16u, C.l2hint);
                    if (nc)
                        bulk_store_row(P.cols + out0 + hc, (u32)__cvta_generic_to_shared(bcol + hc), nc * 16u, C.l2hint);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                // the few entries before / after the aligned bodies (read before this warp frees the buffer again)
                if (lane >= 4 && lane < 4 + hv)
                    P.vals[out0 + lane - 4] = bval[lane - 4];
                if (lane >= 8 && lane - 8 + hv + 2 * nv < M)
                    P.vals[out0 + hv + 2 * nv + lane - 8] = bval[hv + 2 * nv + lane - 8];
                if (lane >= 12 && lane - 12 < hc)
                    P.cols[out0 + lane - 12] = bcol[lane - 12];
                if (lane >= 16 && lane < 20 && lane - 16 + hc + 4 * nc < M)
                    P.cols[out0 + hc + 4 * nc + lane - 16] = bcol[hc + 4 * nc + lane - 16];
                if (lane == 20)
                    P.lowcnt[r] = (int)(self_off + __ldg(C.B.selfj + rb)) + 1; // slots up to and including the diagonal
            }
        }
    }
    if ((threadIdx.x & 255u) == 224u)
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // all rows written before the CTA retires
}

} // namespace
