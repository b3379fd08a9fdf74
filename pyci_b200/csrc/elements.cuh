// Slater-Condon matrix elements of the excitation candidates, in the reference's operation order
// (/root/reference/pyci/src/sparseop.cpp:220-502): diagonal elements, the per-row single-excitation
// tables in shared memory, and candidate c -> (excited strings, signed element).  Shared by the
// construction kernels (build.cu) and the heat-bath / ENPT2 kernels (hci.cu); compile with --fmad=false.
#pragma once
#include "enumerate.cuh"

namespace {

// ---- matrix elements, in the reference's operation order ------------------------------------

// DOCI diagonal, sparseop.cpp:228-236,253: val1 + 2*val2
__device__ double diag_doci(const BuildParams &P, u64 det) {
    const int n = P.n;
    double val1 = 0.0, val2 = 0.0;
    for (u64 wi = det; wi; wi &= wi - 1) {
        const int k = __ffsll((long long)wi) - 1;
        val1 += __ldg(P.v + k * (n + 1));
        val2 += __ldg(P.h + k);
        for (u64 wj = wi & (wi - 1); wj; wj &= wj - 1)
            val2 += __ldg(P.w + k * n + (__ffsll((long long)wj) - 1));
    }
    return val1 + val2 * 2;
}

// FullCI diagonal, sparseop.cpp:283-292,367-372 (GenCI :443-449 is the alpha-only part); occupied
// orbitals are visited in ascending order like fill_occs does
__device__ double diag_twobody(const BuildParams &P, u64 da, u64 db) {
    const long n1 = P.n, n2 = n1 * n1, n3 = n2 * n1;
    double val2 = 0.0;
    for (u64 wi = da; wi; wi &= wi - 1) {
        const long ii = __ffsll((long long)wi) - 1, ioff = n3 * ii;
        val2 += __ldg(P.one_mo + (n1 + 1) * ii);
        for (u64 wk = wi & (wi - 1); wk; wk &= wk - 1) {
            const long kk = __ffsll((long long)wk) - 1, koff = ioff + n2 * kk;
            val2 += __ldg(P.two_mo + koff + n1 * ii + kk) - __ldg(P.two_mo + koff + n1 * kk + ii);
        }
        for (u64 wk = db; wk; wk &= wk - 1) {
            const long kk = __ffsll((long long)wk) - 1;
            val2 += __ldg(P.two_mo + ioff + n2 * kk + n1 * ii + kk);
        }
    }
    for (u64 wi = db; wi; wi &= wi - 1) {
        const long ii = __ffsll((long long)wi) - 1, ioff = n3 * ii;
        val2 += __ldg(P.one_mo + (n1 + 1) * ii);
        for (u64 wk = wi & (wi - 1); wk; wk &= wk - 1) {
            const long kk = __ffsll((long long)wk) - 1, koff = ioff + n2 * kk;
            val2 += __ldg(P.two_mo + koff + n1 * ii + kk) - __ldg(P.two_mo + koff + n1 * kk + ii);
        }
    }
    return val2;
}

// ---- per-row excitation tables ---------------------------------------------------------------------
// Single excitations of the row determinant, built once per row in shared memory (structure of arrays):
// excited string, signed matrix element (fill pass only), partial two_mo offset and packed (i, a, parity).
// Alpha-beta doubles are then one table entry per spin; same-spin doubles use the static pair table.
struct RowTables {
    u64 *sa_str, *sb_str;
    double *sa_val, *sb_val;
    u32 *sa_off, *sb_off, *sa_meta, *sb_meta;
};

__host__ __device__ inline size_t tables_bytes(u32 nSa, u32 nSb) { return (size_t)24 * ((nSa + 1) & ~1u) + (size_t)24 * ((nSb + 1) & ~1u); }
// the strings alone (all a pass that does not evaluate elements reads): they come first in the layout
__host__ __device__ inline size_t table_strings_bytes(u32 nSa, u32 nSb) { return (size_t)8 * ((nSa + 1) & ~1u) + (size_t)8 * ((nSb + 1) & ~1u); }

__device__ __forceinline__ RowTables carve_tables(unsigned char *base, u32 nSa, u32 nSb) {
    const u32 ea = (nSa + 1) & ~1u, eb = (nSb + 1) & ~1u;
    RowTables T;
    T.sa_str = reinterpret_cast<u64 *>(base);
    T.sb_str = T.sa_str + ea;
    T.sa_val = reinterpret_cast<double *>(T.sb_str + eb);
    T.sb_val = T.sa_val + ea;
    T.sa_off = reinterpret_cast<u32 *>(T.sb_val + eb);
    T.sb_off = T.sa_off + ea;
    T.sa_meta = T.sb_off + eb;
    T.sb_meta = T.sa_meta + ea;
    return T;
}

template<int KIND, bool VAL>
__device__ __forceinline__ void build_tables(const BuildParams &P, const RowShared &rs, const RowTables &T, u32 nSa,
                                             u32 nSb) {
    const long n1 = P.n, n2 = n1 * n1, n3 = n2 * n1;
    const int na = rs.nocc[0], nb = (KIND == PYCI_FULLCI) ? rs.nocc[1] : 0;
    const u32 nva = (u32)P.nvir_a, nvb = (u32)P.nvir_b;
    for (u32 t = threadIdx.x; t < nSa + nSb; t += blockDim.x) {
        if (t < nSa) {
            const u32 io = t / nva, ia = t - io * nva;
            const long i = rs.occ[0][io], a = rs.vir[0][ia];
            const int par = parity_single(rs.det[0], (int)i, (int)a);
            T.sa_str[t] = rs.det[0] ^ (1ULL << i) ^ (1ULL << a);
            if (VAL) { // VAL = false callers allocate the strings only (table_strings_bytes)
                T.sa_off[t] = (u32)(n3 * i + n1 * a);
                T.sa_meta[t] = (u32)i | ((u32)a << 8) | ((u32)par << 16);
            }
            if (VAL) { // sparseop.cpp:303-315 (GenCI :459-466)
                const long ioff = n3 * i;
                double val1 = __ldg(P.one_mo + n1 * i + a);
                for (int q = 0; q < na; ++q) {
                    const long kk = rs.occ[0][q], koff = ioff + n2 * kk;
                    val1 += __ldg(P.two_mo + koff + n1 * a + kk) - __ldg(P.two_mo + koff + n1 * kk + a);
                }
                for (int q = 0; q < nb; ++q) {
                    const long kk = rs.occ[1][q];
                    val1 += __ldg(P.two_mo + ioff + n2 * kk + n1 * a + kk);
                }
                T.sa_val[t] = apply_sign(val1, par);
            }
        } else {
            const u32 tb = t - nSa;
            const u32 io = tb / nvb, ia = tb - io * nvb;
            const long i = rs.occ[1][io], a = rs.vir[1][ia];
            const int par = parity_single(rs.det[1], (int)i, (int)a);
            T.sb_str[tb] = rs.det[1] ^ (1ULL << i) ^ (1ULL << a);
            if (VAL) {
                T.sb_off[tb] = (u32)(n2 * i + a);
                T.sb_meta[tb] = (u32)i | ((u32)a << 8) | ((u32)par << 16);
            }
            if (VAL) { // sparseop.cpp:382-394
                const long ioff = n3 * i;
                double val1 = __ldg(P.one_mo + n1 * i + a);
                for (int q = 0; q < na; ++q) {
                    const long kk = rs.occ[0][q];
                    val1 += __ldg(P.two_mo + ioff + n2 * kk + n1 * a + kk);
                }
                for (int q = 0; q < nb; ++q) {
                    const long kk = rs.occ[1][q], koff = ioff + n2 * kk;
                    val1 += __ldg(P.two_mo + koff + n1 * a + kk) - __ldg(P.two_mo + koff + n1 * kk + a);
                }
                T.sb_val[tb] = apply_sign(val1, par);
            }
        }
    }
}

// candidate c of the row -> excited strings (A, B) and, in the fill pass, the signed matrix element.
// Segment order: alpha-beta doubles | alpha-alpha doubles | beta-beta doubles | alpha singles | beta singles;
// consecutive candidates (the lanes of a warp) share a segment, so evaluation does not diverge.
template<int KIND, bool VAL>
__device__ __forceinline__ void candidate(const BuildParams &P, const RowShared &rs, const RowTables &T,
                                          const uchar2 *__restrict__ pairs, u32 c, u64 &A, u64 &B, double &val) {
    const long n1 = P.n, n2 = n1 * n1, n3 = n2 * n1;
    A = rs.det[0];
    B = rs.det[1];
    if (KIND == PYCI_DOCI) { // pair excitation k -> l, element v[k,l], no phase (sparseop.cpp:237-249)
        const u32 io = fdiv(c, P.dVa), ia = c - io * (u32)P.nvir_a;
        const int k = rs.occ[0][io], l = rs.vir[0][ia];
        A ^= (1ULL << k) | (1ULL << l);
        if (VAL)
            val = __ldg(P.v + k * n1 + l);
        return;
    }
    if (KIND == PYCI_FULLCI) {
        if (c < P.nAB) { // sparseop.cpp:318-337
            const u32 sa = fdiv(c, P.dSb), sb = c - sa * P.nSb;
            A = T.sa_str[sa];
            B = T.sb_str[sb];
            if (VAL) {
                const int par = ((T.sa_meta[sa] ^ T.sb_meta[sb]) >> 16) & 1;
                val = apply_sign(__ldg(P.two_mo + (T.sa_off[sa] + T.sb_off[sb])), par);
            }
            return;
        }
        c -= P.nAB;
    }
    if (c < P.nDa) { // sparseop.cpp:339-358 (GenCI :470-490)
        const u32 po = fdiv(c, P.dPva), pv = c - po * P.nPva;
        if (!VAL && rs.pm[0][0] != nullptr) { // strings only: per-row pair masks
            A ^= rs.pm[0][0][po] ^ rs.pm[0][1][pv];
            return;
        }
        const uchar2 o = pairs[po], v = pairs[pv];
        const long i = rs.occ[0][o.x], k = rs.occ[0][o.y], a = rs.vir[0][v.x], l = rs.vir[0][v.y];
        A ^= (1ULL << i) | (1ULL << k) | (1ULL << a) | (1ULL << l);
        if (VAL) {
            const long koff = n3 * i + n2 * k;
            const double x = __ldg(P.two_mo + koff + n1 * a + l) - __ldg(P.two_mo + koff + n1 * l + a);
            val = apply_sign(x, parity_double(rs.det[0], (int)i, (int)k, (int)a, (int)l));
        }
        return;
    }
    c -= P.nDa;
    if (KIND == PYCI_FULLCI) {
        if (c < P.nDb) { // sparseop.cpp:397-416
            const u32 po = fdiv(c, P.dPvb), pv = c - po * P.nPvb;
            if (!VAL && rs.pm[1][0] != nullptr) {
                B ^= rs.pm[1][0][po] ^ rs.pm[1][1][pv];
                return;
            }
            const uchar2 o = pairs[po], v = pairs[pv];
            const long i = rs.occ[1][o.x], k = rs.occ[1][o.y], a = rs.vir[1][v.x], l = rs.vir[1][v.y];
            B ^= (1ULL << i) | (1ULL << k) | (1ULL << a) | (1ULL << l);
            if (VAL) {
                const long koff = n3 * i + n2 * k;
                const double x = __ldg(P.two_mo + koff + n1 * a + l) - __ldg(P.two_mo + koff + n1 * l + a);
                val = apply_sign(x, parity_double(rs.det[1], (int)i, (int)k, (int)a, (int)l));
            }
            return;
        }
        c -= P.nDb;
    }
    if (c < P.nSa) {
        A = T.sa_str[c];
        if (VAL)
            val = T.sa_val[c];
        return;
    }
    c -= P.nSa;
    B = T.sb_str[c];
    if (VAL)
        val = T.sb_val[c];
}

// Signed matrix element of candidate c of the row WITHOUT the per-row single-excitation tables: for rows whose few
// stored entries are already known (recorded hits of a selected space), where building the tables -- every single
// excitation of the row, nocc two-electron terms each -- would cost far more than the entries themselves.  Same
// operations in the same order as build_tables / candidate (sparseop.cpp:303-315,318-337,339-358,382-416,459-490).
template<int KIND>
__device__ __forceinline__ double hit_element(const BuildParams &P, const RowShared &rs, const uchar2 *__restrict__ pairs,
                                              u32 c) {
    const long n1 = P.n, n2 = n1 * n1, n3 = n2 * n1;
    u64 A, B;
    u32 code;
    decode<KIND>(P, rs, pairs, c, A, B, code);
    const int type = code >> 24;
    const long i = (code >> 18) & 63, a = (code >> 12) & 63, k = (code >> 6) & 63, l = code & 63;
    const int na = rs.nocc[0], nb = (KIND == PYCI_FULLCI) ? rs.nocc[1] : 0;
    switch (type) {
    case T_PAIR:
        return __ldg(P.v + i * n1 + a);
    case T_AB: {
        const int par = parity_single(rs.det[0], (int)i, (int)a) ^ parity_single(rs.det[1], (int)k, (int)l);
        return apply_sign(__ldg(P.two_mo + ((u32)(n3 * i + n1 * a) + (u32)(n2 * k + l))), par);
    }
    case T_AA:
    case T_BB: {
        const long koff = n3 * i + n2 * k;
        const double x = __ldg(P.two_mo + koff + n1 * a + l) - __ldg(P.two_mo + koff + n1 * l + a);
        return apply_sign(x, parity_double(rs.det[type == T_AA ? 0 : 1], (int)i, (int)k, (int)a, (int)l));
    }
    case T_SA: {
        const long ioff = n3 * i;
        double val1 = __ldg(P.one_mo + n1 * i + a);
        for (int q = 0; q < na; ++q) {
            const long kk = rs.occ[0][q], koff = ioff + n2 * kk;
            val1 += __ldg(P.two_mo + koff + n1 * a + kk) - __ldg(P.two_mo + koff + n1 * kk + a);
        }
        for (int q = 0; q < nb; ++q) {
            const long kk = rs.occ[1][q];
            val1 += __ldg(P.two_mo + ioff + n2 * kk + n1 * a + kk);
        }
        return apply_sign(val1, parity_single(rs.det[0], (int)i, (int)a));
    }
    default: { // T_SB
        const long ioff = n3 * i;
        double val1 = __ldg(P.one_mo + n1 * i + a);
        for (int q = 0; q < na; ++q) {
            const long kk = rs.occ[0][q];
            val1 += __ldg(P.two_mo + ioff + n2 * kk + n1 * a + kk);
        }
        for (int q = 0; q < nb; ++q) {
            const long kk = rs.occ[1][q], koff = ioff + n2 * kk;
            val1 += __ldg(P.two_mo + koff + n1 * a + kk) - __ldg(P.two_mo + koff + n1 * kk + a);
        }
        return apply_sign(val1, parity_single(rs.det[1], (int)i, (int)a));
    }
    }
}

} // namespace
