// Excitation enumerator shared by the construction kernels (build.cu) and the RDM kernel (rdm.cu):
// the device restatement of the occ x vir loop nests of SparseOp::add_row
// (/root/reference/pyci/src/sparseop.cpp:220-502) and compute_rdms (rdm.cpp:20-632), flattened to
// one candidate index per excitation so that the threads of a CTA can split them evenly.
#pragma once
#include "common.cuh"

namespace {

enum { T_AB = 0, T_AA = 1, T_BB = 2, T_SA = 3, T_SB = 4, T_DIAG = 5, T_PAIR = 6 };

// division of a 31-bit candidate index by a launch-time constant: q = (umulhi(c, mul) + c) >> sh
struct FastDiv {
    u32 d, mul, sh;
};
inline FastDiv make_fastdiv(u32 d) {
    FastDiv f;
    f.d = d ? d : 1;
    u32 s = 0;
    while ((1ULL << s) < f.d)
        ++s;
    f.sh = s;
    f.mul = (u32)((((1ULL << 32) * ((1ULL << s) - f.d)) / f.d) + 1ULL);
    return f;
}

struct BuildParams {
    const u64 *dets;
    int nwords;
    int n;
    int nocc_a, nocc_b, nvir_a, nvir_b;
    u32 nSa, nSb, nDa, nDb, nAB, nPva, nPvb, ncand;
    long row0, nloc, ncol;
    const double *one_mo, *two_mo, *h, *v, *w;
    int kl_sym; // two_mo[i,k,a,l] == two_mo[i,l,a,k] for all indices (pyci_ham::kl_sym)
    long *indptr;
    int *cols;
    double *vals;
    int *lowcnt;
    double *diag;
    int *rowcnt;
    int maxrow; // capacity of the shared row buffer (entries)
    int npairs_dim;
    const double *coeffs; // RDMs only
    double *rdm1, *rdm2;  // RDMs only
    FastDiv dSb, dPva, dPvb, dVa, dVb; // divisors of decode: nSb, nPva, nPvb, nvir_a, nvir_b
    int sort_passes, sort_dbits;       // LSD radix sort of a row: passes x digit bits cover the column index
};

#ifdef __CUDACC__
__device__ __forceinline__ u32 fdiv(u32 c, const FastDiv &f) { return (__umulhi(c, f.mul) + c) >> f.sh; }
#endif

__device__ __forceinline__ u32 pack_code(int type, int i, int a, int k, int l) {
    return ((u32)type << 24) | ((u32)i << 18) | ((u32)a << 12) | ((u32)k << 6) | (u32)l;
}

struct RowShared {
    u64 det[2];
    int count;
    int nocc[2];
    unsigned char occ[2][64];
    unsigned char vir[2][64];
    // per-row masks of the same-spin double excitations (dynamic shared memory, or null): pm[spin][0][po] = bits of
    // the occupied pair po, pm[spin][1][pv] = bits of the virtual pair pv; a double excitation is then two loads
    // and two XORs instead of two pair look-ups, four orbital look-ups and four shifts
    u64 *pm[2][2];
};

// bytes of the pair-mask tables of one CTA
inline size_t pair_mask_bytes(const BuildParams &P, int kind) {
    auto c2 = [](long m) { return (size_t)(m * (m - 1) / 2); };
    const long va = P.n - P.nocc_a, vb = P.n - P.nocc_b;
    size_t e = c2(P.nocc_a) + c2(va);
    if (kind == PYCI_FULLCI)
        e += c2(P.nocc_b) + c2(vb);
    return 8 * (e + 4);
}

// carve the tables (once per CTA; thread 0) ...
__device__ __forceinline__ void pair_masks_carve(RowShared &rs, const BuildParams &P, unsigned char *base, int nspin) {
    if (threadIdx.x == 0) {
        u64 *q = reinterpret_cast<u64 *>(base);
        for (int s = 0; s < 2; ++s) {
            const int no = s ? P.nocc_b : P.nocc_a, nv = P.n - no;
            rs.pm[s][0] = rs.pm[s][1] = nullptr;
            if (s < nspin) {
                rs.pm[s][0] = q;
                q += no * (no - 1) / 2;
                rs.pm[s][1] = q;
                q += nv * (nv - 1) / 2;
            }
        }
    }
}
__device__ __forceinline__ void pair_masks_none(RowShared &rs) {
    if (threadIdx.x == 0)
        rs.pm[0][0] = rs.pm[0][1] = rs.pm[1][0] = rs.pm[1][1] = nullptr;
}

// ... and fill them for the current row (after row_setup and a barrier; followed by a barrier)
__device__ __forceinline__ void pair_masks_build(const RowShared &rs, const BuildParams &P, const uchar2 *__restrict__ pairs,
                                                 int nspin) {
    for (int s = 0; s < nspin; ++s) {
        const int no = s ? P.nocc_b : P.nocc_a, nv = P.n - no;
        const int npo = no * (no - 1) / 2, npv = nv * (nv - 1) / 2;
        u64 *mo = rs.pm[s][0], *mv = rs.pm[s][1];
        for (int t = threadIdx.x; t < npo + npv; t += blockDim.x) {
            if (t < npo) {
                const uchar2 o = pairs[t];
                mo[t] = (1ULL << rs.occ[s][o.x]) | (1ULL << rs.occ[s][o.y]);
            } else {
                const uchar2 v = pairs[t - npo];
                mv[t - npo] = (1ULL << rs.vir[s][v.x]) | (1ULL << rs.vir[s][v.y]);
            }
        }
    }
}

// fill_occs / fill_virs (common.cpp:85-113) for one or two 64-bit strings, by warp 0
__device__ __forceinline__ void row_setup(RowShared &rs, const BuildParams &P, long row, int nspin) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const u32 lt = (1u << lane) - 1u;
        for (int s = 0; s < nspin; ++s) {
            const u64 w = P.dets[row * P.nwords + s];
            int bo = 0, bv = 0;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int pos = lane + 32 * half;
                const bool inb = pos < P.n;
                const bool occ = inb && ((w >> pos) & 1ULL);
                const bool vir = inb && !occ;
                const u32 mo = __ballot_sync(0xffffffffu, occ);
                const u32 mv = __ballot_sync(0xffffffffu, vir);
                if (occ)
                    rs.occ[s][bo + __popc(mo & lt)] = (unsigned char)pos;
                if (vir)
                    rs.vir[s][bv + __popc(mv & lt)] = (unsigned char)pos;
                bo += __popc(mo);
                bv += __popc(mv);
            }
            if (lane == 0) {
                rs.det[s] = w;
                rs.nocc[s] = bo;
            }
        }
        if (lane == 0) {
            if (nspin == 1)
                rs.det[1] = 0ULL;
            rs.count = 0;
        }
    }
}

// candidate c -> excited strings (A', B') and excitation code.  Segment order:
// alpha-beta doubles | alpha-alpha doubles | beta-beta doubles | alpha singles | beta singles.
// One-spin kinds only have the "alpha" segments.  Pairs (x<y) are enumerated as p = y(y-1)/2 + x.
template<int KIND>
__device__ __forceinline__ void decode(const BuildParams &P, const RowShared &rs,
                                       const uchar2 *__restrict__ pairs, u32 c, u64 &A, u64 &B, u32 &code) {
    A = rs.det[0];
    B = rs.det[1];
    if (KIND == PYCI_DOCI) {
        const u32 io = c / (u32)P.nvir_a, ia = c - io * (u32)P.nvir_a;
        const int k = rs.occ[0][io], l = rs.vir[0][ia];
        A ^= (1ULL << k) | (1ULL << l);
        code = pack_code(T_PAIR, k, l, 0, 0);
        return;
    }
    if (KIND == PYCI_FULLCI) {
        if (c < P.nAB) {
            const u32 sa = c / P.nSb, sb = c - sa * P.nSb;
            const u32 io = sa / (u32)P.nvir_a, ia = sa - io * (u32)P.nvir_a;
            const u32 ko = sb / (u32)P.nvir_b, la = sb - ko * (u32)P.nvir_b;
            const int i = rs.occ[0][io], a = rs.vir[0][ia], k = rs.occ[1][ko], l = rs.vir[1][la];
            A ^= (1ULL << i) | (1ULL << a);
            B ^= (1ULL << k) | (1ULL << l);
            code = pack_code(T_AB, i, a, k, l);
            return;
        }
        c -= P.nAB;
    }
    if (c < P.nDa) {
        const u32 po = c / P.nPva, pv = c - po * P.nPva;
        const uchar2 o = pairs[po], v = pairs[pv];
        const int i = rs.occ[0][o.x], k = rs.occ[0][o.y], a = rs.vir[0][v.x], l = rs.vir[0][v.y];
        A ^= (1ULL << i) | (1ULL << k) | (1ULL << a) | (1ULL << l);
        code = pack_code(T_AA, i, a, k, l);
        return;
    }
    c -= P.nDa;
    if (KIND == PYCI_FULLCI) {
        if (c < P.nDb) {
            const u32 po = c / P.nPvb, pv = c - po * P.nPvb;
            const uchar2 o = pairs[po], v = pairs[pv];
            const int i = rs.occ[1][o.x], k = rs.occ[1][o.y], a = rs.vir[1][v.x], l = rs.vir[1][v.y];
            B ^= (1ULL << i) | (1ULL << k) | (1ULL << a) | (1ULL << l);
            code = pack_code(T_BB, i, a, k, l);
            return;
        }
        c -= P.nDb;
    }
    if (c < P.nSa) {
        const u32 io = c / (u32)P.nvir_a, ia = c - io * (u32)P.nvir_a;
        const int i = rs.occ[0][io], a = rs.vir[0][ia];
        A ^= (1ULL << i) | (1ULL << a);
        code = pack_code(T_SA, i, a, 0, 0);
        return;
    }
    c -= P.nSa;
    {
        const u32 io = c / (u32)P.nvir_b, ia = c - io * (u32)P.nvir_b;
        const int i = rs.occ[1][io], a = rs.vir[1][ia];
        B ^= (1ULL << i) | (1ULL << a);
        code = pack_code(T_SB, i, a, 0, 0);
    }
}


// strings only (no excitation code): same-spin doubles through the per-row pair masks when they are there
template<int KIND>
__device__ __forceinline__ void decode_dets(const BuildParams &P, const RowShared &rs, const uchar2 *__restrict__ pairs,
                                            u32 c, u64 &A, u64 &B) {
    if (KIND != PYCI_DOCI && rs.pm[0][0] != nullptr) {
        u32 d = c;
        bool ok = true;
        if (KIND == PYCI_FULLCI) {
            ok = d >= P.nAB;
            d -= P.nAB;
        }
        if (ok && d < P.nDa) {
            const u32 po = fdiv(d, P.dPva), pv = d - po * P.nPva;
            A = rs.det[0] ^ rs.pm[0][0][po] ^ rs.pm[0][1][pv];
            B = rs.det[1];
            return;
        }
        if (KIND == PYCI_FULLCI && ok && d - P.nDa < P.nDb) {
            d -= P.nDa;
            const u32 po = fdiv(d, P.dPvb), pv = d - po * P.nPvb;
            A = rs.det[0];
            B = rs.det[1] ^ rs.pm[1][0][po] ^ rs.pm[1][1][pv];
            return;
        }
    }
    u32 code;
    decode<KIND>(P, rs, pairs, c, A, B, code);
}

__device__ __forceinline__ void fill_pairs(uchar2 *pairs, int m) {
    // pairs[y(y-1)/2 + x] = (x, y) for x < y < m
    for (int y = 1 + threadIdx.x; y < m; y += blockDim.x)
        for (int x = 0; x < y; ++x)
            pairs[y * (y - 1) / 2 + x] = make_uchar2((unsigned char)x, (unsigned char)y);
}

template<int KM>
struct SlotOf;
template<>
struct SlotOf<KEY32> {
    typedef Slot32 type;
};
template<>
struct SlotOf<KEY64> {
    typedef Slot64 type;
};
template<>
struct SlotOf<KEY128> {
    typedef Slot128 type;
};

template<int KM>
DetIndex<KM> make_index(const pyci_wfn *wfn) {
    DetIndex<KM> ix;
    ix.slots = reinterpret_cast<const typename SlotOf<KM>::type *>(wfn->slots);
    ix.mask = wfn->mask;
    ix.shift = (wfn->kind == PYCI_FULLCI) ? (int)wfn->nbasis : 0;
    ix.bloom = wfn->bloom;
    ix.bmask = wfn->bmask;
    return ix;
}

// candidate-segment sizes for a wave function (host)
inline int enum_params_init(BuildParams &P, const pyci_wfn *wfn) {
    memset(&P, 0, sizeof(P));
    const int n = (int)wfn->nbasis;
    const int kind = wfn->kind;
    P.dets = wfn->dets;
    P.nwords = wfn->nwords;
    P.n = n;
    P.nocc_a = (int)wfn->nocc_up;
    P.nvir_a = n - P.nocc_a;
    P.nocc_b = (kind == PYCI_FULLCI) ? (int)wfn->nocc_dn : 0;
    P.nvir_b = (kind == PYCI_FULLCI) ? n - P.nocc_b : 0;
    auto c2 = [](long m) { return (u32)(m * (m - 1) / 2); };
    P.nSa = (u32)(P.nocc_a * P.nvir_a);
    if (kind == PYCI_DOCI) {
        P.ncand = P.nSa;
    } else {
        P.nPva = c2(P.nvir_a);
        P.nDa = c2(P.nocc_a) * P.nPva;
        if (kind == PYCI_FULLCI) {
            P.nSb = (u32)(P.nocc_b * P.nvir_b);
            P.nPvb = c2(P.nvir_b);
            P.nDb = c2(P.nocc_b) * P.nPvb;
            P.nAB = P.nSa * P.nSb;
        }
        const double tot = (double)P.nSa + P.nSb + (double)P.nDa + P.nDb + (double)P.nSa * P.nSb;
        if (tot > 2.0e9)
            PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "excitation space per determinant too large (%g)", tot);
        P.ncand = P.nSa + P.nSb + P.nDa + P.nDb + P.nAB;
    }
    P.npairs_dim = std::max(std::max(P.nocc_a, P.nvir_a), std::max(P.nocc_b, P.nvir_b)) + 1;
    // avoid division by zero in decode when a segment is empty
    if (P.nPva == 0) P.nPva = 1;
    if (P.nPvb == 0) P.nPvb = 1;
    if (P.nSb == 0) { P.nSb = 1; P.nAB = 0; }
    if (P.nvir_a == 0) P.nvir_a = 1;
    if (P.nvir_b == 0) P.nvir_b = 1;
    P.dSb = make_fastdiv(P.nSb);
    P.dPva = make_fastdiv(P.nPva);
    P.dPvb = make_fastdiv(P.nPvb);
    P.dVa = make_fastdiv((u32)P.nvir_a);
    P.dVb = make_fastdiv((u32)P.nvir_b);
    return PYCI_OK;
}

inline size_t pair_table_bytes(const BuildParams &P) {
    return sizeof(uchar2) * (size_t)(P.npairs_dim * (P.npairs_dim - 1) / 2 + 1);
}

} // namespace
