// Generic multi-word construction: determinants of more than 64 orbitals (nword = ceil(nbasis / 64) > 1,
// /root/reference/pyci/src/common.cpp:280-282; every add_row of sparseop.cpp:220-502 walks nword words).
//
// This is the SLOW path -- none of the five benchmark configurations needs it, the reference's own tests only build
// wave functions of 65 and 129 orbitals (pyci/test/test_wavefunction.py:45) -- kept simple on purpose: one warp per
// row, the lanes split the row's single excitations and each lane enumerates the double excitations nested under its
// singles exactly as the reference's loop nest does, probing an open-addressing table of determinant indices (keys are
// compared word by word in the determinant array).  Two passes (count, fill) around the shared int64 scan; rows are
// written unsorted into scratch and ordered by a rank sort.  The operator that comes out is the ordinary pyci_op:
// SpMV, solve, export and get_element are the single-word code.  RDMs, add_hci, ENPT2, index_dets and update of such
// wave functions stay PYCI_ERR_UNSUPPORTED.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

constexpr int MW_MAXW = 8;    // words per determinant: nbasis <= 256 (4 words per string, two strings)
constexpr int MW_MAXORB = 256;
constexpr int MW_WARPS = 4;

struct MWParams {
    const u64 *dets;
    int nw, W; // words per string, words per determinant
    int kind, n, na, nb;
    long ndet, row0, nloc, ncol;
    const double *one_mo, *two_mo, *h, *v, *w;
    const int *slots;
    u32 mask;
    const long *indptr;
    int *rowcnt;   // count pass
    int *tcols;    // fill pass: unsorted rows
    double *tvals;
    double *diag;
};

__device__ __forceinline__ u32 mw_hash(const u64 *d, int W) {
    u64 h = 0x9e3779b97f4a7c15ULL;
    for (int q = 0; q < W; ++q) {
        h ^= d[q];
        h ^= h >> 33;
        h *= 0xff51afd7ed558ccdULL;
        h ^= h >> 33;
    }
    return (u32)(h ^ (h >> 29));
}

__device__ __forceinline__ bool mw_equal(const u64 *a, const u64 *b, int W) {
    for (int q = 0; q < W; ++q)
        if (a[q] != b[q])
            return false;
    return true;
}

// Wfn::index_det (onespinwfn.cpp:123-126, twospinwfn.cpp:129-132) for multi-word strings
__device__ __forceinline__ int mw_find(const MWParams &P, const u64 *d) {
    u32 p = mw_hash(d, P.W) & P.mask;
    for (;;) {
        const int s = P.slots[p];
        if (s < 0)
            return -1;
        if (mw_equal(P.dets + (size_t)s * P.W, d, P.W))
            return s;
        p = (p + 1) & P.mask;
    }
}

__global__ void mw_insert_kernel(const u64 *dets, int W, long ndet, int *slots, u32 mask, u64 valid_last, int nw, int na,
                                 int nb, int nstr, int *bad) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ndet)
        return;
    const u64 *d = dets + (size_t)i * W;
    // occupations (Wfn::init / add_det preconditions): electrons per string, nothing beyond nbasis
    for (int s = 0; s < nstr; ++s) {
        int pc = 0;
        for (int q = 0; q < nw; ++q)
            pc += __popcll(d[s * nw + q]);
        if (pc != (s ? nb : na) || (d[s * nw + nw - 1] & ~valid_last))
            atomicMin(bad + 1, (int)i);
    }
    u32 p = mw_hash(d, W) & mask;
    for (;;) {
        const int cur = atomicCAS(slots + p, -1, (int)i);
        if (cur == -1)
            return;
        if (mw_equal(dets + (size_t)cur * W, d, W)) { // duplicate determinant
            atomicAdd(bad, 1);
            return;
        }
        p = (p + 1) & mask;
    }
}

__device__ __forceinline__ void mw_flip(u64 *s, int o) { s[o >> 6] ^= 1ULL << (o & 63); }

// phase_single_det (common.cpp:163-194): occupied orbitals strictly between i and a, over the words
__device__ __forceinline__ int mw_parity(const u64 *s, int i, int a) {
    const int lo = min(i, a) + 1, hi = max(i, a); // bits [lo, hi)
    int pc = 0;
    for (int q = lo >> 6; q <= (hi - 1) >> 6 && lo < hi; ++q) {
        u64 m = ~0ULL;
        if (q == (lo >> 6))
            m &= ~0ULL << (lo & 63);
        if (q == ((hi - 1) >> 6) && (hi & 63))
            m &= (1ULL << (hi & 63)) - 1ULL;
        pc += __popcll(s[q] & m);
    }
    return pc & 1;
}

// phase_double_det (common.cpp:196-261)
__device__ __forceinline__ int mw_parity2(const u64 *s, int i1, int i2, int a1, int a2) {
    return (mw_parity(s, i1, a1) + mw_parity(s, i2, a2) + ((i2 < a1) || (i1 > a2))) & 1;
}

struct MWRow { // per warp, shared memory
    unsigned short occ[2][MW_MAXORB], vir[2][MW_MAXORB];
    int cnt;
};

// ordered occupied / virtual lists of one string (fill_occs / fill_virs, common.cpp:85-113)
__device__ void mw_lists(const u64 *s, int n, unsigned short *occ, unsigned short *vir, int lane) {
    const u32 lt = (1u << lane) - 1u;
    int bo = 0, bv = 0;
    for (int base = 0; base < n; base += 32) {
        const int o = base + lane;
        const bool in = o < n, on = in && ((s[o >> 6] >> (o & 63)) & 1ULL);
        const u32 mo = __ballot_sync(0xffffffffu, on), mv = __ballot_sync(0xffffffffu, in && !on);
        if (on)
            occ[bo + __popc(mo & lt)] = (unsigned short)o;
        if (in && !on)
            vir[bv + __popc(mv & lt)] = (unsigned short)o;
        bo += __popc(mo);
        bv += __popc(mv);
    }
}

// one found entry: counted, or written (unsorted) into the row's scratch segment
template<bool FILL>
__device__ __forceinline__ void mw_emit(const MWParams &P, MWRow &R, long base, int &count, int col, double val) {
    if (FILL) {
        const int at = atomicAdd(&R.cnt, 1);
        P.tcols[base + at] = col;
        P.tvals[base + at] = val;
    } else {
        ++count;
    }
}

template<bool FILL>
__global__ void __launch_bounds__(32 * MW_WARPS) mw_rows_kernel(MWParams P) {
    __shared__ MWRow rows[MW_WARPS];
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    MWRow &R = rows[wq];
    const long n1 = P.n, n2 = n1 * n1, n3 = n2 * n1;
    const int nw = P.nw, W = P.W, na = P.na, nb = P.nb, va = P.n - na, vb = P.n - nb;
    for (long r = (long)blockIdx.x * MW_WARPS + wq; r < P.nloc; r += (long)gridDim.x * MW_WARPS) {
        const long row = P.row0 + r;
        const u64 *rd = P.dets + (size_t)row * W;
        __syncwarp();
        mw_lists(rd, P.n, R.occ[0], R.vir[0], lane);
        if (P.kind == PYCI_FULLCI)
            mw_lists(rd + nw, P.n, R.occ[1], R.vir[1], lane);
        if (lane == 0)
            R.cnt = 0;
        __syncwarp();
        const long base = FILL ? P.indptr[r] : 0;
        int count = 0;
        u64 d[MW_MAXW];
        for (int q = 0; q < W; ++q)
            d[q] = rd[q];
        if (P.kind == PYCI_DOCI) {
            // pair excitations k -> l, element v[k,l] (sparseop.cpp:237-249)
            for (int u = lane; u < na * va; u += 32) {
                const int k = R.occ[0][u / va], l = R.vir[0][u % va];
                mw_flip(d, k);
                mw_flip(d, l);
                const int jd = mw_find(P, d);
                if (jd >= 0 && jd < P.ncol)
                    mw_emit<FILL>(P, R, base, count, jd, P.v[k * n1 + l]);
                mw_flip(d, k);
                mw_flip(d, l);
            }
        } else {
            const int nspin = (P.kind == PYCI_FULLCI) ? 2 : 1;
            for (int sp = 0; sp < nspin; ++sp) {
                // spin sp singles with everything nested under them: sparseop.cpp:294-358 (alpha, with the alpha-beta
                // doubles), :374-416 (beta); GenCI :451-490 is the alpha part without a beta string
                const int no = sp ? nb : na, nv = sp ? vb : va;
                u64 *ds = d + sp * nw;
                const u64 *rs = rd + sp * nw;
                for (int u = lane; u < no * nv; u += 32) {
                    const int io = u / nv, ja = u % nv;
                    const long ii = R.occ[sp][io], jj = R.vir[sp][ja], ioff = n3 * ii;
                    mw_flip(ds, (int)ii);
                    mw_flip(ds, (int)jj);
                    const int sign1 = mw_parity(rs, (int)ii, (int)jj);
                    int jd = mw_find(P, d);
                    if (jd >= 0 && jd < P.ncol) {
                        double val1 = P.one_mo[n1 * ii + jj];
                        if (sp == 0) { // :303-312 (GenCI :459-466)
                            for (int k = 0; k < na; ++k) {
                                const long kk = R.occ[0][k], koff = ioff + n2 * kk;
                                val1 += P.two_mo[koff + n1 * jj + kk] - P.two_mo[koff + n1 * kk + jj];
                            }
                            if (nspin == 2)
                                for (int k = 0; k < nb; ++k) {
                                    const long kk = R.occ[1][k];
                                    val1 += P.two_mo[ioff + n2 * kk + n1 * jj + kk];
                                }
                        } else { // :382-394
                            for (int k = 0; k < na; ++k) {
                                const long kk = R.occ[0][k];
                                val1 += P.two_mo[ioff + n2 * kk + n1 * jj + kk];
                            }
                            for (int k = 0; k < nb; ++k) {
                                const long kk = R.occ[1][k], koff = ioff + n2 * kk;
                                val1 += P.two_mo[koff + n1 * jj + kk] - P.two_mo[koff + n1 * kk + jj];
                            }
                        }
                        mw_emit<FILL>(P, R, base, count, jd, apply_sign(val1, sign1));
                    }
                    if (sp == 0 && nspin == 2) { // alpha-beta doubles, :318-337
                        u64 *db = d + nw;
                        for (int k = 0; k < nb; ++k) {
                            const long kk = R.occ[1][k], koff = ioff + n2 * kk;
                            for (int l = 0; l < vb; ++l) {
                                const long ll = R.vir[1][l];
                                mw_flip(db, (int)kk);
                                mw_flip(db, (int)ll);
                                jd = mw_find(P, d);
                                if (jd >= 0 && jd < P.ncol)
                                    mw_emit<FILL>(P, R, base, count, jd,
                                                  apply_sign(P.two_mo[koff + n1 * jj + ll],
                                                             sign1 ^ mw_parity(rd + nw, (int)kk, (int)ll)));
                                mw_flip(db, (int)kk);
                                mw_flip(db, (int)ll);
                            }
                        }
                    }
                    for (int k = io + 1; k < no; ++k) { // same-spin doubles, :339-358 / :397-416 (GenCI :470-490)
                        const long kk = R.occ[sp][k], koff = ioff + n2 * kk;
                        for (int l = ja + 1; l < nv; ++l) {
                            const long ll = R.vir[sp][l];
                            mw_flip(ds, (int)kk);
                            mw_flip(ds, (int)ll);
                            jd = mw_find(P, d);
                            if (jd >= 0 && jd < P.ncol) {
                                const double x = P.two_mo[koff + n1 * jj + ll] - P.two_mo[koff + n1 * ll + jj];
                                mw_emit<FILL>(P, R, base, count, jd,
                                              apply_sign(x, mw_parity2(rs, (int)ii, (int)kk, (int)jj, (int)ll)));
                            }
                            mw_flip(ds, (int)kk);
                            mw_flip(ds, (int)ll);
                        }
                    }
                    mw_flip(ds, (int)ii);
                    mw_flip(ds, (int)jj);
                }
            }
        }
        // ---- diagonal (sparseop.cpp:228-236,253 DOCI; :283-292,367-372,421-424 FullCI; :443-449,496-499 GenCI),
        // by one lane in the reference's summation order
        if (lane == 0) {
            double dg;
            if (P.kind == PYCI_DOCI) {
                double val1 = 0.0, val2 = 0.0;
                for (int i = 0; i < na; ++i) {
                    const long k = R.occ[0][i];
                    val1 += P.v[k * (n1 + 1)];
                    val2 += P.h[k];
                    for (int j = i + 1; j < na; ++j)
                        val2 += P.w[k * n1 + R.occ[0][j]];
                }
                dg = val1 + val2 * 2;
            } else {
                double val2 = 0.0;
                for (int i = 0; i < na; ++i) {
                    const long ii = R.occ[0][i], ioff = n3 * ii;
                    val2 += P.one_mo[(n1 + 1) * ii];
                    for (int k = i + 1; k < na; ++k) {
                        const long kk = R.occ[0][k], koff = ioff + n2 * kk;
                        val2 += P.two_mo[koff + n1 * ii + kk] - P.two_mo[koff + n1 * kk + ii];
                    }
                    if (P.kind == PYCI_FULLCI)
                        for (int k = 0; k < nb; ++k) {
                            const long kk = R.occ[1][k];
                            val2 += P.two_mo[ioff + n2 * kk + n1 * ii + kk];
                        }
                }
                if (P.kind == PYCI_FULLCI)
                    for (int i = 0; i < nb; ++i) {
                        const long ii = R.occ[1][i], ioff = n3 * ii;
                        val2 += P.one_mo[(n1 + 1) * ii];
                        for (int k = i + 1; k < nb; ++k) {
                            const long kk = R.occ[1][k], koff = ioff + n2 * kk;
                            val2 += P.two_mo[koff + n1 * ii + kk] - P.two_mo[koff + n1 * kk + ii];
                        }
                    }
                dg = val2;
            }
            if (FILL) {
                P.diag[r] = dg;
                if (row < P.ncol)
                    mw_emit<true>(P, R, base, count, (int)row, dg);
            } else if (row < P.ncol) {
                ++count;
            }
        }
        if (!FILL) {
            for (int o = 16; o > 0; o >>= 1)
                count += __shfl_xor_sync(0xffffffffu, count, o);
            if (lane == 0)
                P.rowcnt[r] = count;
        }
    }
}

// sort_row (sparseop.cpp:214-218) by ranks: entry e goes to the number of entries with a smaller column (columns of a
// row are distinct); one warp per row
__global__ void __launch_bounds__(128) mw_sort_kernel(const long *indptr, const int *tcols, const double *tvals, long row0,
                                                      long nloc, int *cols, double *vals, int *lowcnt) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long r = warp; r < nloc; r += nwarps) {
        const long b = indptr[r], m = indptr[r + 1] - b;
        int low = 0;
        for (long e = lane; e < m; e += 32) {
            const int c = tcols[b + e];
            long rank = 0;
            for (long f = 0; f < m; ++f)
                rank += tcols[b + f] < c;
            cols[b + rank] = c;
            vals[b + rank] = tvals[b + e];
            low += ((long)c <= row0 + r);
        }
        for (int o = 16; o > 0; o >>= 1)
            low += __shfl_xor_sync(0xffffffffu, low, o);
        if (lane == 0)
            lowcnt[r] = low;
    }
}

} // namespace

int mw_index_build(pyci_wfn *wfn) {
    pyci_ctx *ctx = wfn->ctx;
    cudaStream_t st = ctx->stream;
    u64 cap = 16;
    while (cap < 2 * (u64)wfn->ndet)
        cap <<= 1;
    if (cap > (1ULL << 31))
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "too many determinants for the device index (%ld)", wfn->ndet);
    dev_free(wfn->slots);
    wfn->slots = nullptr;
    wfn->index_valid = false;
    wfn->mask = (u32)(cap - 1);
    PYCI_CUDA(dev_malloc(&wfn->slots, sizeof(int) * (size_t)cap));
    PYCI_CUDA(cudaMemsetAsync(wfn->slots, 0xFF, sizeof(int) * (size_t)cap, st));
    if (wfn->ndet > 0) {
        int *bad = nullptr;
        const int init[2] = {0, 0x7fffffff};
        int h[2] = {0, 0x7fffffff};
        PYCI_CUDA(dev_malloc(&bad, 2 * sizeof(int)));
        PYCI_CUDA(cudaMemcpyAsync(bad, init, 2 * sizeof(int), cudaMemcpyHostToDevice, st));
        const int nw = (int)((wfn->nbasis + 63) / 64), nstr = wfn->kind == PYCI_FULLCI ? 2 : 1;
        const u64 valid_last = (wfn->nbasis & 63) ? ((1ULL << (wfn->nbasis & 63)) - 1ULL) : ~0ULL;
        mw_insert_kernel<<<(unsigned)((wfn->ndet + 255) / 256), 256, 0, st>>>(
            wfn->dets, nw * nstr, wfn->ndet, reinterpret_cast<int *>(wfn->slots), wfn->mask, valid_last, nw,
            (int)wfn->nocc_up, (int)wfn->nocc_dn, nstr, bad);
        ctx->launches++;
        PYCI_CUDA(cudaMemcpyAsync(h, bad, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        PYCI_CUDA(cudaStreamSynchronize(st));
        dev_free(bad);
        if (h[1] != 0x7fffffff)
            PYCI_FAIL(PYCI_ERR_VALUE, "determinant %d does not have the declared occupation", h[1]);
        if (h[0])
            PYCI_FAIL(PYCI_ERR_VALUE, "wave function contains %d duplicate determinant(s)", h[0]);
    }
    wfn->index_valid = true;
    return PYCI_OK;
}

int mw_op_build(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, pyci_op *op) {
    cudaStream_t st = ctx->stream;
    PYCI_NVTX("pyci:build(multi-word slow path)");
    if (wfn->nbasis > MW_MAXORB)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "nbasis = %ld: the multi-word device path handles nbasis <= %d", wfn->nbasis, MW_MAXORB);
    MWParams P;
    memset(&P, 0, sizeof(P));
    P.dets = wfn->dets;
    P.nw = (int)((wfn->nbasis + 63) / 64);
    P.W = P.nw * (wfn->kind == PYCI_FULLCI ? 2 : 1);
    P.kind = wfn->kind;
    P.n = (int)wfn->nbasis;
    P.na = (int)wfn->nocc_up;
    P.nb = wfn->kind == PYCI_FULLCI ? (int)wfn->nocc_dn : 0;
    P.ndet = wfn->ndet;
    P.row0 = op->row0;
    P.nloc = op->nloc;
    P.ncol = op->ncol;
    P.one_mo = ham->one_mo;
    P.two_mo = ham->two_mo;
    P.h = ham->h;
    P.v = ham->v;
    P.w = ham->w;
    P.slots = reinterpret_cast<const int *>(wfn->slots);
    P.mask = wfn->mask;
    P.diag = op->diag;
    const long nloc = op->nloc;
    int *rowcnt = nullptr, *tcols = nullptr;
    double *tvals = nullptr;
    PYCI_CUDA(cudaEventRecord(ctx->ev[0], st));
    auto body = [&]() -> int {
        PYCI_CUDA(dev_malloc(&rowcnt, sizeof(int) * (size_t)(nloc + 1)));
        P.rowcnt = rowcnt;
        const unsigned grid = (unsigned)std::max<long>(1, std::min<long>((nloc + MW_WARPS - 1) / MW_WARPS, (long)ctx->sm_count * 8));
        if (nloc > 0) {
            mw_rows_kernel<false><<<grid, 32 * MW_WARPS, 0, st>>>(P);
            ctx->launches++;
        }
        PYCI_TRY(scan_counts(ctx, rowcnt, nloc, op->indptr, nullptr));
        long nnz = 0;
        PYCI_CUDA(cudaMemcpyAsync(&nnz, op->indptr + nloc, sizeof(long), cudaMemcpyDeviceToHost, st));
        PYCI_CUDA(cudaEventRecord(ctx->ev[1], st));
        PYCI_CUDA(cudaStreamSynchronize(st));
        op->nnz = nnz;
        PYCI_CUDA(dev_malloc(&op->vals, sizeof(double) * (size_t)(nnz + 4)));
        PYCI_CUDA(dev_malloc(&op->cols, sizeof(int) * (size_t)(nnz + 4)));
        PYCI_CUDA(dev_malloc(&tvals, sizeof(double) * (size_t)(nnz + 4)));
        PYCI_CUDA(dev_malloc(&tcols, sizeof(int) * (size_t)(nnz + 4)));
        P.indptr = op->indptr;
        P.tcols = tcols;
        P.tvals = tvals;
        PYCI_CUDA(cudaEventRecord(ctx->ev[2], st));
        if (nloc > 0) {
            mw_rows_kernel<true><<<grid, 32 * MW_WARPS, 0, st>>>(P);
            mw_sort_kernel<<<(unsigned)std::max<long>(1, std::min<long>((nloc + 3) / 4, (long)ctx->sm_count * 8)), 128, 0, st>>>(
                op->indptr, tcols, tvals, op->row0, nloc, op->cols, op->vals, op->lowcnt);
            ctx->launches += 2;
        }
        PYCI_CUDA(cudaEventRecord(ctx->ev[3], st));
        PYCI_CUDA(cudaStreamSynchronize(st));
        PYCI_CUDA(cudaGetLastError());
        return PYCI_OK;
    };
    const int rc = body();
    dev_free(rowcnt);
    dev_free(tcols);
    dev_free(tvals);
    PYCI_TRY(rc);
    float ms01 = 0, ms23 = 0;
    cudaEventElapsedTime(&ms01, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ms23, ctx->ev[2], ctx->ev[3]);
    op->times[0] = wfn->hash_seconds;
    op->times[1] = ms01 * 1e-3;
    op->times[2] = ms23 * 1e-3;
    op->times[3] = op->times[1] + op->times[2];
    op->fill_seconds = op->times[2];
    op->fill_kernel = "mw_rows_kernel";
    op->count_kernel = "mw_rows_kernel";
    op->size_ref = op->symmetric ? -1 : op->nnz;
    return PYCI_OK;
}
