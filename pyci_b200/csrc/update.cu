// Incremental growth of an operator: the device form of SparseOp::update
// (/root/reference/pyci/src/sparseop.cpp:175-201), which appends the rows of the determinants added to the wave
// function since the operator was built (the step after add_hci in every selected-CI loop,
// pyci/test/test_routines.py:450-453).
//
// The reference stores the lower triangle, so appending rows [n0, n1) is all it does.  The device keeps FULL
// rows for its gather SpMV: the old rows also gain the columns [n0, n1).  Those entries are the transposes of the
// new rows' entries with column < n0, and because every new column index exceeds every old one they go to the END
// of their row, ordered by the new row they come from.  So nothing is enumerated or probed for the old rows:
//
//   1. build rows [n0, n1) x columns [0, n1) with the ordinary construction (count / scan / fill),
//   2. count the new entries per old row, scan -> row pointer of the grown operator,
//   3. copy the old rows to their new places (one warp per row),
//   4. radix-sort the transposed entries by (old row, new row) and drop them behind the old rows,
//   5. append the new rows.
//
// Cost: enumeration for the NEW determinants only, plus one pass over the stored entries.  Exported CSR (lower
// triangle) is bit-identical to a fresh build: rows < n0 export only columns <= row, which are untouched.
#include <algorithm>

#include "radix.cuh"

namespace {

__global__ void count_transposed_kernel(const int *__restrict__ cols, long nnz, int n0, int *__restrict__ add) {
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (long)gridDim.x * blockDim.x) {
        const int c = cols[e];
        if (c < n0)
            atomicAdd(add + c, 1);
    }
}

// len[r] = old row length + transposed entries joining row r
__global__ void grown_length_kernel(const long *__restrict__ old_indptr, const int *__restrict__ add, long n0,
                                    int *__restrict__ len) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n0)
        len[r] = (int)(old_indptr[r + 1] - old_indptr[r]) + add[r];
}

// one warp per row: row r of (sp, sc, sv) -> position dp[r] (+ shift) of (dc, dv)
__global__ void copy_rows_kernel(const long *__restrict__ sp, const int *__restrict__ sc, const double *__restrict__ sv,
                                 const long *__restrict__ dp, long shift, int *__restrict__ dc, double *__restrict__ dv,
                                 long nrows) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long r = warp; r < nrows; r += nwarps) {
        const long s = sp[r], d = dp[r] + shift, cnt = sp[r + 1] - s;
        for (long e = lane; e < cnt; e += 32) {
            dc[d + e] = sc[s + e];
            dv[d + e] = sv[s + e];
        }
    }
}

// new-row entries with column < n0 -> (column << 32 | new row, value), compacted with a warp-aggregated cursor
__global__ void gather_transposed_kernel(const long *__restrict__ indptr, const int *__restrict__ cols,
                                         const double *__restrict__ vals, long nrows, int row0, int n0,
                                         unsigned long long *__restrict__ keys, double *__restrict__ out,
                                         unsigned long long *cursor) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long r = warp; r < nrows; r += nwarps) {
        const long s = indptr[r], e1 = indptr[r + 1];
        for (long base = s; base < e1; base += 32) {
            const long e = base + lane;
            const int c = (e < e1) ? cols[e] : n0;
            const bool take = c < n0;
            const u32 m = __ballot_sync(0xffffffffu, take);
            if (!m)
                continue;
            unsigned long long start = 0;
            if (lane == 0)
                start = atomicAdd(cursor, (unsigned long long)__popc(m));
            start = __shfl_sync(0xffffffffu, start, 0);
            if (take) {
                const unsigned long long d = start + __popc(m & ((1u << lane) - 1u));
                keys[d] = ((unsigned long long)(u32)c << 32) | (u32)(row0 + r);
                out[d] = vals[e];
            }
        }
    }
}

// sorted transposed entry k (old row c = key >> 32, column = new row) -> behind the old entries of row c
__global__ void place_transposed_kernel(const unsigned long long *__restrict__ keys, const double *__restrict__ vals,
                                        long n, const long *__restrict__ addptr, const long *__restrict__ new_indptr,
                                        const long *__restrict__ old_indptr, int *__restrict__ dc, double *__restrict__ dv) {
    for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[k];
        const long c = (long)(key >> 32);
        const long pos = new_indptr[c] + (old_indptr[c + 1] - old_indptr[c]) + (k - addptr[c]);
        dc[pos] = (int)(u32)key;
        dv[pos] = vals[k];
    }
}

__global__ void shift_indptr_kernel(const long *__restrict__ src, long n, long shift, long *__restrict__ dst) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n)
        dst[i] = src[i] + shift;
}

void release_op_arrays(pyci_op &t) {
    dev_free(t.indptr);
    dev_free(t.cols);
    dev_free(t.vals);
    dev_free(t.lowcnt);
    dev_free(t.diag);
    t.indptr = nullptr;
    t.cols = nullptr;
    t.vals = nullptr;
    t.lowcnt = nullptr;
    t.diag = nullptr;
}

} // namespace

// Non-symmetric operators.  The reference appends the rows [nrow, ndet) with ncol = ndet and leaves the rows it
// already has as they are (sparseop.cpp:188-199): an old row keeps the columns it was built with and does not gain
// the new determinants -- the grown operator has FEWER entries than a fresh non-symmetric build of the same wave
// function (tests/golden/update.npz pins this against the compiled reference).  The device does the same: the new
// rows are built with the ordinary construction and placed behind the old arrays; nothing old is touched.
int op_append_rows_impl(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, pyci_op *op) {
    cudaStream_t st = ctx->stream;
    const long n0 = op->nrow, n1 = wfn->ndet, nnew = n1 - n0;
    PYCI_CUDA(cudaEventRecord(ctx->ev[0], st));
    pyci_op T;
    T.ctx = ctx;
    T.nrow = n1;
    T.ncol = n1;
    T.row0 = n0;
    T.nloc = nnew;
    T.npad = std::max<long>(nnew, 1);
    T.symmetric = 0;
    T.ecore = ham->ecore;
    long *nip = nullptr;
    int *ncols = nullptr, *nlow = nullptr;
    double *nvals = nullptr, *ndiag = nullptr;
    auto body = [&]() -> int {
        if (nnew > 0) {
            PYCI_CUDA(dev_malloc(&T.indptr, sizeof(long) * (size_t)(nnew + 1)));
            PYCI_CUDA(dev_malloc(&T.lowcnt, sizeof(int) * (size_t)(nnew + 1)));
            PYCI_CUDA(dev_malloc(&T.diag, sizeof(double) * (size_t)nnew));
            PYCI_CUDA(cudaMemsetAsync(T.lowcnt, 0, sizeof(int) * (size_t)(nnew + 1), st));
            PYCI_CUDA(cudaMemsetAsync(T.diag, 0, sizeof(double) * (size_t)nnew, st));
            PYCI_TRY(op_build_impl(ctx, ham, wfn, &T));
        }
        const long old_total = op->nnz, total = old_total + T.nnz;
        PYCI_CUDA(dev_malloc(&nip, sizeof(long) * (size_t)(n1 + 1)));
        PYCI_CUDA(dev_malloc(&nvals, sizeof(double) * (size_t)(total + 4)));
        PYCI_CUDA(dev_malloc(&ncols, sizeof(int) * (size_t)(total + 4)));
        PYCI_CUDA(dev_malloc(&nlow, sizeof(int) * (size_t)(n1 + 1)));
        PYCI_CUDA(dev_malloc(&ndiag, sizeof(double) * (size_t)std::max<long>(n1, 1)));
        PYCI_CUDA(cudaMemcpyAsync(nip, op->indptr, sizeof(long) * (size_t)(n0 + 1), cudaMemcpyDeviceToDevice, st));
        if (old_total > 0) {
            PYCI_CUDA(cudaMemcpyAsync(ncols, op->cols, sizeof(int) * (size_t)old_total, cudaMemcpyDeviceToDevice, st));
            PYCI_CUDA(cudaMemcpyAsync(nvals, op->vals, sizeof(double) * (size_t)old_total, cudaMemcpyDeviceToDevice, st));
        }
        if (n0 > 0) {
            PYCI_CUDA(cudaMemcpyAsync(nlow, op->lowcnt, sizeof(int) * (size_t)n0, cudaMemcpyDeviceToDevice, st));
            PYCI_CUDA(cudaMemcpyAsync(ndiag, op->diag, sizeof(double) * (size_t)n0, cudaMemcpyDeviceToDevice, st));
        }
        if (nnew > 0) {
            if (T.nnz > 0) {
                PYCI_CUDA(cudaMemcpyAsync(ncols + old_total, T.cols, sizeof(int) * (size_t)T.nnz, cudaMemcpyDeviceToDevice, st));
                PYCI_CUDA(cudaMemcpyAsync(nvals + old_total, T.vals, sizeof(double) * (size_t)T.nnz, cudaMemcpyDeviceToDevice, st));
            }
            shift_indptr_kernel<<<(unsigned)((nnew + 256) / 256), 256, 0, st>>>(T.indptr, nnew, old_total, nip + n0);
            ctx->launches++;
            PYCI_CUDA(cudaMemcpyAsync(nlow + n0, T.lowcnt, sizeof(int) * (size_t)nnew, cudaMemcpyDeviceToDevice, st));
            PYCI_CUDA(cudaMemcpyAsync(ndiag + n0, T.diag, sizeof(double) * (size_t)nnew, cudaMemcpyDeviceToDevice, st));
            PYCI_CUDA(cudaGetLastError());
        }
        dev_free(op->indptr);
        dev_free(op->cols);
        dev_free(op->vals);
        dev_free(op->lowcnt);
        dev_free(op->diag);
        dev_free(op->xbuf);
        dev_free(op->ybuf);
        dev_free(op->spmv_part);
        op->indptr = nip;
        op->cols = ncols;
        op->ge_row = -1;
        op->vals = nvals;
        op->lowcnt = nlow;
        op->diag = ndiag;
        nip = nullptr;
        ncols = nlow = nullptr;
        nvals = ndiag = nullptr;
        op->xbuf = op->ybuf = nullptr;
        op->spmv_part = nullptr;
        op->spmv_part_n = 0;
        op->spmv_tpr = 0;
        op->nrow = op->ncol = n1;
        op->row0 = 0;
        op->nloc = n1;
        op->npad = std::max<long>(1, n1);
        op->nnz = total;
        op->size_ref = total;
        op->ecore = ham->ecore;
        if (nnew > 0)
            op->fill_kernel = T.fill_kernel;
        return PYCI_OK;
    };
    const int rc = body();
    release_op_arrays(T);
    dev_free(nip);
    dev_free(ncols);
    dev_free(nvals);
    dev_free(nlow);
    dev_free(ndiag);
    PYCI_TRY(rc);
    PYCI_CUDA(cudaEventRecord(ctx->ev[1], st));
    PYCI_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    op->times[0] = wfn->hash_seconds;
    op->times[1] = T.times[1];
    op->times[2] = T.times[2];
    op->times[3] = ms * 1e-3;
    return PYCI_OK;
}

int op_update_impl(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, pyci_op *op) {
    if (!op->symmetric)
        return op_append_rows_impl(ctx, ham, wfn, op);
    cudaStream_t st = ctx->stream;
    const long n0 = op->nrow, n1 = wfn->ndet;
    if (n1 == n0)
        return PYCI_OK;
    const long nnew = n1 - n0;
    PYCI_CUDA(cudaEventRecord(ctx->ev[0], st));

    // ---- 1. the new rows, all columns
    pyci_op T;
    T.ctx = ctx;
    T.nrow = n1;
    T.ncol = n1;
    T.row0 = n0;
    T.nloc = nnew;
    T.npad = nnew;
    T.symmetric = 1;
    T.ecore = ham->ecore;
    int *add = nullptr, *len = nullptr;
    long *addptr = nullptr, *nip = nullptr;
    int *ncols = nullptr, *nlow = nullptr;
    double *nvals = nullptr, *ndiag = nullptr, *tvals = nullptr, *tvals2 = nullptr;
    unsigned long long *tkeys = nullptr, *tkeys2 = nullptr, *cursor = nullptr;
    void *tmp = nullptr;
    auto body = [&]() -> int {
        PYCI_CUDA(dev_malloc(&T.indptr, sizeof(long) * (size_t)(nnew + 1)));
        PYCI_CUDA(dev_malloc(&T.lowcnt, sizeof(int) * (size_t)(nnew + 1)));
        PYCI_CUDA(dev_malloc(&T.diag, sizeof(double) * (size_t)nnew));
        PYCI_CUDA(cudaMemsetAsync(T.lowcnt, 0, sizeof(int) * (size_t)(nnew + 1), st));
        PYCI_CUDA(cudaMemsetAsync(T.diag, 0, sizeof(double) * (size_t)nnew, st));
        PYCI_TRY(op_build_impl(ctx, ham, wfn, &T));

        // ---- 2. transposed entries per old row, row pointer of the grown operator
        PYCI_CUDA(dev_malloc(&add, sizeof(int) * (size_t)(n0 + 1)));
        PYCI_CUDA(dev_malloc(&len, sizeof(int) * (size_t)(n0 + 1)));
        PYCI_CUDA(dev_malloc(&addptr, sizeof(long) * (size_t)(n0 + 1)));
        PYCI_CUDA(dev_malloc(&nip, sizeof(long) * (size_t)(n1 + 1)));
        PYCI_CUDA(cudaMemsetAsync(add, 0, sizeof(int) * (size_t)(n0 + 1), st));
        const unsigned gsz = (unsigned)ctx->sm_count * 8;
        if (T.nnz > 0 && n0 > 0) {
            count_transposed_kernel<<<gsz, 256, 0, st>>>(T.cols, T.nnz, (int)n0, add);
            ctx->launches++;
        }
        if (n0 > 0) {
            grown_length_kernel<<<(unsigned)((n0 + 255) / 256), 256, 0, st>>>(op->indptr, add, n0, len);
            ctx->launches++;
        }
        PYCI_TRY(scan_counts(ctx, add, n0, addptr, nullptr));
        PYCI_TRY(scan_counts(ctx, len, n0, nip, nullptr));
        long nt = 0, old_total = 0;
        PYCI_CUDA(cudaMemcpyAsync(&nt, addptr + n0, sizeof(long), cudaMemcpyDeviceToHost, st));
        PYCI_CUDA(cudaMemcpyAsync(&old_total, nip + n0, sizeof(long), cudaMemcpyDeviceToHost, st));
        PYCI_CUDA(cudaStreamSynchronize(st));
        if (nt >= (1L << 31) - 1)
            PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "too many new entries for one incremental update (%ld)", nt);
        const long total = old_total + T.nnz;

        // ---- 3. old rows to their new places
        PYCI_CUDA(dev_malloc(&ncols, sizeof(int) * (size_t)(total + 4)));
        PYCI_CUDA(dev_malloc(&nvals, sizeof(double) * (size_t)(total + 4)));
        PYCI_CUDA(dev_malloc(&nlow, sizeof(int) * (size_t)(n1 + 1)));
        PYCI_CUDA(dev_malloc(&ndiag, sizeof(double) * (size_t)n1));
        if (n0 > 0) {
            copy_rows_kernel<<<gsz, 256, 0, st>>>(op->indptr, op->cols, op->vals, nip, 0, ncols, nvals, n0);
            ctx->launches++;
        }
        // ---- 4. transposed entries, sorted by (old row, new row)
        if (nt > 0) {
            PYCI_CUDA(dev_malloc(&tkeys, sizeof(unsigned long long) * (size_t)nt));
            PYCI_CUDA(dev_malloc(&tkeys2, sizeof(unsigned long long) * (size_t)nt));
            PYCI_CUDA(dev_malloc(&tvals, sizeof(double) * (size_t)nt));
            PYCI_CUDA(dev_malloc(&tvals2, sizeof(double) * (size_t)nt));
            PYCI_CUDA(dev_malloc(&cursor, sizeof(unsigned long long)));
            PYCI_CUDA(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), st));
            gather_transposed_kernel<<<gsz, 256, 0, st>>>(T.indptr, T.cols, T.vals, nnew, (int)n0, (int)n0, tkeys, tvals, cursor);
            ctx->launches++;
            int bits = 1;
            while ((1L << bits) < n1)
                ++bits;
            bool in_alt = false;
            PYCI_TRY(radix_sort_pairs<double>(ctx, tkeys, tkeys2, tvals, tvals2, nt, 32 + bits, &in_alt));
            if (!in_alt) { // place_transposed_kernel reads the *2 buffers
                std::swap(tkeys, tkeys2);
                std::swap(tvals, tvals2);
            }
            place_transposed_kernel<<<gsz, 256, 0, st>>>(tkeys2, tvals2, nt, addptr, nip, op->indptr, ncols, nvals);
            ctx->launches++;
        }
        // ---- 5. the new rows behind them; row pointer, prefix counts, diagonal
        copy_rows_kernel<<<gsz, 256, 0, st>>>(T.indptr, T.cols, T.vals, T.indptr, old_total, ncols, nvals, nnew);
        shift_indptr_kernel<<<(unsigned)((nnew + 256) / 256), 256, 0, st>>>(T.indptr, nnew, old_total, nip + n0);
        ctx->launches += 2;
        PYCI_CUDA(cudaMemcpyAsync(nlow, op->lowcnt, sizeof(int) * (size_t)n0, cudaMemcpyDeviceToDevice, st));
        PYCI_CUDA(cudaMemcpyAsync(nlow + n0, T.lowcnt, sizeof(int) * (size_t)nnew, cudaMemcpyDeviceToDevice, st));
        PYCI_CUDA(cudaMemcpyAsync(ndiag, op->diag, sizeof(double) * (size_t)n0, cudaMemcpyDeviceToDevice, st));
        PYCI_CUDA(cudaMemcpyAsync(ndiag + n0, T.diag, sizeof(double) * (size_t)nnew, cudaMemcpyDeviceToDevice, st));
        PYCI_CUDA(cudaGetLastError());

        // ---- swap in
        dev_free(op->indptr);
        dev_free(op->cols);
        dev_free(op->vals);
        dev_free(op->lowcnt);
        dev_free(op->diag);
        dev_free(op->xbuf);
        dev_free(op->ybuf);
        dev_free(op->spmv_part);
        op->indptr = nip;
        op->cols = ncols;
        op->ge_row = -1;
        op->vals = nvals;
        op->lowcnt = nlow;
        op->diag = ndiag;
        nip = nullptr;
        ncols = nullptr;
        nvals = nullptr;
        nlow = nullptr;
        ndiag = nullptr;
        op->xbuf = op->ybuf = nullptr;
        op->spmv_part = nullptr;
        op->spmv_part_n = 0;
        op->spmv_tpr = 0; // the mean row length changed: choose the launch shape again
        op->nrow = op->ncol = n1;
        op->row0 = 0;
        op->nloc = n1;
        op->npad = std::max<long>(1, n1);
        op->nnz = total;
        op->size_ref = -1; // summed again from the grown lowcnt on first use
        op->ecore = ham->ecore;
        op->fill_kernel = T.fill_kernel;
        return PYCI_OK;
    };
    const int rc = body();
    release_op_arrays(T);
    dev_free(add);
    dev_free(len);
    dev_free(addptr);
    dev_free(nip);
    dev_free(ncols);
    dev_free(nvals);
    dev_free(nlow);
    dev_free(ndiag);
    dev_free(tkeys);
    dev_free(tkeys2);
    dev_free(tvals);
    dev_free(tvals2);
    dev_free(cursor);
    dev_free(tmp);
    PYCI_TRY(rc);
    PYCI_CUDA(cudaEventRecord(ctx->ev[1], st));
    PYCI_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    op->times[0] = wfn->hash_seconds;
    op->times[1] = T.times[1];
    op->times[2] = T.times[2];
    op->times[3] = ms * 1e-3; // the whole update
    return PYCI_OK;
}
