// nnz-balanced row partition of a row-sharded operator ("contiguous row ranges balanced by nnz").
//
// Construction shards the rows uniformly: the number of stored entries of a row is not known before it is built,
// and every rank must know which rows are its own before it starts.  For a complete space every row has the same
// length and that is the end of it.  The rows of a selected space do not (the determinants near the reference
// connect to far more of the space than the tail does), so the gather SpMV -- one pass over the stored entries per
// Davidson iteration, as fast as its slowest rank -- wants equal ENTRIES per rank, not equal rows.
//
// op_rebalance runs once, after the fill: the ranks agree on boundaries b_0 = 0 <= b_1 <= ... <= b_R = nrow such that
// rank p's rows [b_p, b_p+1) hold as close to total/R entries as whole rows allow, and each rank ships the rows it no
// longer owns to their new owners.  Both partitions are contiguous and ordered, so what moves between a pair of
// ranks is one contiguous row range = one contiguous slice of each CSR array: five point-to-point transfers per pair
// (row pointer slice, columns, values, lower-triangle counts, diagonal) in NCCL groups over NVLink, typically only
// between neighbours.  Afterwards op->bounds holds the partition, op->npad the largest row count (the stride of the
// solver's vectors), and op_allgather_rows gathers unequal shards straight into their global positions.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "common.cuh"

namespace {

// thread p < ntargets: the global row at which the running entry count reaches targets[p], if that happens inside
// this rank's rows (global entry offsets [g0, g0 + nnz)); 0 otherwise -- the ranks' answers are summed
__global__ void boundary_rows_kernel(const long *__restrict__ indptr, long nloc, long row0, long g0, long nnz,
                                     const long *__restrict__ targets, int ntargets, long *__restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= ntargets)
        return;
    const long t = targets[p] - g0;
    long ans = 0;
    if (t >= 0 && t < nnz) {
        long lo = 0, hi = nloc; // smallest local row j with indptr[j] >= t
        while (lo < hi) {
            const long mid = (lo + hi) >> 1;
            if (indptr[mid] >= t)
                hi = mid;
            else
                lo = mid + 1;
        }
        ans = row0 + lo;
    }
    out[p] = ans;
}

__global__ void gather_longs_kernel(const long *__restrict__ src, const long *__restrict__ idx, int n, long *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n)
        out[k] = src[idx[k]];
}

// dst[j] = src[j] - src[0] + base for j <= nrows: the row pointer slice of a received piece, re-based to its place
__global__ void rebase_indptr_kernel(const long *__restrict__ src, long nrows, long base, long *__restrict__ dst) {
    const long first = src[0];
    for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j <= nrows; j += (long)gridDim.x * blockDim.x)
        dst[j] = src[j] - first + base;
}

struct Overlap {
    long lo, hi; // global rows
    long n() const { return std::max<long>(0, hi - lo); }
};
inline Overlap overlap(long a0, long a1, long b0, long b1) { return Overlap{std::max(a0, b0), std::min(a1, b1)}; }

// stage[p * ld + j] -> out[bounds[p] + j], j < bounds[p + 1] - bounds[p]: blockIdx.y = p
__global__ void compact_shards_kernel(const double *__restrict__ stage, long ld, const long *__restrict__ bounds,
                                      double *__restrict__ out) {
    const int p = blockIdx.y;
    const long b0 = bounds[p], cnt = bounds[p + 1] - b0;
    const double *src = stage + (size_t)p * ld;
    for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += (long)gridDim.x * blockDim.x)
        out[b0 + j] = src[j];
}

} // namespace

// Unequal shards: ONE ncclAllGather of the common stride (every rank's vectors are allocated with it) into a staging
// buffer, then one pass that closes the gaps -- a group of R broadcasts straight into place moves the same bytes but
// ran at half the rate of the all-gather (50 M rows on 8 GPUs: 1.4 ms against 0.71 ms; the extra pass is 0.13 ms).
// PYCI_B200_GATHER_BCAST=1 selects the broadcast group.
int op_allgather_rows(pyci_ctx *ctx, pyci_op *op, const double *send_dev, double *recv_dev) {
    if (op->bounds.empty())
        return comm_allgather_f64(ctx, send_dev, recv_dev, op->npad);
    static const bool bcast = getenv("PYCI_B200_GATHER_BCAST") != nullptr;
    if (bcast)
        return comm_allgatherv_f64(ctx, send_dev, recv_dev, op->bounds.data());
    const int R = ctx->nranks;
    if (!op->gather_stage) {
        PYCI_CUDA(dev_malloc(&op->gather_stage, sizeof(double) * (size_t)op->npad * (size_t)R));
        PYCI_CUDA(dev_malloc(&op->bounds_dev, sizeof(long) * (size_t)(R + 1)));
        PYCI_CUDA(cudaMemcpyAsync(op->bounds_dev, op->bounds.data(), sizeof(long) * (size_t)(R + 1), cudaMemcpyHostToDevice,
                                  ctx->stream)); // (op->bounds outlives the copy: it changes only with the operator)
    }
    PYCI_TRY(comm_allgather_f64(ctx, send_dev, op->gather_stage, op->npad));
    const dim3 grid((unsigned)std::max<long>(1, std::min<long>((op->npad + 255) / 256, (long)ctx->sm_count * 4)), (unsigned)R);
    compact_shards_kernel<<<grid, 256, 0, ctx->stream>>>(op->gather_stage, op->npad, op->bounds_dev, recv_dev);
    ctx->launches++;
    PYCI_CUDA(cudaGetLastError());
    return PYCI_OK;
}


int op_rebalance(pyci_ctx *ctx, pyci_op *op) {
    const int R = ctx->nranks, me = ctx->rank;
    // (complete spaces: every row has the same length -- no collective is spent on finding that out)
    if (R <= 1 || R > 64 || op->foreign || op->nrow != op->ncol || !strcmp(op->count_kernel, "analytic") ||
        getenv("PYCI_B200_NO_REBALANCE"))
        return PYCI_OK;
    PYCI_NVTX("pyci:rebalance(rows to equal stored entries per rank)");
    cudaStream_t st = ctx->stream;
    const long nrow = op->nrow;

    // ---- entries per rank; nothing to do when the uniform blocks are already even
    std::vector<long> nnz((size_t)R, 0);
    nnz[(size_t)me] = op->nnz;
    PYCI_TRY(comm_allreduce_sum_i64_host(ctx, nnz.data(), R));
    const long total = std::accumulate(nnz.begin(), nnz.end(), 0L);
    const long most = *std::max_element(nnz.begin(), nnz.end());
    double min_ratio = 1.05; // fullest rank more than 5 % over the mean (gathering unequal shards costs an extra pass)
    if (const char *e = getenv("PYCI_B200_REBALANCE_MIN"))
        min_ratio = atof(e);
    if (total <= 0 || nrow < 4L * R || (double)most * R <= min_ratio * (double)total)
        return PYCI_OK;

    // old (uniform) partition
    std::vector<long> ob((size_t)R + 1);
    for (int p = 0; p <= R; ++p)
        ob[(size_t)p] = std::min(nrow, op->npad * p);
    std::vector<long> g((size_t)R + 1, 0); // global entry offset of every rank's first row
    for (int p = 0; p < R; ++p)
        g[(size_t)p + 1] = g[(size_t)p] + nnz[(size_t)p];

    long *dtmp = nullptr, *didx = nullptr, *dstage = nullptr;
    long *nip = nullptr;
    int *ncols = nullptr, *nlow = nullptr;
    double *nvals = nullptr, *ndiag = nullptr;
    auto body = [&]() -> int {
        // ---- new boundaries: the row at which the running entry count reaches p * total / R
        std::vector<long> nb((size_t)R + 1, 0);
        {
            std::vector<long> targets((size_t)R - 1);
            for (int p = 1; p < R; ++p)
                targets[(size_t)p - 1] = (long)((double)total * p / R);
            PYCI_CUDA(dev_malloc(&dtmp, sizeof(long) * 2 * (size_t)R));
            PYCI_CUDA(cudaMemcpyAsync(dtmp, targets.data(), sizeof(long) * (size_t)(R - 1), cudaMemcpyHostToDevice, st));
            boundary_rows_kernel<<<1, 64, 0, st>>>(op->indptr, op->nloc, op->row0, g[(size_t)me], op->nnz, dtmp, R - 1, dtmp + R);
            ctx->launches++;
            std::vector<long> found((size_t)R - 1, 0);
            PYCI_CUDA(cudaMemcpyAsync(found.data(), dtmp + R, sizeof(long) * (size_t)(R - 1), cudaMemcpyDeviceToHost, st));
            PYCI_CUDA(cudaStreamSynchronize(st));
            PYCI_TRY(comm_allreduce_sum_i64_host(ctx, found.data(), R - 1));
            for (int p = 1; p < R; ++p)
                nb[(size_t)p] = std::max(nb[(size_t)p - 1], std::min(nrow, found[(size_t)p - 1]));
            nb[(size_t)R] = nrow;
        }
        const long my0 = nb[(size_t)me], my1 = nb[(size_t)me + 1], nloc2 = my1 - my0;

        // ---- what this rank sends: its old rows that fall into every other rank's new range
        // (entry offsets of the piece ends come from the local row pointer: one small gather)
        std::vector<Overlap> sp((size_t)R), rp((size_t)R);
        std::vector<long> idx;
        for (int p = 0; p < R; ++p) {
            sp[(size_t)p] = overlap(ob[(size_t)me], ob[(size_t)me + 1], nb[(size_t)p], nb[(size_t)p + 1]);
            rp[(size_t)p] = overlap(ob[(size_t)p], ob[(size_t)p + 1], my0, my1);
            const Overlap &o = sp[(size_t)p];
            idx.push_back(o.n() ? o.lo - op->row0 : 0);
            idx.push_back(o.n() ? o.hi - op->row0 : 0);
        }
        std::vector<long> ends((size_t)2 * R, 0);
        PYCI_CUDA(dev_malloc(&didx, sizeof(long) * 4 * (size_t)R));
        PYCI_CUDA(cudaMemcpyAsync(didx, idx.data(), sizeof(long) * 2 * (size_t)R, cudaMemcpyHostToDevice, st));
        gather_longs_kernel<<<(2 * R + 63) / 64, 64, 0, st>>>(op->indptr, didx, 2 * R, didx + 2 * R);
        ctx->launches++;
        PYCI_CUDA(cudaMemcpyAsync(ends.data(), didx + 2 * R, sizeof(long) * 2 * (size_t)R, cudaMemcpyDeviceToHost, st));
        PYCI_CUDA(cudaStreamSynchronize(st));
        // entries sent by q to p, every pair (each rank fills its own row of the matrix)
        std::vector<long> mat((size_t)R * R, 0);
        for (int p = 0; p < R; ++p)
            mat[(size_t)me * R + p] = ends[(size_t)2 * p + 1] - ends[(size_t)2 * p];
        PYCI_TRY(comm_allreduce_sum_i64_host(ctx, mat.data(), R * R));

        // ---- offsets, in rows and in entries, of every piece on both sides
        std::vector<long> s_rows((size_t)R), s_row_off((size_t)R), s_ent((size_t)R), s_ent_off((size_t)R);
        std::vector<long> r_rows((size_t)R), r_row_off((size_t)R + 1, 0), r_ent((size_t)R), r_ent_off((size_t)R + 1, 0);
        for (int p = 0; p < R; ++p) {
            s_rows[(size_t)p] = sp[(size_t)p].n();
            s_row_off[(size_t)p] = s_rows[(size_t)p] ? sp[(size_t)p].lo - op->row0 : 0;
            s_ent[(size_t)p] = mat[(size_t)me * R + p];
            s_ent_off[(size_t)p] = ends[(size_t)2 * p];
            r_rows[(size_t)p] = rp[(size_t)p].n();
            r_ent[(size_t)p] = mat[(size_t)p * R + me];
            r_row_off[(size_t)p + 1] = r_row_off[(size_t)p] + r_rows[(size_t)p]; // senders in rank order = row order
            r_ent_off[(size_t)p + 1] = r_ent_off[(size_t)p] + r_ent[(size_t)p];
        }
        if (r_row_off[(size_t)R] != nloc2)
            PYCI_FAIL(PYCI_ERR_RUNTIME, "rebalance: the pieces of rank %d cover %ld rows, expected %ld", me, r_row_off[(size_t)R], nloc2);
        const long nnz2 = r_ent_off[(size_t)R];
        long ld2 = 1;
        for (int p = 0; p < R; ++p)
            ld2 = std::max(ld2, nb[(size_t)p + 1] - nb[(size_t)p]);

        // ---- the new arrays; row pointer slices (rows + 1 each) are staged and re-based
        PYCI_CUDA(dev_malloc(&nvals, sizeof(double) * (size_t)(nnz2 + 4)));
        PYCI_CUDA(dev_malloc(&ncols, sizeof(int) * (size_t)(nnz2 + 4)));
        PYCI_CUDA(dev_malloc(&nip, sizeof(long) * (size_t)(nloc2 + 1)));
        PYCI_CUDA(dev_malloc(&nlow, sizeof(int) * (size_t)(nloc2 + 1)));
        PYCI_CUDA(dev_malloc(&ndiag, sizeof(double) * (size_t)ld2));
        PYCI_CUDA(dev_malloc(&dstage, sizeof(long) * (size_t)(nloc2 + R + 1)));
        PYCI_CUDA(cudaMemsetAsync(ndiag, 0, sizeof(double) * (size_t)ld2, st));
        PYCI_CUDA(cudaMemsetAsync(nip, 0, sizeof(long) * (size_t)(nloc2 + 1), st));
        PYCI_CUDA(cudaMemsetAsync(nlow, 0, sizeof(int) * (size_t)(nloc2 + 1), st));
        std::vector<long> sc((size_t)R), so((size_t)R), rc((size_t)R), ro((size_t)R);
        auto exchange = [&](const void *src, void *dst, long unit, const std::vector<long> &scount, const std::vector<long> &soff,
                            const std::vector<long> &rcount, const std::vector<long> &roff, long extra) -> int {
            for (int p = 0; p < R; ++p) { // `extra` elements more per non-empty piece (the closing row pointer)
                sc[(size_t)p] = scount[(size_t)p] > 0 || (extra && s_rows[(size_t)p] > 0) ? (scount[(size_t)p] + extra) * unit : 0;
                so[(size_t)p] = soff[(size_t)p] * unit;
                rc[(size_t)p] = rcount[(size_t)p] > 0 || (extra && r_rows[(size_t)p] > 0) ? (rcount[(size_t)p] + extra) * unit : 0;
                ro[(size_t)p] = (roff[(size_t)p] + (extra ? p : 0)) * unit;
            }
            return comm_alltoallv_bytes(ctx, src, sc.data(), so.data(), dst, rc.data(), ro.data());
        };
        PYCI_TRY(exchange(op->vals, nvals, 8, s_ent, s_ent_off, r_ent, r_ent_off, 0));
        PYCI_TRY(exchange(op->cols, ncols, 4, s_ent, s_ent_off, r_ent, r_ent_off, 0));
        PYCI_TRY(exchange(op->lowcnt, nlow, 4, s_rows, s_row_off, r_rows, r_row_off, 0));
        PYCI_TRY(exchange(op->diag, ndiag, 8, s_rows, s_row_off, r_rows, r_row_off, 0));
        PYCI_TRY(exchange(op->indptr, dstage, 8, s_rows, s_row_off, r_rows, r_row_off, 1));
        for (int p = 0; p < R; ++p) {
            if (r_rows[(size_t)p] <= 0)
                continue;
            const long nr = r_rows[(size_t)p];
            rebase_indptr_kernel<<<(unsigned)std::max<long>(1, std::min<long>((nr + 256) / 256, 1024)), 256, 0, st>>>(
                dstage + r_row_off[(size_t)p] + p, nr, r_ent_off[(size_t)p], nip + r_row_off[(size_t)p]);
            ctx->launches++;
        }
        PYCI_CUDA(cudaGetLastError());
        PYCI_CUDA(cudaStreamSynchronize(st)); // the host vectors above are read by the copies; old arrays go next

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
        op->row0 = my0;
        op->nloc = nloc2;
        op->npad = ld2;
        op->nnz = nnz2;
        op->size_ref = op->symmetric ? -1 : nnz2;
        op->bounds = nb;
        return PYCI_OK;
    };
    const int rc = body();
    dev_free(dtmp);
    dev_free(didx);
    dev_free(dstage);
    dev_free(nip);
    dev_free(ncols);
    dev_free(nvals);
    dev_free(nlow);
    dev_free(ndiag);
    return rc;
}
