// fp64 CSR sparse matrix-vector product: the device form of SparseOp::perform_op /
// perform_op_symm (/root/reference/pyci/src/sparseop.cpp:96-112).
//
// The device keeps FULL rows (both triangles) with int32 columns, so the symmetric product is the
// same gather kernel as the general one: no atomics, deterministic summation order.  32..256 threads
// stream one row: values as 16-byte (double2) and columns as 8-byte (int2) read-only loads that
// bypass L1 allocation (each is used once), x gathered through the read-only path (it is re-used
// across rows and lives in L2), shuffle reduction, one store per row.  Algorithmic traffic is
// 12 B per stored non-zero + 8 B (indptr) + 8 B (y) per row + 8 B per column of x.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace {

__device__ __forceinline__ double2 ld_stream_f64x2(const double *p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ int2 ld_stream_s32x2(const int *p) {
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double ld_stream_f64(const double *p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ int ld_stream_s32(const int *p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

constexpr int SPMV_BLOCK = 256;

// Short rows (selected CI, DOCI: ~10^2 entries): one warp per row, no register pipeline (a row is only a
// few trips), 32 registers so that 8 CTAs = 64 warps per SM hide the per-row latency chain
// (row pointer -> values/columns -> x gathers -> reduction).
__global__ void __launch_bounds__(SPMV_BLOCK, 8)
spmv_short_rows(const long *__restrict__ indptr, const int *__restrict__ cols, const double *__restrict__ vals,
                const double *__restrict__ x, double *__restrict__ y, long nrows) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * SPMV_BLOCK + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * SPMV_BLOCK) >> 5;
    for (long r = warp; r < nrows; r += nwarps) {
        const long start = __ldg(indptr + r), end = __ldg(indptr + r + 1);
        double acc0 = 0.0, acc1 = 0.0;
        long p = start;
        if ((p & 1) && p < end) {
            if (lane == 0)
                acc0 = ld_stream_f64(vals + p) * __ldg(x + ld_stream_s32(cols + p));
            ++p;
        }
        const long nvec = (end - p) >> 1;
        long q = lane;
        for (; q + 32 < nvec; q += 64) {
            const double2 v0 = ld_stream_f64x2(vals + p + 2 * q);
            const int2 c0 = ld_stream_s32x2(cols + p + 2 * q);
            const double2 v1 = ld_stream_f64x2(vals + p + 2 * (q + 32));
            const int2 c1 = ld_stream_s32x2(cols + p + 2 * (q + 32));
            const double x00 = __ldg(x + c0.x), x01 = __ldg(x + c0.y);
            const double x10 = __ldg(x + c1.x), x11 = __ldg(x + c1.y);
            acc0 = fma(v0.x, x00, acc0);
            acc1 = fma(v0.y, x01, acc1);
            acc0 = fma(v1.x, x10, acc0);
            acc1 = fma(v1.y, x11, acc1);
        }
        if (q < nvec) {
            const double2 v0 = ld_stream_f64x2(vals + p + 2 * q);
            const int2 c0 = ld_stream_s32x2(cols + p + 2 * q);
            acc0 = fma(v0.x, __ldg(x + c0.x), acc0);
            acc1 = fma(v0.y, __ldg(x + c0.y), acc1);
        }
        const long tail = p + 2 * nvec;
        if (tail < end && lane == 31)
            acc1 = fma(ld_stream_f64(vals + tail), __ldg(x + ld_stream_s32(cols + tail)), acc1);
        double acc = acc0 + acc1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0)
            y[r] = acc;
    }
}

// TPR threads (a power of two, 32..256) stream one row; a CTA of 256 threads holds 256/TPR rows at a time.
// Long rows use more threads per row: fewer rows are in flight, so the set of 2 MB pages being streamed
// stays small, and every thread still issues two independent 16-byte value loads per trip.
template<int TPR>
__global__ void __launch_bounds__(SPMV_BLOCK)
spmv_rows(const long *__restrict__ indptr, const int *__restrict__ cols, const double *__restrict__ vals,
          const double *__restrict__ x, double *__restrict__ y, long nrows) {
    constexpr int RPB = SPMV_BLOCK / TPR; // rows per block
    constexpr int WPR = TPR / 32;         // warps per row
    __shared__ double partial[SPMV_BLOCK / 32];
    const int lane = threadIdx.x & 31;
    const int sub = threadIdx.x / TPR, t = threadIdx.x % TPR;
    for (long base = (long)blockIdx.x * RPB; base < nrows; base += (long)gridDim.x * RPB) {
        const long r = base + sub;
        double acc0 = 0.0, acc1 = 0.0;
        if (r < nrows) {
            const long start = __ldg(indptr + r), end = __ldg(indptr + r + 1);
            // peel to an even element index so the 16-byte / 8-byte vector loads are aligned
            long p = start;
            if ((p & 1) && p < end) {
                if (t == 0)
                    acc0 = ld_stream_f64(vals + p) * __ldg(x + ld_stream_s32(cols + p));
                ++p;
            }
            const long nvec = (end - p) >> 1; // pairs
            const double *vp = vals + p;
            const int *cp = cols + p;
            // software pipeline: the (value, column) pairs of trip k+1 are requested before the x gathers of
            // trip k are waited for, so the HBM stream never drains while a warp sits on its gathers
            long q = t;
            double2 v0 = make_double2(0.0, 0.0), v1 = v0;
            int2 c0 = make_int2(0, 0), c1 = c0;
            bool h0 = q < nvec, h1 = q + TPR < nvec;
            if (h0) {
                v0 = ld_stream_f64x2(vp + 2 * q);
                c0 = ld_stream_s32x2(cp + 2 * q);
            }
            if (h1) {
                v1 = ld_stream_f64x2(vp + 2 * (q + TPR));
                c1 = ld_stream_s32x2(cp + 2 * (q + TPR));
            }
            while (h0) {
                const long qn = q + 2 * TPR;
                const bool n0 = qn < nvec, n1 = qn + TPR < nvec;
                double2 w0 = make_double2(0.0, 0.0), w1 = w0;
                int2 d0 = make_int2(0, 0), d1 = d0;
                if (n0) {
                    w0 = ld_stream_f64x2(vp + 2 * qn);
                    d0 = ld_stream_s32x2(cp + 2 * qn);
                }
                if (n1) {
                    w1 = ld_stream_f64x2(vp + 2 * (qn + TPR));
                    d1 = ld_stream_s32x2(cp + 2 * (qn + TPR));
                }
                const double x00 = __ldg(x + c0.x), x01 = __ldg(x + c0.y);
                acc0 = fma(v0.x, x00, acc0);
                acc1 = fma(v0.y, x01, acc1);
                if (h1) {
                    const double x10 = __ldg(x + c1.x), x11 = __ldg(x + c1.y);
                    acc0 = fma(v1.x, x10, acc0);
                    acc1 = fma(v1.y, x11, acc1);
                }
                v0 = w0; c0 = d0; v1 = w1; c1 = d1;
                h0 = n0; h1 = n1;
                q = qn;
            }
            const long tail = p + 2 * nvec;
            if (tail < end && t == TPR - 1)
                acc1 = fma(ld_stream_f64(vals + tail), __ldg(x + ld_stream_s32(cols + tail)), acc1);
        }
        double acc = acc0 + acc1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (WPR == 1) {
            if (lane == 0 && r < nrows)
                y[r] = acc;
        } else {
            if (lane == 0)
                partial[threadIdx.x >> 5] = acc;
            __syncthreads();
            if (t == 0 && r < nrows) {
                double sum = 0.0;
#pragma unroll
                for (int wq = 0; wq < WPR; ++wq)
                    sum += partial[sub * WPR + wq];
                y[r] = sum;
            }
            __syncthreads();
        }
    }
}

} // namespace

int spmv_launch(pyci_op *op, const double *x_dev, double *y_dev) {
    pyci_ctx *ctx = op->ctx;
    if (op->nloc <= 0)
        return PYCI_OK;
    // threads per row from the mean row length (measured: 64 threads x 4 CTAs/SM is best for rows of ~2000)
    if (op->spmv_tpr == 0) {
        const long avg = op->nnz / std::max<long>(op->nloc, 1);
        int tpr = avg >= 512 ? 64 : avg >= 320 ? 32 : 1; // 1 = spmv_short_rows
        if (const char *e = getenv("PYCI_B200_SPMV_TPR")) // tuning knob
            tpr = atoi(e);
        op->spmv_tpr = (tpr == 256 || tpr == 128 || tpr == 64 || tpr == 1) ? tpr : 32;
        if (tpr == 1)
            op->spmv_ctas = 8;
        if (const char *e = getenv("PYCI_B200_SPMV_CTAS"))
            op->spmv_ctas = std::max(1, atoi(e));
    }
    const int tpr = op->spmv_tpr;
    if (tpr == 1) {
        const long blocks = (op->nloc * 32 + SPMV_BLOCK - 1) / SPMV_BLOCK;
        const long g = std::min<long>(blocks, (long)ctx->sm_count * op->spmv_ctas);
        spmv_short_rows<<<(unsigned)g, SPMV_BLOCK, 0, ctx->stream>>>(op->indptr, op->cols, op->vals, x_dev, y_dev, op->nloc);
        ctx->launches++;
        PYCI_CUDA(cudaGetLastError());
        return PYCI_OK;
    }
    const long rpb = SPMV_BLOCK / tpr;
    const long blocks_needed = (op->nloc + rpb - 1) / rpb;
    // persistent-ish grid: a multiple of the SM count
    const long grid = std::min<long>(blocks_needed, (long)ctx->sm_count * op->spmv_ctas);
    switch (tpr) {
    case 256:
        spmv_rows<256><<<(unsigned)grid, SPMV_BLOCK, 0, ctx->stream>>>(op->indptr, op->cols, op->vals, x_dev, y_dev, op->nloc);
        break;
    case 128:
        spmv_rows<128><<<(unsigned)grid, SPMV_BLOCK, 0, ctx->stream>>>(op->indptr, op->cols, op->vals, x_dev, y_dev, op->nloc);
        break;
    case 64:
        spmv_rows<64><<<(unsigned)grid, SPMV_BLOCK, 0, ctx->stream>>>(op->indptr, op->cols, op->vals, x_dev, y_dev, op->nloc);
        break;
    default:
        spmv_rows<32><<<(unsigned)grid, SPMV_BLOCK, 0, ctx->stream>>>(op->indptr, op->cols, op->vals, x_dev, y_dev, op->nloc);
        break;
    }
    ctx->launches++;
    PYCI_CUDA(cudaGetLastError());
    return PYCI_OK;
}
