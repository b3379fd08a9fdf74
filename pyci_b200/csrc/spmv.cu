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
                const double *__restrict__ x, double *__restrict__ y, long nrows, long chunk) {
    const int lane = threadIdx.x & 31;
    // chunk > 0: every CTA walks a contiguous range of `chunk` rows, eight (one per warp) at a time -- neighbouring
    // rows of a selected space gather neighbouring x, so the sectors one batch pulled into L1 serve the next;
    // chunk == 0: rows dealt round-robin over all warps of the grid
    const long warp = chunk > 0 ? (long)blockIdx.x * chunk + (threadIdx.x >> 5)
                                : ((long)blockIdx.x * SPMV_BLOCK + threadIdx.x) >> 5;
    const long nwarps = chunk > 0 ? (long)(SPMV_BLOCK / 32) : ((long)gridDim.x * SPMV_BLOCK) >> 5;
    const long rend = chunk > 0 ? min(nrows, ((long)blockIdx.x + 1) * chunk) : nrows;
    for (long r = warp; r < rend; r += nwarps) {
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

// The same rows with lane l of a trip on entry 32 k + l (scalar loads) instead of on the pair (2q, 2q + 1).  ncu of the
// kernel above (profiles/r2v): the L1 data pipe is what is full -- l1tex__data_pipe_lsu_wavefronts at 85 % of its
// peak, ~17 wavefronts per gather instruction -- not HBM (58 %) and not the issue slots (42 %).  The columns of a row
// ascend, and in a selected space 31 % of neighbouring entries gather from the same 128-byte line of x (14 % from the
// same sector): with the pair mapping the two fall into different gather instructions (x of the even entries, then x
// of the odd ones); here neighbouring entries sit in neighbouring lanes of ONE instruction and share its wavefront.
// TRIPS x 32 entries of a row are requested before the first gather is waited for.
template<int TRIPS, int MINB>
__global__ void __launch_bounds__(SPMV_BLOCK, MINB)
spmv_short_rows_seq(const long *__restrict__ indptr, const int *__restrict__ cols, const double *__restrict__ vals,
                    const double *__restrict__ x, double *__restrict__ y, long nrows, long chunk) {
    const int lane = threadIdx.x & 31;
    constexpr int NW = SPMV_BLOCK / 32;
    const long row0 = (long)blockIdx.x * chunk;
    const int nr = (int)(min(nrows, row0 + chunk) - row0); // rows of this CTA
    const long *ip = indptr + row0;
    for (int i = threadIdx.x >> 5; i < nr; i += NW) {
        const long start = __ldg(ip + i);
        const int len = (int)(__ldg(ip + i + 1) - start);
        const double *vp = vals + start;
        const int *cp = cols + start;
        double acc = 0.0;
        for (int e = lane; e < len; e += 32 * TRIPS) {
            double v[TRIPS];
            int c[TRIPS];
#pragma unroll
            for (int k = 0; k < TRIPS; ++k) {
                // (slots beyond the row's end hold zero and column 0: their products vanish)
                v[k] = 0.0;
                c[k] = 0;
                if (e + 32 * k < len) {
                    v[k] = ld_stream_f64(vp + e + 32 * k);
                    c[k] = ld_stream_s32(cp + e + 32 * k);
                }
            }
#pragma unroll
            for (int k = 0; k < TRIPS; ++k)
                acc = fma(v[k], __ldg(x + c[k]), acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0)
            y[row0 + i] = acc;
    }
}

// (ncu of the entry-per-lane kernel, profiles/r3o: 628 us against 723 us on 2 M determinants, DRAM 58 -> 66 %, the L1
// data pipe 85 -> 93 % busy -- still the limit.  The reduction's shuffles go through the same pipe; reducing two rows
// of a warp together, five shuffle steps per two rows, was measured and is slower, 1.83 against 1.75 ms: the second
// row waits behind the first.)
// Measured on the 5 M-determinant selected space (2.02 ms, 0.72 of the HBM peak; ncu: 81 % of the warp samples wait on
// a long scoreboard, L1 hit rate of the gathers 59 %) and NOT kept -- none of them moves the number, the x gathers are
// what the warps wait for: touching the warp's next row with prefetch.global.L2 (2.007 ms); three / four trips of a
// row requested before the first gather at 40 / 46 registers (2.45 / 2.27 ms: the occupancy lost costs more); an L2
// evict-first policy on the matrix stream so that x stays resident (2.021 ms; 1.006 -> 1.007 of peak on the long rows
// of config 3, 0.904 -> 0.910 at config 4 on one GPU); the bulk-copy stream kernel below (3.9 ms).

// TPR threads (a power of two, 32..256) stream one row; a CTA of 256 threads holds 256/TPR rows at a time.
// Long rows use more threads per row: fewer rows are in flight, so the set of 2 MB pages being streamed
// stays small, and every thread still issues two independent 16-byte value loads per trip.
template<int TPR, int BLOCK, int DEPTH>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK)
spmv_rows(const long *__restrict__ indptr, const int *__restrict__ cols, const double *__restrict__ vals,
          const double *__restrict__ x, double *__restrict__ y, long nrows) {
    constexpr int RPB = BLOCK / TPR; // rows per block, streamed in lockstep: consecutive rows gather from the same
                                     // stretches of x at the same time, so their sectors are shared in L1
    constexpr int WPR = TPR / 32;    // warps per row
    __shared__ double partial[BLOCK / 32];
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
            // software pipeline, DEPTH trips deep: a trip is two (value pair, column pair) loads per thread; the
            // trips k+1 .. k+DEPTH-1 are in flight while the x gathers of trip k are waited for, so the HBM stream
            // never drains while a warp sits on its gathers.  Slots are compile-time indices (no register moves).
            double2 v[DEPTH][2];
            int2 c[DEPTH][2];
            long q = t; // first pair of slot 0's trip
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) {
                const long qd = q + (long)d * 2 * TPR;
                v[d][0] = v[d][1] = make_double2(0.0, 0.0);
                c[d][0] = c[d][1] = make_int2(0, 0);
                if (qd < nvec) {
                    v[d][0] = ld_stream_f64x2(vp + 2 * qd);
                    c[d][0] = ld_stream_s32x2(cp + 2 * qd);
                }
                if (qd + TPR < nvec) {
                    v[d][1] = ld_stream_f64x2(vp + 2 * (qd + TPR));
                    c[d][1] = ld_stream_s32x2(cp + 2 * (qd + TPR));
                }
            }
            while (q < nvec) {
#pragma unroll
                for (int d = 0; d < DEPTH; ++d) {
                    const long qd = q + (long)d * 2 * TPR;
                    // slots beyond the row's end hold zeros and column 0: their products vanish, no predicate needed
                    const double x00 = __ldg(x + c[d][0].x), x01 = __ldg(x + c[d][0].y);
                    const double x10 = __ldg(x + c[d][1].x), x11 = __ldg(x + c[d][1].y);
                    const double2 u0 = v[d][0], u1 = v[d][1];
                    const long qn = qd + (long)DEPTH * 2 * TPR; // refill this slot before waiting on the gathers
                    v[d][0] = v[d][1] = make_double2(0.0, 0.0);
                    c[d][0] = c[d][1] = make_int2(0, 0);
                    if (qn < nvec) {
                        v[d][0] = ld_stream_f64x2(vp + 2 * qn);
                        c[d][0] = ld_stream_s32x2(cp + 2 * qn);
                    }
                    if (qn + TPR < nvec) {
                        v[d][1] = ld_stream_f64x2(vp + 2 * (qn + TPR));
                        c[d][1] = ld_stream_s32x2(cp + 2 * (qn + TPR));
                    }
                    acc0 = fma(u0.x, x00, acc0);
                    acc1 = fma(u0.y, x01, acc1);
                    acc0 = fma(u1.x, x10, acc0);
                    acc1 = fma(u1.y, x11, acc1);
                }
                q += (long)DEPTH * 2 * TPR;
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
            if (BLOCK > 256)
                __syncthreads(); // keep the rows of a batch in lockstep (L1 sharing of the x gathers)
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


// ---- bulk-copy (TMA) staged stream ------------------------------------------------------------------
// Long rows, the Davidson hot loop.  Every warp owns a contiguous range of rows (equal stored entries per warp),
// i.e. one sequential stream of values and one of columns, and pulls it through a private ring of shared-memory
// tiles with cp.async.bulk (UBLKCP) completing on an mbarrier: up to ~200 KB per SM are in flight without holding
// a register, against ~50 KB for register-staged loads at the occupancy the gathers allow.  A lane takes pairs
// (16-byte value pair + 8-byte column pair, conflict-free), gathers x through the read-only path and keeps a
// private accumulator for the current row; lanes are reduced once per row, in a fixed order (deterministic).
// No block-level synchronisation after the ring is set up: producer and consumer of a ring are the same warp.
constexpr int ST_TW = 256;                     // entries per tile: 2 KB of values + 1 KB of columns
constexpr int ST_NP = ST_TW / 64;              // pairs per lane per tile
constexpr int ST_TILE_BYTES = ST_TW * 12;

__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@!p bra WAIT_%=;\n"
                 "}" ::"r"(bar), "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void *src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// part[k] = first row of warp k's range: smallest row whose first entry is at or beyond k/nparts of the stored entries
__global__ void spmv_partition_kernel(const long *__restrict__ indptr, long nrows, int nparts, long *__restrict__ part) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nparts)
        return;
    if (k == nparts) {
        part[k] = nrows;
        return;
    }
    const long nnz = indptr[nrows];
    const long target = (long)(((__int128)nnz * k) / nparts);
    long lo = 0, hi = nrows; // smallest r in [0, nrows] with indptr[r] >= target
    while (lo < hi) {
        const long mid = (lo + hi) >> 1;
        if (indptr[mid] >= target)
            hi = mid;
        else
            lo = mid + 1;
    }
    part[k] = lo;
}

template<int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1)
spmv_stream(const long *__restrict__ indptr, const int *__restrict__ cols, const double *__restrict__ vals,
            const double *__restrict__ x, double *__restrict__ y, const long *__restrict__ part, int depth) {
    extern __shared__ __align__(128) unsigned char st_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *ring = st_smem + (size_t)warp * depth * ST_TILE_BYTES;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(st_smem + (size_t)WARPS * depth * ST_TILE_BYTES) + warp * depth;
    const u32 ring_s = (u32)__cvta_generic_to_shared(ring), bars_s = (u32)__cvta_generic_to_shared(bars);
    const long gw = (long)blockIdx.x * WARPS + warp;
    const long r0 = part[gw], r1 = part[gw + 1];
    if (r0 >= r1)
        return;
    const long e0 = __ldg(indptr + r0), e1 = __ldg(indptr + r1);
    const long base = e0 & ~3L; // tiles start on a 16-byte boundary of the column stream
    const long ntiles = (e1 - base + ST_TW - 1) / ST_TW;
    if (lane == 0) {
        for (int s = 0; s < depth; ++s)
            mbar_init(bars_s + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](long t, int s) { // lane 0: request tile t into stage s = t % depth
        const long lo = base + t * ST_TW;
        const u32 cnt = (u32)min((long)ST_TW, ((e1 - lo) + 3) & ~3L); // entries, multiple of 4 (arrays are padded)
        const u32 dst = ring_s + s * ST_TILE_BYTES, bar = bars_s + 8 * s;
        mbar_expect_tx(bar, cnt * 12u);
        bulk_g2s(dst, vals + lo, cnt * 8u, bar);
        bulk_g2s(dst + ST_TW * 8, cols + lo, cnt * 4u, bar);
    };
    if (lane == 0)
        for (int t = 0; t < min((long)depth, ntiles); ++t)
            issue(t, t);

    long row = r0, pos = e0;
    long row_end = __ldg(indptr + row + 1);
    long next_end = (row + 2 <= r1) ? __ldg(indptr + row + 2) : e1; // one row ahead: its latency hides behind a row
    double acc0 = 0.0, acc1 = 0.0;
    int s = 0;
    u32 parity = 0;
    for (long t = 0; t < ntiles; ++t, ++s) {
        if (s == depth) {
            s = 0;
            parity ^= 1u;
        }
        const long lo = base + t * ST_TW, hi = min(lo + ST_TW, e1);
        mbar_wait(bars_s + 8 * s, parity);
        const double2 *tv = reinterpret_cast<const double2 *>(ring + s * ST_TILE_BYTES);
        const int2 *tc = reinterpret_cast<const int2 *>(ring + s * ST_TILE_BYTES + ST_TW * 8);
        double2 v[ST_NP];
        int2 c[ST_NP];
#pragma unroll
        for (int j = 0; j < ST_NP; ++j) {
            v[j] = tv[lane + 32 * j];
            c[j] = tc[lane + 32 * j];
        }
        __syncwarp(); // every lane holds its part of the tile: the stage can be refilled
        if (lane == 0 && t + depth < ntiles)
            issue(t + depth, s);
        const int npair = (int)((hi - lo + 1) >> 1); // pairs holding at least one entry below hi
        double xa[ST_NP], xb[ST_NP];
#pragma unroll
        for (int j = 0; j < ST_NP; ++j) {
            const bool in = lane + 32 * j < npair; // columns beyond the stream's end are not ours to dereference
            xa[j] = in ? __ldg(x + c[j].x) : 0.0;
            xb[j] = (in && lo + 2 * (lane + 32 * j) + 1 < hi) ? __ldg(x + c[j].y) : 0.0;
        }
        if (pos == lo && row_end >= lo + ST_TW) { // the whole tile lies inside the current row
#pragma unroll
            for (int j = 0; j < ST_NP; ++j) {
                acc0 = fma(v[j].x, xa[j], acc0);
                acc1 = fma(v[j].y, xb[j], acc1);
            }
            pos = lo + ST_TW;
        } else {
            for (;;) {
                while (row < r1 && row_end <= pos) { // finished (or empty) rows
                    double a = acc0 + acc1;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
                        a += __shfl_xor_sync(0xffffffffu, a, o);
                    if (lane == 0)
                        y[row] = a;
                    acc0 = acc1 = 0.0;
                    ++row;
                    row_end = next_end;
                    next_end = (row + 2 <= r1) ? __ldg(indptr + row + 2) : e1;
                }
                if (pos >= hi || row >= r1)
                    break;
                const long seg_hi = min(row_end, hi);
                const int a = (int)(pos - lo), b = (int)(seg_hi - lo); // tile-relative entry range of this row
#pragma unroll
                for (int j = 0; j < ST_NP; ++j) {
                    const int e = 2 * (lane + 32 * j);
                    if (e >= a && e < b)
                        acc0 = fma(v[j].x, xa[j], acc0);
                    if (e + 1 >= a && e + 1 < b)
                        acc1 = fma(v[j].y, xb[j], acc1);
                }
                pos = seg_hi;
            }
        }
    }
    while (row < r1) { // rows that end exactly at the end of the stream (and trailing empty rows)
        double a = acc0 + acc1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0)
            y[row] = a;
        acc0 = acc1 = 0.0;
        ++row;
    }
}


template<int WARPS>
int stream_launch_t(pyci_op *op, const double *x_dev, double *y_dev, int depth) {
    pyci_ctx *ctx = op->ctx;
    const size_t smem = (size_t)WARPS * depth * ST_TILE_BYTES + (size_t)WARPS * depth * 8;
    PYCI_CUDA(cudaFuncSetAttribute(spmv_stream<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    spmv_stream<WARPS><<<(unsigned)ctx->sm_count, 32 * WARPS, smem, ctx->stream>>>(op->indptr, op->cols, op->vals, x_dev,
                                                                                 y_dev, op->spmv_part, depth);
    ctx->launches++;
    PYCI_CUDA(cudaGetLastError());
    return PYCI_OK;
}

int stream_launch(pyci_op *op, const double *x_dev, double *y_dev) {
    pyci_ctx *ctx = op->ctx;
    const int warps = -op->spmv_tpr;
    const int maxdepth = (int)(((long)ctx->smem_optin - 1024) / ((long)warps * (ST_TILE_BYTES + 8)));
    const int depth = std::max(1, std::min(op->spmv_ctas, maxdepth));
    const int nparts = ctx->sm_count * warps;
    if (op->spmv_part_n != nparts) { // row ranges of equal stored entries, one per warp (once per operator and shape)
        dev_free(op->spmv_part);
        op->spmv_part = nullptr;
        PYCI_CUDA(dev_malloc(&op->spmv_part, sizeof(long) * (size_t)(nparts + 1)));
        spmv_partition_kernel<<<(nparts + 1 + 255) / 256, 256, 0, ctx->stream>>>(op->indptr, op->nloc, nparts, op->spmv_part);
        ctx->launches++;
        PYCI_CUDA(cudaGetLastError());
        op->spmv_part_n = nparts;
    }
    switch (warps) {
    case 8:
        return stream_launch_t<8>(op, x_dev, y_dev, depth);
    case 16:
        return stream_launch_t<16>(op, x_dev, y_dev, depth);
    case 24:
        return stream_launch_t<24>(op, x_dev, y_dev, depth);
    default:
        return stream_launch_t<32>(op, x_dev, y_dev, depth);
    }
}

} // namespace

int spmv_launch(pyci_op *op, const double *x_dev, double *y_dev) {
    PYCI_NVTX("pyci:spmv");
    pyci_ctx *ctx = op->ctx;
    if (op->nloc <= 0)
        return PYCI_OK;
    // launch shape from the mean row length.  Measured on cfg3 (rows of 2221, profiles/): one warp per row, 16
    // consecutive rows in lockstep per CTA of 512 threads, 2 CTAs per SM, 2 trips in flight = 0.95 of the measured
    // HBM peak; 64 threads per row x 4 CTAs of 256 = 0.91; deeper pipelines spill or lose occupancy.
    if (op->spmv_tpr == 0) {
        const long avg = op->nnz / std::max<long>(op->nloc, 1);
        int tpr = avg >= 320 ? 32 : 1; // 1 = spmv_short_rows
        op->spmv_block = avg >= 512 ? 512 : 256;
        op->spmv_depth = 2;
        if (const char *e = getenv("PYCI_B200_SPMV_TPR")) // tuning knobs
            tpr = atoi(e);
        op->spmv_tpr = (tpr == 256 || tpr == 128 || tpr == 64 || tpr == 1 || tpr == -8 || tpr == -16 || tpr == -24 ||
                        tpr == -32) ? tpr : 32;
        if (const char *e = getenv("PYCI_B200_SPMV_DEPTH")) {
            const int d = atoi(e);
            if (d >= 2 && d <= 4)
                op->spmv_depth = d;
        }
        if (const char *e = getenv("PYCI_B200_SPMV_BLOCK")) {
            const int b = atoi(e);
            if (b == 256 || b == 512 || b == 1024)
                op->spmv_block = b;
        }
        op->spmv_ctas = 1024 / op->spmv_block;
        if (tpr == 1)
            op->spmv_ctas = 8;
        if (tpr < 0)
            op->spmv_ctas = (227 * 1024 - 4096) / (-tpr * ST_TILE_BYTES); // deepest ring that fits shared memory
        if (const char *e = getenv("PYCI_B200_SPMV_CTAS"))
            op->spmv_ctas = std::max(1, atoi(e));
    }
    const int tpr = op->spmv_tpr;
    if (tpr < 0)
        return stream_launch(op, x_dev, y_dev);
    if (tpr == 1) {
        const long blocks = (op->nloc * 32 + SPMV_BLOCK - 1) / SPMV_BLOCK;
        const long g = std::min<long>(blocks, (long)ctx->sm_count * op->spmv_ctas);
        // contiguous ranges of 64 rows per CTA: 2.02 ms against 2.66 ms with rows dealt round-robin (5 M-determinant
        // selected space, 9.6 GB; 32-256 rows are within 4 % of each other, 8 and 2048 lose).  PYCI_B200_SPMV_CHUNK=k
        // overrides, 0 = round-robin
        static const long chunk_env = getenv("PYCI_B200_SPMV_CHUNK") ? atol(getenv("PYCI_B200_SPMV_CHUNK")) : 64;
        long chunk = 0, gg = g;
        if (chunk_env > 0) {
            chunk = (chunk_env + 7) & ~7L;
            gg = (op->nloc + chunk - 1) / chunk;
        }
        // the entry-per-lane kernel (contiguous ranges only); PYCI_B200_SPMV_SEQ=0: the pair-per-lane kernel, k: k trips
        // requested per loop iteration.  5 M determinants (x = 40 MB, L2-resident): 2.013 ms -> 1.76 / 1.76 / 1.74 ms at
        // k = 1 / 2 / 3 = 0.83 of the HBM peak, 1.92 / 2.12 ms at 4 / 6.  One rank's shard of the 50 M-determinant
        // operator (x = 400 MB, tools/spmv_shard.py): 3.96 ms -> 2.95 / 4.31 / 3.05 / 3.00 / 3.28 ms at k = 1 / 2 / 3 / 4 / 6
        // (k = 2 is reproducibly the slow one there: the compiler's interleaving of its unrolled loop) -- one trip per
        // iteration, which the compiler unrolls four times and pipelines itself, is the default.
        static const int seq = getenv("PYCI_B200_SPMV_SEQ") ? atoi(getenv("PYCI_B200_SPMV_SEQ")) : 1;
#define PYCI_SEQ(T, B)                                                                                              \
    spmv_short_rows_seq<T, B><<<(unsigned)gg, SPMV_BLOCK, 0, ctx->stream>>>(op->indptr, op->cols, op->vals, x_dev, y_dev, \
                                                                            op->nloc, chunk)
        if (chunk > 0 && seq == 2)
            PYCI_SEQ(2, 8);
        else if (chunk > 0 && seq == 1)
            PYCI_SEQ(1, 8);
        else if (chunk > 0 && seq == 3)
            PYCI_SEQ(3, 8);
        else if (chunk > 0 && seq == 4)
            PYCI_SEQ(4, 6);
        else if (chunk > 0 && seq == 6)
            PYCI_SEQ(6, 5);
        else
            spmv_short_rows<<<(unsigned)gg, SPMV_BLOCK, 0, ctx->stream>>>(op->indptr, op->cols, op->vals, x_dev, y_dev, op->nloc,
                                                                      chunk);
#undef PYCI_SEQ
        ctx->launches++;
        PYCI_CUDA(cudaGetLastError());
        return PYCI_OK;
    }
    const int block = op->spmv_block;
    const long rpb = block / tpr;
    const long blocks_needed = (op->nloc + rpb - 1) / rpb;
    // persistent-ish grid: a multiple of the SM count
    const long grid = std::min<long>(blocks_needed, (long)ctx->sm_count * op->spmv_ctas);
    static const int carve = getenv("PYCI_B200_SPMV_CARVEOUT") ? atoi(getenv("PYCI_B200_SPMV_CARVEOUT")) : -1;
    const int depth = op->spmv_depth;
#define PYCI_SPMV_CASE_D(T, B, D)                                                                                   \
    if (tpr == T && block == B && depth == D) {                                                                     \
        if (carve >= 0)                                                                                             \
            cudaFuncSetAttribute(spmv_rows<T, B, D>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);        \
        spmv_rows<T, B, D><<<(unsigned)grid, B, 0, ctx->stream>>>(op->indptr, op->cols, op->vals, x_dev, y_dev, op->nloc); \
    }
#define PYCI_SPMV_CASE(T, B) PYCI_SPMV_CASE_D(T, B, 2) PYCI_SPMV_CASE_D(T, B, 3) PYCI_SPMV_CASE_D(T, B, 4)
    PYCI_SPMV_CASE(256, 256)
    PYCI_SPMV_CASE(128, 256)
    PYCI_SPMV_CASE(64, 256)
    PYCI_SPMV_CASE(32, 256)
    PYCI_SPMV_CASE(256, 512)
    PYCI_SPMV_CASE(128, 512)
    PYCI_SPMV_CASE(64, 512)
    PYCI_SPMV_CASE(32, 512)
    PYCI_SPMV_CASE(256, 1024)
    PYCI_SPMV_CASE(128, 1024)
    PYCI_SPMV_CASE(64, 1024)
    PYCI_SPMV_CASE(32, 1024)
#undef PYCI_SPMV_CASE
#undef PYCI_SPMV_CASE_D
    ctx->launches++;
    PYCI_CUDA(cudaGetLastError());
    return PYCI_OK;
}
