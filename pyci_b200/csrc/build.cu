// Construction of the CI Hamiltonian in CSR on the device: the B200 form of
// SparseOp::update / add_row / sort_row (/root/reference/pyci/src/sparseop.cpp:186-502).
//
//   index   : open-addressing hash of the determinant strings (replaces Wfn::index_det's
//             SpookyHash + flat_hash_map, pyci.h:121-128,157)
//   count   : one CTA per row, every thread enumerates excitations with bit tricks and probes the
//             hash; per-row hit counts                      (skipped when the space is complete)
//   scan    : int64 exclusive scan of the counts -> indptr
//   fill    : same enumeration; hits are appended (warp-aggregated) to a shared-memory row buffer as
//             (column << 32 | excitation code), sorted by column in shared memory (sort_row,
//             sparseop.cpp:214-218), then matrix elements are evaluated from the codes
//             (Slater-Condon rules with the reference's phases and summation order) and the row is
//             streamed to HBM with coalesced stores.
#include <algorithm>
#include <cstring>

#include "enumerate.cuh"

namespace {

// ---- matrix elements, in the reference's operation order ------------------------------------

// DOCI diagonal, sparseop.cpp:228-236,253: val1 + 2*val2
__device__ double diag_doci(const BuildParams &P, const RowShared &rs) {
    const int n = P.n, no = rs.nocc[0];
    double val1 = 0.0, val2 = 0.0;
    for (int i = 0; i < no; ++i) {
        const int k = rs.occ[0][i];
        val1 += __ldg(P.v + k * (n + 1));
        val2 += __ldg(P.h + k);
        for (int j = i + 1; j < no; ++j)
            val2 += __ldg(P.w + k * n + rs.occ[0][j]);
    }
    return val1 + val2 * 2;
}

// FullCI diagonal, sparseop.cpp:283-292,367-372 (GenCI :443-449 is the alpha-only part)
template<int KIND>
__device__ double diag_twobody(const BuildParams &P, const RowShared &rs) {
    const long n1 = P.n, n2 = n1 * n1, n3 = n2 * n1;
    const int na = rs.nocc[0], nb = (KIND == PYCI_FULLCI) ? rs.nocc[1] : 0;
    double val2 = 0.0;
    for (int i = 0; i < na; ++i) {
        const long ii = rs.occ[0][i], ioff = n3 * ii;
        val2 += __ldg(P.one_mo + (n1 + 1) * ii);
        for (int k = i + 1; k < na; ++k) {
            const long kk = rs.occ[0][k], koff = ioff + n2 * kk;
            val2 += __ldg(P.two_mo + koff + n1 * ii + kk) - __ldg(P.two_mo + koff + n1 * kk + ii);
        }
        for (int k = 0; k < nb; ++k) {
            const long kk = rs.occ[1][k];
            val2 += __ldg(P.two_mo + ioff + n2 * kk + n1 * ii + kk);
        }
    }
    for (int i = 0; i < nb; ++i) {
        const long ii = rs.occ[1][i], ioff = n3 * ii;
        val2 += __ldg(P.one_mo + (n1 + 1) * ii);
        for (int k = i + 1; k < nb; ++k) {
            const long kk = rs.occ[1][k], koff = ioff + n2 * kk;
            val2 += __ldg(P.two_mo + koff + n1 * ii + kk) - __ldg(P.two_mo + koff + n1 * kk + ii);
        }
    }
    return val2;
}

template<int KIND>
__device__ double element(const BuildParams &P, const RowShared &rs, u32 code) {
    const int type = code >> 24;
    const long i = (code >> 18) & 63, a = (code >> 12) & 63, k = (code >> 6) & 63, l = code & 63;
    const long n1 = P.n, n2 = n1 * n1, n3 = n2 * n1;
    switch (type) {
    case T_PAIR: // sparseop.cpp:246: v[k*n + l], no phase
        return __ldg(P.v + i * n1 + a);
    case T_AB: { // sparseop.cpp:330-332
        const int par = parity_single(rs.det[0], (int)i, (int)a) ^ parity_single(rs.det[1], (int)k, (int)l);
        return apply_sign(__ldg(P.two_mo + n3 * i + n2 * k + n1 * a + l), par);
    }
    case T_AA:
    case T_BB: { // sparseop.cpp:350-353 / :408-411
        const u64 d = rs.det[type == T_AA ? 0 : 1];
        const long koff = n3 * i + n2 * k;
        const double x = __ldg(P.two_mo + koff + n1 * a + l) - __ldg(P.two_mo + koff + n1 * l + a);
        return apply_sign(x, parity_double(d, (int)i, (int)k, (int)a, (int)l));
    }
    case T_SA: { // sparseop.cpp:303-315 (GenCI :459-466)
        const long ioff = n3 * i;
        double val1 = __ldg(P.one_mo + n1 * i + a);
        const int na = rs.nocc[0], nb = (KIND == PYCI_FULLCI) ? rs.nocc[1] : 0;
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
    case T_SB: { // sparseop.cpp:382-394
        const long ioff = n3 * i;
        double val1 = __ldg(P.one_mo + n1 * i + a);
        const int na = rs.nocc[0], nb = rs.nocc[1];
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
    default: // T_DIAG
        return (KIND == PYCI_DOCI) ? diag_doci(P, rs) : diag_twobody<KIND>(P, rs);
    }
}

constexpr int UNROLL = 4;

// ---- count pass ---------------------------------------------------------------------------------
template<int KIND, int KM>
__global__ void __launch_bounds__(256) count_kernel(BuildParams P, DetIndex<KM> index, int npairs_dim) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uchar2 *pairs = reinterpret_cast<uchar2 *>(smem_raw);
    __shared__ RowShared rs;
    __shared__ int warp_sums[8];
    fill_pairs(pairs, npairs_dim);
    const int nspin = (KIND == PYCI_FULLCI) ? 2 : 1;
    for (long r = blockIdx.x; r < P.nloc; r += gridDim.x) {
        const long row = P.row0 + r;
        __syncthreads();
        row_setup(rs, P, row, nspin);
        __syncthreads();
        int cnt = 0;
        for (u32 base = 0; base < P.ncand; base += UNROLL * blockDim.x) {
            int hit[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const u32 c = base + u * blockDim.x + threadIdx.x;
                hit[u] = -1;
                if (c < P.ncand) {
                    u64 A, B;
                    u32 code;
                    decode<KIND>(P, rs, pairs, c, A, B, code);
                    hit[u] = index.find(A, B);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                cnt += (hit[u] >= 0 && hit[u] < P.ncol);
        }
        // block reduce
        for (int o = 16; o > 0; o >>= 1)
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if ((threadIdx.x & 31) == 0)
            warp_sums[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = (row < P.ncol) ? 1 : 0; // diagonal, sparseop.cpp:252-255
            for (int wq = 0; wq < (int)((blockDim.x + 31) >> 5); ++wq)
                tot += warp_sums[wq];
            P.rowcnt[r] = tot;
        }
    }
}

// ---- fill pass ----------------------------------------------------------------------------------

// sort buf[0,m) ascending; normalised bitonic network (all comparators ascending), virtual +inf padding
__device__ void block_sort(u64 *buf, int m) {
    int P2 = 1;
    while (P2 < m)
        P2 <<= 1;
    const int half = P2 >> 1;
    for (int k = 2; k <= P2; k <<= 1) {
        // flip step
        for (int t = threadIdx.x; t < half; t += blockDim.x) {
            const int hk = k >> 1;
            const int blk = t / hk, off = t - blk * hk;
            const int i = blk * k + off, j = blk * k + (k - 1 - off);
            if (j < m) {
                const u64 x = buf[i], y = buf[j];
                if (x > y) {
                    buf[i] = y;
                    buf[j] = x;
                }
            }
        }
        __syncthreads();
        for (int js = k >> 2; js >= 1; js >>= 1) {
            for (int t = threadIdx.x; t < half; t += blockDim.x) {
                const int i = 2 * js * (t / js) + (t % js), j = i + js;
                if (j < m) {
                    const u64 x = buf[i], y = buf[j];
                    if (x > y) {
                        buf[i] = y;
                        buf[j] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
}

template<int KIND, int KM>
__global__ void __launch_bounds__(256) fill_kernel(BuildParams P, DetIndex<KM> index, int npairs_dim) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);
    uchar2 *pairs = reinterpret_cast<uchar2 *>(buf + P.maxrow);
    __shared__ RowShared rs;
    fill_pairs(pairs, npairs_dim);
    const int nspin = (KIND == PYCI_FULLCI) ? 2 : 1;
    const int lane = threadIdx.x & 31;
    const u32 lt = (1u << lane) - 1u;
    for (long r = blockIdx.x; r < P.nloc; r += gridDim.x) {
        const long row = P.row0 + r;
        __syncthreads();
        row_setup(rs, P, row, nspin);
        __syncthreads();
        if (threadIdx.x == 0 && row < P.ncol)
            buf[atomicAdd(&rs.count, 1)] = ((u64)row << 32) | pack_code(T_DIAG, 0, 0, 0, 0);
        for (u32 base = 0; base < P.ncand; base += UNROLL * blockDim.x) {
            int hit[UNROLL];
            u32 codes[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const u32 c = base + u * blockDim.x + threadIdx.x;
                hit[u] = -1;
                codes[u] = 0;
                if (c < P.ncand) {
                    u64 A, B;
                    decode<KIND>(P, rs, pairs, c, A, B, codes[u]);
                    hit[u] = index.find(A, B);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                // warp-aggregated append: one shared atomic per warp per step
                const bool keep = hit[u] >= 0 && hit[u] < P.ncol;
                const u32 m = __ballot_sync(0xffffffffu, keep);
                if (m) {
                    int slot = 0;
                    const int leader = __ffs(m) - 1;
                    if (lane == leader)
                        slot = atomicAdd(&rs.count, __popc(m));
                    slot = __shfl_sync(0xffffffffu, slot, leader) + __popc(m & lt);
                    if (keep)
                        buf[slot] = ((u64)(u32)hit[u] << 32) | codes[u];
                }
            }
        }
        __syncthreads();
        const int m = rs.count;
        block_sort(buf, m);
        // evaluate and stream out
        const long out0 = P.indptr[r];
        for (int e = threadIdx.x; e < m; e += blockDim.x) {
            const u64 kv = buf[e];
            const long col = (long)(kv >> 32);
            const u32 code = (u32)kv;
            const double val = element<KIND>(P, rs, code);
            P.cols[out0 + e] = (int)col;
            P.vals[out0 + e] = val;
            if ((code >> 24) == T_DIAG)
                P.diag[r] = val;
            // number of entries with col <= row: rows are sorted, so it is a prefix length
            const bool le = col <= row;
            const bool next_gt = (e + 1 == m) || ((long)(buf[e + 1] >> 32) > row);
            if (le && next_gt)
                P.lowcnt[r] = e + 1;
        }
    }
}

// ---- int64 exclusive scan of the row counts -------------------------------------------------------
constexpr int SCAN_BLOCK = 1024;

__global__ void scan_block_sums(const int *cnt, long n, long *blocksum) {
    __shared__ long ws[32];
    const long i = (long)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    long v = (i < n) ? cnt[i] : 0;
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0)
        ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        long t = ws[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1)
            t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0)
            blocksum[blockIdx.x] = t;
    }
}

// single block: exclusive scan of blocksum[nb] in place, total -> blocksum[nb]
__global__ void scan_of_sums(long *blocksum, long nb) {
    __shared__ long ws[32];
    __shared__ long carry;
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    for (long base = 0; base < nb; base += SCAN_BLOCK) {
        const long i = base + threadIdx.x;
        const long v = (i < nb) ? blocksum[i] : 0;
        long x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const long t = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o)
                x += t;
        }
        if ((threadIdx.x & 31) == 31)
            ws[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            long t = ws[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                const long q = __shfl_up_sync(0xffffffffu, t, o);
                if (threadIdx.x >= o)
                    t += q;
            }
            ws[threadIdx.x] = t;
        }
        __syncthreads();
        const long woff = (threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0;
        const long incl = x + woff + carry;
        if (i < nb)
            blocksum[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1)
            carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        blocksum[nb] = carry;
}

__global__ void scan_finish(const int *cnt, long n, const long *blocksum, long *indptr, int *maxcnt) {
    __shared__ long ws[32];
    const long i = (long)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const long v = (i < n) ? cnt[i] : 0;
    long x = v;
    for (int o = 1; o < 32; o <<= 1) {
        const long t = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o)
            x += t;
    }
    if ((threadIdx.x & 31) == 31)
        ws[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        long t = ws[threadIdx.x];
        for (int o = 1; o < 32; o <<= 1) {
            const long q = __shfl_up_sync(0xffffffffu, t, o);
            if (threadIdx.x >= o)
                t += q;
        }
        ws[threadIdx.x] = t;
    }
    __syncthreads();
    const long woff = (threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0;
    if (i < n) {
        indptr[i] = blocksum[blockIdx.x] + woff + x - v;
        if (i == n - 1)
            indptr[n] = blocksum[blockIdx.x] + woff + x;
    }
    int mx = (int)v;
    for (int o = 16; o > 0; o >>= 1)
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0)
        atomicMax(maxcnt, mx);
}

__global__ void uniform_counts(int *cnt, long n, int value) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        cnt[i] = value;
}

// ---- hash index build -------------------------------------------------------------------------------

template<int KM>
__global__ void insert_kernel(typename SlotOf<KM>::type *slots, u32 mask, int shift, const u64 *dets,
                              int nwords, long ndet) {
    const long idet = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idet >= ndet)
        return;
    const u64 a = dets[idet * nwords], b = (nwords == 2) ? dets[idet * nwords + 1] : 0ULL;
    DetIndex<KM> ix;
    ix.slots = slots;
    ix.mask = mask;
    ix.shift = shift;
    u32 p;
    if constexpr (KM == KEY128)
        p = ix.home(a, b);
    else
        p = ix.home(ix.key(a, b));
    for (;;) {
        // claim an empty slot (val == -1); keys are written afterwards -- determinants are unique, so
        // no key comparison is needed while inserting (verified by verify_kernel)
        if (atomicCAS(&slots[p].val, -1, (int)idet) == -1) {
            if constexpr (KM == KEY32)
                slots[p].key = ix.key(a, b);
            else if constexpr (KM == KEY64)
                slots[p].key = ix.key(a, b);
            else {
                slots[p].k0 = a;
                slots[p].k1 = b;
            }
            return;
        }
        p = (p + 1) & mask;
    }
}

template<int KM>
__global__ void verify_kernel(DetIndex<KM> ix, const u64 *dets, int nwords, long ndet, int *bad) {
    const long idet = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idet >= ndet)
        return;
    const u64 a = dets[idet * nwords], b = (nwords == 2) ? dets[idet * nwords + 1] : 0ULL;
    if (ix.find(a, b) != (int)idet)
        atomicAdd(bad, 1);
}

template<int KM>
__global__ void lookup_kernel(DetIndex<KM> ix, const u64 *dets, int nwords, long n, long *out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const u64 a = dets[i * nwords], b = (nwords == 2) ? dets[i * nwords + 1] : 0ULL;
    out[i] = ix.find(a, b);
}

size_t slot_bytes(int km) { return km == KEY32 ? sizeof(Slot32) : km == KEY64 ? sizeof(Slot64) : sizeof(Slot128); }

template<int KM>
int build_index_t(pyci_wfn *wfn) {
    pyci_ctx *ctx = wfn->ctx;
    typedef typename SlotOf<KM>::type slot_t;
    const long ndet = wfn->ndet;
    const int threads = 256;
    const unsigned blocks = (unsigned)((ndet + threads - 1) / threads);
    if (ndet > 0) {
        insert_kernel<KM><<<blocks, threads, 0, ctx->stream>>>(reinterpret_cast<slot_t *>(wfn->slots), wfn->mask,
                                                             (wfn->kind == PYCI_FULLCI) ? (int)wfn->nbasis : 0,
                                                             wfn->dets, wfn->nwords, ndet);
        ctx->launches++;
        int *bad = nullptr;
        PYCI_CUDA(cudaMalloc(&bad, sizeof(int)));
        PYCI_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
        verify_kernel<KM><<<blocks, threads, 0, ctx->stream>>>(make_index<KM>(wfn), wfn->dets, wfn->nwords, ndet, bad);
        ctx->launches++;
        int hbad = 0;
        PYCI_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(bad);
        if (hbad)
            PYCI_FAIL(PYCI_ERR_VALUE, "wave function contains %d duplicate determinant(s)", hbad);
    }
    return PYCI_OK;
}

double binom_d(long n, long k) {
    if (k < 0 || k > n)
        return 0.0;
    double b = 1.0;
    for (long d = 1; d <= k; ++d)
        b = b * (double)(n - k + d) / (double)d;
    return b;
}

template<int KIND, int KM>
int run_build(pyci_ctx *ctx, const pyci_wfn *wfn, pyci_op *op, BuildParams &P, int npairs_dim) {
    cudaStream_t st = ctx->stream;
    const DetIndex<KM> ix = make_index<KM>(wfn);
    const long nloc = op->nloc;
    const size_t pair_bytes = pair_table_bytes(P);

    // block size from the amount of per-row work
    auto pick_block = [](long work) { return work <= 256 ? 32 : work <= 1024 ? 64 : work <= 4096 ? 128 : 256; };

    PYCI_CUDA(cudaEventRecord(ctx->ev[0], st));
    int *rowcnt = nullptr;
    PYCI_CUDA(cudaMalloc(&rowcnt, sizeof(int) * (size_t)(nloc + 1)));
    P.rowcnt = rowcnt;
    const bool analytic = wfn->complete && op->ncol == wfn->ndet;
    if (nloc > 0) {
        if (analytic) {
            // complete space: every excitation is in the wfn, so each row holds ncand + 1 entries
            uniform_counts<<<(unsigned)((nloc + 255) / 256), 256, 0, st>>>(rowcnt, nloc, (int)P.ncand + 1);
            ctx->launches++;
        } else {
            const int block = pick_block((long)P.ncand / 4);
            int per_sm = 1;
            PYCI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, count_kernel<KIND, KM>, block, pair_bytes));
            const long grid = std::min<long>(nloc, (long)ctx->sm_count * std::max(per_sm, 1));
            count_kernel<KIND, KM><<<(unsigned)grid, block, pair_bytes, st>>>(P, ix, npairs_dim);
            ctx->launches++;
        }
    }
    // scan
    const long nb = (nloc + SCAN_BLOCK - 1) / SCAN_BLOCK;
    long *blocksum = nullptr;
    int *maxcnt = nullptr;
    PYCI_CUDA(cudaMalloc(&blocksum, sizeof(long) * (size_t)(nb + 2)));
    PYCI_CUDA(cudaMalloc(&maxcnt, sizeof(int)));
    PYCI_CUDA(cudaMemsetAsync(maxcnt, 0, sizeof(int), st));
    PYCI_CUDA(cudaMemsetAsync(op->indptr, 0, sizeof(long) * (size_t)(nloc + 1), st));
    if (nloc > 0) {
        scan_block_sums<<<(unsigned)nb, SCAN_BLOCK, 0, st>>>(rowcnt, nloc, blocksum);
        scan_of_sums<<<1, SCAN_BLOCK, 0, st>>>(blocksum, nb);
        scan_finish<<<(unsigned)nb, SCAN_BLOCK, 0, st>>>(rowcnt, nloc, blocksum, op->indptr, maxcnt);
        ctx->launches += 3;
    }
    long nnz = 0;
    int maxrow = 0;
    PYCI_CUDA(cudaMemcpyAsync(&nnz, op->indptr + nloc, sizeof(long), cudaMemcpyDeviceToHost, st));
    PYCI_CUDA(cudaMemcpyAsync(&maxrow, maxcnt, sizeof(int), cudaMemcpyDeviceToHost, st));
    PYCI_CUDA(cudaEventRecord(ctx->ev[1], st));
    PYCI_CUDA(cudaStreamSynchronize(st));
    cudaFree(blocksum);
    cudaFree(maxcnt);
    cudaFree(rowcnt);
    P.rowcnt = nullptr;

    op->nnz = nnz;
    PYCI_CUDA(cudaMalloc(&op->cols, sizeof(int) * (size_t)std::max<long>(nnz, 1)));
    PYCI_CUDA(cudaMalloc(&op->vals, sizeof(double) * (size_t)std::max<long>(nnz, 1)));
    P.cols = op->cols;
    P.vals = op->vals;
    P.maxrow = (maxrow + 1) & ~1;
    const size_t smem = sizeof(u64) * (size_t)P.maxrow + pair_bytes;
    if ((long)smem > (long)ctx->smem_optin)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED,
                  "a matrix row holds %d entries; rows above %ld entries do not fit the shared-memory row buffer",
                  maxrow, (long)((ctx->smem_optin - pair_bytes) / sizeof(u64)));
    if (nloc > 0 && nnz > 0) {
        const int block = pick_block(std::max<long>((long)P.ncand / 4, maxrow));
        PYCI_CUDA(cudaFuncSetAttribute(fill_kernel<KIND, KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        PYCI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_kernel<KIND, KM>, block, smem));
        const long grid = std::min<long>(nloc, (long)ctx->sm_count * std::max(per_sm, 1));
        fill_kernel<KIND, KM><<<(unsigned)grid, block, smem, st>>>(P, ix, npairs_dim);
        ctx->launches++;
    }
    PYCI_CUDA(cudaEventRecord(ctx->ev[2], st));
    PYCI_CUDA(cudaStreamSynchronize(st));
    PYCI_CUDA(cudaGetLastError());
    float ms01 = 0, ms12 = 0;
    cudaEventElapsedTime(&ms01, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ms12, ctx->ev[1], ctx->ev[2]);
    op->times[0] = wfn->hash_seconds;
    op->times[1] = ms01 * 1e-3;
    op->times[2] = ms12 * 1e-3;
    op->times[3] = op->times[1] + op->times[2];
    return PYCI_OK;
}

template<int KIND>
int dispatch_key(pyci_ctx *ctx, const pyci_wfn *wfn, pyci_op *op, BuildParams &P, int npairs_dim) {
    switch (wfn->keymode) {
    case KEY32:
        return run_build<KIND, KEY32>(ctx, wfn, op, P, npairs_dim);
    case KEY64:
        return run_build<KIND, KEY64>(ctx, wfn, op, P, npairs_dim);
    default:
        if constexpr (KIND == PYCI_FULLCI)
            return run_build<KIND, KEY128>(ctx, wfn, op, P, npairs_dim);
        else
            PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "one-spin wave functions use 32- or 64-bit keys");
    }
}

__global__ void lowcnt_sum_kernel(const int *lowcnt, long n, unsigned long long *out) {
    long acc = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        acc += lowcnt[i];
    for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc)
        atomicAdd(out, (unsigned long long)acc);
}

} // namespace

int wfn_build_index(pyci_wfn *wfn) {
    pyci_ctx *ctx = wfn->ctx;
    // capacity: power of two with load factor in (0.25, 0.5]
    u64 cap = 16;
    while (cap < 2 * (u64)wfn->ndet)
        cap <<= 1;
    if (cap > (1ULL << 31))
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "too many determinants for the device index (%ld)", wfn->ndet);
    wfn->mask = (u32)(cap - 1);
    const size_t bytes = slot_bytes(wfn->keymode) * (size_t)cap;
    PYCI_CUDA(cudaMalloc(&wfn->slots, bytes));
    PYCI_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    PYCI_CUDA(cudaMemsetAsync(wfn->slots, 0xFF, bytes, ctx->stream));
    int rc;
    switch (wfn->keymode) {
    case KEY32:
        rc = build_index_t<KEY32>(wfn);
        break;
    case KEY64:
        rc = build_index_t<KEY64>(wfn);
        break;
    default:
        rc = build_index_t<KEY128>(wfn);
        break;
    }
    PYCI_TRY(rc);
    PYCI_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    wfn->hash_seconds = ms * 1e-3;
    return PYCI_OK;
}

int wfn_index_dets_impl(pyci_wfn *wfn, long n, const u64 *dets_dev, long *out_dev) {
    pyci_ctx *ctx = wfn->ctx;
    if (n <= 0)
        return PYCI_OK;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    switch (wfn->keymode) {
    case KEY32:
        lookup_kernel<KEY32><<<blocks, 256, 0, ctx->stream>>>(make_index<KEY32>(wfn), dets_dev, wfn->nwords, n, out_dev);
        break;
    case KEY64:
        lookup_kernel<KEY64><<<blocks, 256, 0, ctx->stream>>>(make_index<KEY64>(wfn), dets_dev, wfn->nwords, n, out_dev);
        break;
    default:
        lookup_kernel<KEY128><<<blocks, 256, 0, ctx->stream>>>(make_index<KEY128>(wfn), dets_dev, wfn->nwords, n, out_dev);
        break;
    }
    ctx->launches++;
    PYCI_CUDA(cudaGetLastError());
    return PYCI_OK;
}

int op_build_impl(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, pyci_op *op) {
    BuildParams P;
    PYCI_TRY(enum_params_init(P, wfn));
    const int kind = wfn->kind;
    P.row0 = op->row0;
    P.nloc = op->nloc;
    P.ncol = op->ncol;
    P.one_mo = ham->one_mo;
    P.two_mo = ham->two_mo;
    P.h = ham->h;
    P.v = ham->v;
    P.w = ham->w;
    P.indptr = op->indptr;
    P.lowcnt = op->lowcnt;
    P.diag = op->diag;
    const int npairs_dim = P.npairs_dim;

    int rc;
    if (kind == PYCI_DOCI)
        rc = dispatch_key<PYCI_DOCI>(ctx, wfn, op, P, npairs_dim);
    else if (kind == PYCI_FULLCI)
        rc = dispatch_key<PYCI_FULLCI>(ctx, wfn, op, P, npairs_dim);
    else
        rc = dispatch_key<PYCI_GENCI>(ctx, wfn, op, P, npairs_dim);
    PYCI_TRY(rc);

    // SparseOp::size in the reference's storage
    if (op->symmetric) {
        unsigned long long *acc = nullptr;
        PYCI_CUDA(cudaMalloc(&acc, sizeof(unsigned long long)));
        PYCI_CUDA(cudaMemsetAsync(acc, 0, sizeof(unsigned long long), ctx->stream));
        if (op->nloc > 0) {
            lowcnt_sum_kernel<<<ctx->sm_count, 256, 0, ctx->stream>>>(op->lowcnt, op->nloc, acc);
            ctx->launches++;
        }
        unsigned long long h = 0;
        PYCI_CUDA(cudaMemcpyAsync(&h, acc, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(acc);
        op->size_ref = (long)h;
    } else {
        op->size_ref = op->nnz;
    }
    return PYCI_OK;
}
