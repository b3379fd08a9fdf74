// Heat-bath CI selection and Epstein-Nesbet PT2 on the device: the B200 form of
// add_hci (/root/reference/pyci/src/hci.cpp:22-279) and compute_enpt2 (enpt2.cpp:21-400).
//
// Both walk the excitations of every determinant with the enumerator of the Hamiltonian construction
// (enumerate.cuh / elements.cuh: same candidates, same element arithmetic and summation order) and keep a
// candidate j of determinant i when |H_ji| > eps / |c_i| and j is NOT in the wave function.  The external
// determinants are collected in a second open-addressing table in HBM keyed by the bit-strings:
//
//   add_hci       payload = smallest (row << 24 | position in the reference's loop nest) over all i that
//                 reach j (atomicMin); compaction + radix sort on that payload gives every new determinant
//                 once, in first-encounter order of the reference's serial loop (the reference itself appends
//                 in the iteration order of its hash map, which is unspecified: the SET is the contract).
//   compute_enpt2 payload = sum_i H_ji c_i (fp64 atomicAdd); a second pass evaluates H_jj of every external
//                 determinant (enpt2.cpp:36-66, 246-255) and reduces sum_j payload^2 / (E - ecore - H_jj).
//
// The table grows by re-running the pass when it fills (the load is not known before the walk).  Row-sharded
// over ranks: every rank walks its own rows, the compacted external lists are all-gathered and merged.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>

#include "elements.cuh"
#include "radix.cuh"

namespace {

constexpr u64 EXT_EMPTY = ~0ULL;
enum { MODE_HCI = 0, MODE_PT2 = 1 };

struct ExtTable {
    u64 *k0, *k1, *pay; // [cap] key words (k1 only for two-spin kinds) and payload
    u32 mask;
    unsigned long long *count; // occupied slots
    unsigned long long limit;  // give up (and let the host grow the table) beyond this many
    int *overflow;
};

__device__ __forceinline__ u32 ext_home(u64 a, u64 b) { return mix64(a ^ (b * 0x9e3779b97f4a7c15ULL) ^ (b >> 29)); }

// find-or-insert; returns the slot or -1 when the table is (nearly) full
template<bool TWO>
__device__ __forceinline__ long ext_slot(const ExtTable &T, u64 a, u64 b) {
    u32 p = ext_home(a, TWO ? b : 0ULL) & T.mask;
    for (int probes = 0; probes < 4096; ++probes) {
        u64 cur = *reinterpret_cast<volatile u64 *>(T.k0 + p);
        if (cur == EXT_EMPTY) {
            cur = atomicCAS(reinterpret_cast<unsigned long long *>(T.k0 + p), EXT_EMPTY, a);
            if (cur == EXT_EMPTY) {
                if (TWO) {
                    *reinterpret_cast<volatile u64 *>(T.k1 + p) = b;
                    __threadfence();
                }
                if (atomicAdd(T.count, 1ULL) + 1ULL > T.limit)
                    *T.overflow = 1;
                return (long)p;
            }
        }
        if (cur == a) {
            if (!TWO)
                return (long)p;
            u64 kb;
            while ((kb = *reinterpret_cast<volatile u64 *>(T.k1 + p)) == EXT_EMPTY) {
            }
            if (kb == b)
                return (long)p;
        }
        p = (p + 1) & T.mask;
    }
    *T.overflow = 1;
    return -1;
}

// position of candidate c in the reference's loop nest (hci.cpp:69-187, 201-236, 32-46): monotone key, not dense
struct OrderParams {
    u32 Ma, Mb, offB, nva, nvb, nSb;
};

template<int KIND>
__device__ __forceinline__ u32 order_key(const BuildParams &P, const OrderParams &O, const uchar2 *__restrict__ pairs,
                                         u32 c) {
    if (KIND == PYCI_DOCI)
        return c; // occ-major, vir-minor
    if (KIND == PYCI_FULLCI) {
        if (c < P.nAB) { // (i,a) alpha single, then beta (k,l)
            const u32 sa = fdiv(c, P.dSb), sb = c - sa * P.nSb;
            return sa * O.Ma + 1u + sb;
        }
        c -= P.nAB;
    }
    if (c < P.nDa) { // (i,a) then (k>i, l>a) of the same spin, after the alpha-beta block of (i,a)
        const u32 po = fdiv(c, P.dPva), pv = c - po * P.nPva;
        const uchar2 o = pairs[po], v = pairs[pv];
        return ((u32)o.x * O.nva + v.x) * O.Ma + 1u + O.nSb + (u32)o.y * O.nva + v.y;
    }
    c -= P.nDa;
    if (KIND == PYCI_FULLCI) {
        if (c < P.nDb) {
            const u32 po = fdiv(c, P.dPvb), pv = c - po * P.nPvb;
            const uchar2 o = pairs[po], v = pairs[pv];
            return O.offB + ((u32)o.x * O.nvb + v.x) * O.Mb + 1u + (u32)o.y * O.nvb + v.y;
        }
        c -= P.nDb;
    }
    if (c < P.nSa)
        return c * O.Ma;
    c -= P.nSa;
    return O.offB + c * O.Mb;
}

constexpr int EXT_UNROLL = 4;

// one CTA per determinant row (persistent grid); threads split the row's excitation candidates
template<int KIND, int KM, int MODE>
__global__ void __launch_bounds__(256) ext_walk_kernel(BuildParams P, DetIndex<KM> index, u32 nSa, u32 nSb, ExtTable T,
                                                       double eps, OrderParams O) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const RowTables RT = carve_tables(smem_raw, nSa, nSb);
    uchar2 *pairs = reinterpret_cast<uchar2 *>(smem_raw + tables_bytes(nSa, nSb));
    __shared__ RowShared rs;
    __shared__ int s_over;
    pair_masks_none(rs);
    fill_pairs(pairs, P.npairs_dim);
    constexpr int nspin = (KIND == PYCI_FULLCI) ? 2 : 1;
    constexpr bool TWO = (KIND == PYCI_FULLCI);
    for (long r = blockIdx.x; r < P.nloc; r += gridDim.x) {
        const long row = P.row0 + r;
        __syncthreads();
        if (threadIdx.x == 0)
            s_over = *reinterpret_cast<volatile int *>(T.overflow);
        row_setup(rs, P, row, nspin);
        __syncthreads();
        if (s_over)
            return; // table full: the host re-runs the pass with a larger one (uniform exit)
        if (KIND != PYCI_DOCI) {
            build_tables<KIND, true>(P, rs, RT, nSa, nSb);
            __syncthreads();
        }
        const double ci = __ldg(P.coeffs + row);
        const double eps_i = eps / fabs(ci); // hci.cpp:26,56,193 (inf when c_i == 0: nothing passes)
        for (u32 base = 0; base < P.ncand; base += EXT_UNROLL * blockDim.x) {
            u64 A[EXT_UNROLL], B[EXT_UNROLL];
            double val[EXT_UNROLL];
            int hit[EXT_UNROLL];
            bool want[EXT_UNROLL];
#pragma unroll
            for (int u = 0; u < EXT_UNROLL; ++u) {
                const u32 c = base + u * blockDim.x + threadIdx.x;
                want[u] = false;
                val[u] = 0.0;
                A[u] = B[u] = 0ULL;
                if (c < P.ncand) {
                    candidate<KIND, true>(P, rs, RT, pairs, c, A[u], B[u], val[u]);
                    want[u] = fabs(val[u]) > eps_i;
                }
            }
            find_batch<EXT_UNROLL>(index, A, B, want, hit);
#pragma unroll
            for (int u = 0; u < EXT_UNROLL; ++u) {
                if (!want[u] || hit[u] >= 0)
                    continue;
                const long s = ext_slot<TWO>(T, A[u], B[u]);
                if (s < 0)
                    continue;
                if (MODE == MODE_HCI) {
                    const u32 c = base + u * blockDim.x + threadIdx.x;
                    atomicMin(reinterpret_cast<unsigned long long *>(T.pay + s),
                              ((unsigned long long)row << 24) | order_key<KIND>(P, O, pairs, c));
                } else {
                    atomicAdd(reinterpret_cast<double *>(T.pay + s), val[u] * ci); // enpt2.cpp:124,144,165,...
                }
            }
        }
    }
}

// occupied slots -> dense list (payload, k0, k1), warp-aggregated
__global__ void ext_compact_kernel(ExtTable T, bool two, u64 *out_pay, u64 *out_k0, u64 *out_k1,
                                   unsigned long long *cursor) {
    const long cap = (long)T.mask + 1;
    const int lane = threadIdx.x & 31;
    for (long base = (long)blockIdx.x * blockDim.x; base < cap; base += (long)gridDim.x * blockDim.x) {
        const long s = base + threadIdx.x;
        const bool occ = s < cap && T.k0[s] != EXT_EMPTY;
        const u32 m = __ballot_sync(0xffffffffu, occ);
        if (!m)
            continue;
        unsigned long long start = 0;
        if (lane == 0)
            start = atomicAdd(cursor, (unsigned long long)__popc(m));
        start = __shfl_sync(0xffffffffu, start, 0);
        if (occ) {
            const unsigned long long d = start + __popc(m & ((1u << lane) - 1u));
            out_pay[d] = T.pay[s];
            out_k0[d] = T.k0[s];
            if (two)
                out_k1[d] = T.k1[s];
        }
    }
}

// merge lists gathered from all ranks into one table
template<bool TWO, int MODE>
__global__ void ext_merge_kernel(ExtTable T, const u64 *pay, const u64 *k0, const u64 *k1, long stride, int nranks,
                                 const long *counts) {
    for (int rk = 0; rk < nranks; ++rk) {
        const long cnt = counts[rk];
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (long)gridDim.x * blockDim.x) {
            const long q = rk * stride + i;
            const long s = ext_slot<TWO>(T, k0[q], TWO ? k1[q] : 0ULL);
            if (s < 0)
                continue;
            if (MODE == MODE_HCI)
                atomicMin(reinterpret_cast<unsigned long long *>(T.pay + s), (unsigned long long)pay[q]);
            else
                atomicAdd(reinterpret_cast<double *>(T.pay + s), __longlong_as_double((long long)pay[q]));
        }
    }
}

// rank that owns an external determinant: the same rule pt2_reduce_kernel uses for its share
__device__ __forceinline__ u32 ext_owner(u64 a, u64 b, u32 nranks) { return (ext_home(a, b) >> 7) % nranks; }

template<bool TWO>
__global__ void owner_count_kernel(const u64 *k0, const u64 *k1, long n, u32 nranks, unsigned long long *cnt) {
    __shared__ unsigned int h[64];
    if (threadIdx.x < 64)
        h[threadIdx.x] = 0;
    __syncthreads();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        atomicAdd(&h[ext_owner(k0[i], TWO ? k1[i] : 0ULL, nranks)], 1u);
    __syncthreads();
    if (threadIdx.x < nranks && h[threadIdx.x])
        atomicAdd(cnt + threadIdx.x, (unsigned long long)h[threadIdx.x]);
}

// entries grouped by owner: cursor[p] starts at the first slot of owner p's segment
template<bool TWO>
__global__ void owner_scatter_kernel(const u64 *pay, const u64 *k0, const u64 *k1, long n, u32 nranks,
                                     unsigned long long *cursor, u64 *spay, u64 *s0, u64 *s1) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const u64 a = k0[i], b = TWO ? k1[i] : 0ULL;
        const unsigned long long at = atomicAdd(cursor + ext_owner(a, b, nranks), 1ULL);
        spay[at] = pay[i];
        s0[at] = a;
        if (TWO)
            s1[at] = b;
    }
}

__global__ void ext_gather_dets_kernel(const u32 *order, const u64 *k0, const u64 *k1, long n, int nwords, u64 *out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const u32 s = order[i];
    out[i * nwords] = k0[s];
    if (nwords == 2)
        out[i * nwords + 1] = k1[s];
}

__global__ void iota_kernel(u32 *p, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        p[i] = (u32)i;
}

// Sum over this rank's share of the external determinants of (sum_i H_ji c_i)^2 / (e0 - H_jj).  Every rank holds the
// whole merged list but in its own order (slot placement depends on insertion races), so the share is defined by the
// determinant itself: rank = hash(j) mod nranks.
template<int KIND>
__global__ void __launch_bounds__(256) pt2_reduce_kernel(BuildParams P, const u64 *pay, const u64 *k0, const u64 *k1,
                                                         long n, u32 rank, u32 nranks, double e0, double *out) {
    __shared__ double ws[8];
    double acc = 0.0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const u64 a = k0[i], b = (KIND == PYCI_FULLCI) ? k1[i] : 0ULL;
        if (nranks > 1 && (ext_home(a, b) >> 7) % nranks != rank)
            continue;
        const double s = __longlong_as_double((long long)pay[i]);
        const double diag = diag_twobody(P, a, b);
        acc += s * s / (e0 - diag); // enpt2.cpp:370
    }
    for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0)
        ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
            t += ws[w];
        atomicAdd(out, t);
    }
}

// ---- host side ------------------------------------------------------------------------------------

struct ExtBuffers {
    ExtTable T{};
    long cap = 0;
    bool two = false;
    int rc_alloc(pyci_ctx *ctx, long capacity, bool two_, int mode) {
        two = two_;
        cap = capacity;
        T.mask = (u32)(capacity - 1);
        T.limit = (unsigned long long)(capacity * 0.6);
        PYCI_CUDA(dev_malloc(&T.k0, sizeof(u64) * (size_t)capacity));
        if (two)
            PYCI_CUDA(dev_malloc(&T.k1, sizeof(u64) * (size_t)capacity));
        PYCI_CUDA(dev_malloc(&T.pay, sizeof(u64) * (size_t)capacity));
        PYCI_CUDA(dev_malloc(&T.count, sizeof(unsigned long long)));
        PYCI_CUDA(dev_malloc(&T.overflow, sizeof(int)));
        PYCI_CUDA(cudaMemsetAsync(T.k0, 0xFF, sizeof(u64) * (size_t)capacity, ctx->stream));
        if (two)
            PYCI_CUDA(cudaMemsetAsync(T.k1, 0xFF, sizeof(u64) * (size_t)capacity, ctx->stream));
        // HCI: payload starts at the largest key (atomicMin); PT2: at +0.0 (atomicAdd)
        PYCI_CUDA(cudaMemsetAsync(T.pay, mode == MODE_HCI ? 0xFF : 0x00, sizeof(u64) * (size_t)capacity, ctx->stream));
        PYCI_CUDA(cudaMemsetAsync(T.count, 0, sizeof(unsigned long long), ctx->stream));
        PYCI_CUDA(cudaMemsetAsync(T.overflow, 0, sizeof(int), ctx->stream));
        return PYCI_OK;
    }
    void release() {
        dev_free(T.k0);
        dev_free(T.k1);
        dev_free(T.pay);
        dev_free(T.count);
        dev_free(T.overflow);
        T = ExtTable{};
    }
};

struct ExtList { // dense (payload, k0, k1) on the device
    u64 *pay = nullptr, *k0 = nullptr, *k1 = nullptr;
    long n = 0;
    void release() {
        dev_free(pay);
        dev_free(k0);
        dev_free(k1);
        pay = k0 = k1 = nullptr;
        n = 0;
    }
};

long next_pow2(long x) {
    long c = 1;
    while (c < x)
        c <<= 1;
    return c;
}

int compact(pyci_ctx *ctx, ExtBuffers &E, ExtList &L) {
    unsigned long long cnt = 0;
    PYCI_CUDA(cudaMemcpyAsync(&cnt, E.T.count, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream));
    PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
    L.n = (long)cnt;
    const size_t bytes = sizeof(u64) * (size_t)std::max<long>(L.n, 1);
    PYCI_CUDA(dev_malloc(&L.pay, bytes));
    PYCI_CUDA(dev_malloc(&L.k0, bytes));
    if (E.two)
        PYCI_CUDA(dev_malloc(&L.k1, bytes));
    if (L.n > 0) {
        unsigned long long *cursor = nullptr;
        PYCI_CUDA(dev_malloc(&cursor, sizeof(unsigned long long)));
        PYCI_CUDA(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), ctx->stream));
        const long blocks = std::min<long>((E.cap + 255) / 256, (long)ctx->sm_count * 8);
        ext_compact_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(E.T, E.two, L.pay, L.k0, L.k1, cursor);
        ctx->launches++;
        PYCI_CUDA(cudaGetLastError());
        dev_free(cursor);
    }
    return PYCI_OK;
}

template<int KIND, int KM, int MODE>
int walk_once(pyci_ctx *ctx, const pyci_wfn *wfn, BuildParams &P, const OrderParams &O, double eps, ExtBuffers &E,
              int *overflowed) {
    const DetIndex<KM> ix = make_index<KM>(wfn);
    const u32 nSa = (KIND == PYCI_DOCI) ? 0u : P.nSa;
    const u32 nSb = (KIND == PYCI_FULLCI) ? O.nSb : 0u;
    const size_t smem = tables_bytes(nSa, nSb) + pair_table_bytes(P);
    if ((long)smem > (long)ctx->smem_optin)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "excitation tables (%zu bytes) do not fit shared memory", smem);
    const long work = (long)P.ncand / 4;
    const int block = work <= 128 ? 32 : work <= 512 ? 64 : work <= 1024 ? 128 : 256;
    int per_sm = 1;
    PYCI_CUDA(cudaFuncSetAttribute(ext_walk_kernel<KIND, KM, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PYCI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ext_walk_kernel<KIND, KM, MODE>, block, smem));
    const long grid = std::min<long>(P.nloc, (long)ctx->sm_count * std::max(per_sm, 1));
    if (grid > 0) {
        ext_walk_kernel<KIND, KM, MODE><<<(unsigned)grid, block, smem, ctx->stream>>>(P, ix, nSa, nSb, E.T, eps, O);
        ctx->launches++;
    }
    PYCI_CUDA(cudaGetLastError());
    PYCI_CUDA(cudaMemcpyAsync(overflowed, E.T.overflow, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
    return PYCI_OK;
}

template<int KIND, int MODE>
int walk_key(pyci_ctx *ctx, const pyci_wfn *wfn, BuildParams &P, const OrderParams &O, double eps, ExtBuffers &E,
             int *overflowed) {
    switch (wfn->keymode) {
    case KEY32:
        return walk_once<KIND, KEY32, MODE>(ctx, wfn, P, O, eps, E, overflowed);
    case KEY64:
        return walk_once<KIND, KEY64, MODE>(ctx, wfn, P, O, eps, E, overflowed);
    default:
        if constexpr (KIND == PYCI_FULLCI)
            return walk_once<KIND, KEY128, MODE>(ctx, wfn, P, O, eps, E, overflowed);
        else
            PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "one-spin wave functions use 32- or 64-bit keys");
    }
}

template<int MODE>
int walk_kind(pyci_ctx *ctx, const pyci_wfn *wfn, BuildParams &P, const OrderParams &O, double eps, ExtBuffers &E,
              int *overflowed) {
    if (wfn->kind == PYCI_DOCI) {
        if constexpr (MODE == MODE_HCI)
            return walk_key<PYCI_DOCI, MODE>(ctx, wfn, P, O, eps, E, overflowed);
        else
            PYCI_FAIL(PYCI_ERR_VALUE, "compute_enpt2 of a DOCI wave function runs on its FullCI image (enpt2.cpp:376-380)");
    }
    if (wfn->kind == PYCI_FULLCI)
        return walk_key<PYCI_FULLCI, MODE>(ctx, wfn, P, O, eps, E, overflowed);
    return walk_key<PYCI_GENCI, MODE>(ctx, wfn, P, O, eps, E, overflowed);
}

// Merge `nlists` dense lists laid out as [nlists][stride] (counts[l] valid entries each) into one list without
// duplicate determinants: payloads are combined with atomicMin (HCI) or atomicAdd (PT2).
template<int MODE>
int merge_lists(pyci_ctx *ctx, bool two, const u64 *pay, const u64 *k0, const u64 *k1, long stride,
                const std::vector<long> &counts, ExtList &out) {
    long total = 0;
    for (long c : counts)
        total += c;
    const int nlists = (int)counts.size();
    const long mcap = std::max<long>(1L << 16, next_pow2(2 * std::max<long>(total, 1)));
    if (mcap > (1L << 31))
        PYCI_FAIL(PYCI_ERR_MEMORY, "external-space table would exceed 2^31 slots");
    long *dcounts = nullptr;
    ExtBuffers M;
    auto body = [&]() -> int {
        PYCI_CUDA(dev_malloc(&dcounts, sizeof(long) * (size_t)nlists));
        PYCI_CUDA(cudaMemcpyAsync(dcounts, counts.data(), sizeof(long) * (size_t)nlists, cudaMemcpyHostToDevice, ctx->stream));
        PYCI_TRY(M.rc_alloc(ctx, mcap, two, MODE));
        const unsigned blocks = (unsigned)std::max<long>(1, std::min<long>((stride + 255) / 256, (long)ctx->sm_count * 8));
        if (two)
            ext_merge_kernel<true, MODE><<<blocks, 256, 0, ctx->stream>>>(M.T, pay, k0, k1, stride, nlists, dcounts);
        else
            ext_merge_kernel<false, MODE><<<blocks, 256, 0, ctx->stream>>>(M.T, pay, k0, k1, stride, nlists, dcounts);
        ctx->launches++;
        PYCI_CUDA(cudaGetLastError());
        int over = 0;
        PYCI_CUDA(cudaMemcpyAsync(&over, M.T.overflow, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
        if (over)
            PYCI_FAIL(PYCI_ERR_RUNTIME, "external-space merge table overflowed");
        PYCI_TRY(compact(ctx, M, out));
        return PYCI_OK;
    };
    const int rc = body();
    dev_free(dcounts);
    M.release();
    if (rc != PYCI_OK)
        out.release();
    return rc;
}

// Walk rows [row0, row0 + nloc) of the wave function, growing the table until it holds every external
// determinant they reach, and compact it.
template<int MODE>
int walk_rows(pyci_ctx *ctx, const pyci_wfn *wfn, BuildParams &P, const OrderParams &O, double eps, long row0, long nloc,
              ExtList &out) {
    P.row0 = row0;
    P.nloc = nloc;
    const bool two = wfn->kind == PYCI_FULLCI;
    // first guess: 64 external determinants per row (a heat-bath step typically multiplies the space by 10-50), within
    // an eighth of the free memory; an overflowing pass is repeated with 4x the slots
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const long budget = std::max<long>(1L << 16, (long)(free_b / 8 / 24));
    // (a rank that walks 1/R of the rows still finds most of the external space -- the duplicates it no longer sees are
    // the ones the other ranks' rows would have produced -- so a row-sharded walk sizes its table for R times its rows:
    // a table that is too large costs a memset and a scan, one that is too small costs a whole second walk.  At 8 GPUs
    // the 100 376-determinant case of bench.py walked three times, 60 ms instead of 20.)
    const long rows_guess = std::min<long>(std::max<long>(wfn->ndet, 1), std::max<long>(nloc, 1) * std::max(ctx->nranks, 1));
    long cap = std::max<long>(1L << 16, next_pow2(std::min<long>(64 * rows_guess, budget)));
    // the size that held the last walk of this many rows (compute_enpt2 followed by add_hci, repeated selection steps)
    if (ctx->ext_hint_rows == nloc && ctx->ext_hint_cap > cap)
        cap = ctx->ext_hint_cap;
    ExtBuffers E;
    for (;;) {
        if (cap > (1L << 31))
            PYCI_FAIL(PYCI_ERR_MEMORY, "external-space table would exceed 2^31 slots");
        int rc = E.rc_alloc(ctx, cap, two, MODE);
        int over = 0;
        if (rc == PYCI_OK)
            rc = walk_kind<MODE>(ctx, wfn, P, O, eps, E, &over);
        if (rc != PYCI_OK) {
            E.release();
            return rc;
        }
        if (!over)
            break;
        E.release();
        cap *= 4;
    }
    ctx->ext_hint_rows = nloc;
    ctx->ext_hint_cap = cap;
    const int rc = compact(ctx, E, out);
    E.release();
    if (rc != PYCI_OK)
        out.release();
    return rc;
}

// pad lists to a common stride in one [nlists][stride] allocation per word
int pack_lists(pyci_ctx *ctx, bool two, const std::vector<ExtList> &lists, long stride, u64 **pay, u64 **k0, u64 **k1) {
    const size_t bytes = sizeof(u64) * (size_t)stride * lists.size();
    PYCI_CUDA(dev_malloc(pay, bytes));
    PYCI_CUDA(dev_malloc(k0, bytes));
    if (two)
        PYCI_CUDA(dev_malloc(k1, bytes));
    for (size_t l = 0; l < lists.size(); ++l) {
        const size_t nb = sizeof(u64) * (size_t)lists[l].n;
        PYCI_CUDA(cudaMemcpyAsync(*pay + l * stride, lists[l].pay, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        PYCI_CUDA(cudaMemcpyAsync(*k0 + l * stride, lists[l].k0, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        if (two)
            PYCI_CUDA(cudaMemcpyAsync(*k1 + l * stride, lists[l].k1, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return PYCI_OK;
}

// Owner-computes merge over ranks: every entry of this rank's list goes to the rank that owns its determinant
// (hash of the string), which merges what it receives -- duplicates from different ranks combine there (atomicMin /
// atomicAdd as in the local table).  Each rank ends with the duplicate-free list of the determinants it owns: the merge
// work is divided by the rank count instead of repeated on every rank.
template<int MODE>
int exchange_to_owners(pyci_ctx *ctx, bool two, ExtList &L, ExtList &owned) {
    const int R = ctx->nranks, me = ctx->rank;
    cudaStream_t st = ctx->stream;
    unsigned long long *dcnt = nullptr;
    u64 *spay = nullptr, *s0 = nullptr, *s1 = nullptr, *rpay = nullptr, *r0 = nullptr, *r1 = nullptr;
    auto body = [&]() -> int {
        PYCI_CUDA(dev_malloc(&dcnt, sizeof(unsigned long long) * 2 * (size_t)R));
        PYCI_CUDA(cudaMemsetAsync(dcnt, 0, sizeof(unsigned long long) * 2 * (size_t)R, st));
        const unsigned blocks = (unsigned)std::max<long>(1, std::min<long>((L.n + 255) / 256, (long)ctx->sm_count * 8));
        if (L.n > 0) {
            if (two)
                owner_count_kernel<true><<<blocks, 256, 0, st>>>(L.k0, L.k1, L.n, (u32)R, dcnt);
            else
                owner_count_kernel<false><<<blocks, 256, 0, st>>>(L.k0, L.k1, L.n, (u32)R, dcnt);
            ctx->launches++;
        }
        std::vector<unsigned long long> hc((size_t)R, 0ULL);
        PYCI_CUDA(cudaMemcpyAsync(hc.data(), dcnt, sizeof(unsigned long long) * (size_t)R, cudaMemcpyDeviceToHost, st));
        PYCI_CUDA(cudaStreamSynchronize(st));
        std::vector<long> scount((size_t)R), soff((size_t)R + 1, 0), rcount((size_t)R), roff((size_t)R + 1, 0);
        std::vector<unsigned long long> hcur((size_t)R);
        for (int p = 0; p < R; ++p) {
            scount[(size_t)p] = (long)hc[(size_t)p];
            soff[(size_t)p + 1] = soff[(size_t)p] + scount[(size_t)p];
            hcur[(size_t)p] = (unsigned long long)soff[(size_t)p];
        }
        // counts matrix [sender][receiver] summed over ranks (every rank fills its own row)
        std::vector<long> mat((size_t)R * R, 0);
        for (int p = 0; p < R; ++p)
            mat[(size_t)me * R + p] = scount[(size_t)p];
        PYCI_TRY(comm_allreduce_sum_i64_host(ctx, mat.data(), R * R));
        for (int q = 0; q < R; ++q) {
            rcount[(size_t)q] = mat[(size_t)q * R + me];
            roff[(size_t)q + 1] = roff[(size_t)q] + rcount[(size_t)q];
        }
        const long ns = std::max<long>(soff[(size_t)R], 1), nr = std::max<long>(roff[(size_t)R], 1);
        PYCI_CUDA(dev_malloc(&spay, sizeof(u64) * (size_t)ns));
        PYCI_CUDA(dev_malloc(&s0, sizeof(u64) * (size_t)ns));
        if (two)
            PYCI_CUDA(dev_malloc(&s1, sizeof(u64) * (size_t)ns));
        PYCI_CUDA(dev_malloc(&rpay, sizeof(u64) * (size_t)nr));
        PYCI_CUDA(dev_malloc(&r0, sizeof(u64) * (size_t)nr));
        if (two)
            PYCI_CUDA(dev_malloc(&r1, sizeof(u64) * (size_t)nr));
        if (L.n > 0) {
            PYCI_CUDA(cudaMemcpyAsync(dcnt + R, hcur.data(), sizeof(unsigned long long) * (size_t)R, cudaMemcpyHostToDevice, st));
            if (two)
                owner_scatter_kernel<true><<<blocks, 256, 0, st>>>(L.pay, L.k0, L.k1, L.n, (u32)R, dcnt + R, spay, s0, s1);
            else
                owner_scatter_kernel<false><<<blocks, 256, 0, st>>>(L.pay, L.k0, L.k1, L.n, (u32)R, dcnt + R, spay, s0, s1);
            ctx->launches++;
        }
        const bool trace = getenv("PYCI_B200_HCI_TRACE") != nullptr;
        auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        double tt = now();
        if (trace) {
            cudaStreamSynchronize(st);
            fprintf(stderr, "[pyci_b200 hci] rank %d:   partition by owner done (send %ld, receive %ld)\n", me, soff[(size_t)R], roff[(size_t)R]);
            tt = now();
        }
        PYCI_TRY(comm_alltoallv_u64(ctx, spay, scount.data(), soff.data(), rpay, rcount.data(), roff.data()));
        PYCI_TRY(comm_alltoallv_u64(ctx, s0, scount.data(), soff.data(), r0, rcount.data(), roff.data()));
        if (two)
            PYCI_TRY(comm_alltoallv_u64(ctx, s1, scount.data(), soff.data(), r1, rcount.data(), roff.data()));
        PYCI_CUDA(cudaStreamSynchronize(st)); // hcur is read by the copy above
        if (trace) {
            fprintf(stderr, "[pyci_b200 hci] rank %d:   all-to-all %.2f ms\n", me, 1e3 * (now() - tt));
            tt = now();
        }
        const std::vector<long> one(1, roff[(size_t)R]);
        PYCI_TRY(merge_lists<MODE>(ctx, two, rpay, r0, r1, nr, one, owned));
        if (trace) {
            cudaStreamSynchronize(st);
            fprintf(stderr, "[pyci_b200 hci] rank %d:   merge at the owner %.2f ms (%ld owned)\n", me, 1e3 * (now() - tt), owned.n);
        }
        return PYCI_OK;
    };
    const int rc = body();
    dev_free(dcnt);
    dev_free(spay);
    dev_free(s0);
    dev_free(s1);
    dev_free(rpay);
    dev_free(r0);
    dev_free(r1);
    return rc;
}

// every rank's owned list, concatenated in rank order, on every rank (the owners' lists are disjoint: no merge)
int allgather_lists(pyci_ctx *ctx, bool two, ExtList &mine, ExtList &all) {
    const int R = ctx->nranks;
    cudaStream_t st = ctx->stream;
    std::vector<long> rc((size_t)R, 0);
    rc[(size_t)ctx->rank] = mine.n;
    PYCI_TRY(comm_allreduce_sum_i64_host(ctx, rc.data(), R));
    long stride = 1, total = 0;
    for (long c : rc) {
        stride = std::max(stride, c);
        total += c;
    }
    u64 *sp = nullptr, *s0 = nullptr, *s1 = nullptr, *g = nullptr;
    auto body = [&]() -> int {
        std::vector<ExtList> one(1, mine);
        PYCI_TRY(pack_lists(ctx, two, one, stride, &sp, &s0, &s1)); // padded to the common stride
        PYCI_CUDA(dev_malloc(&g, sizeof(u64) * (size_t)stride * (size_t)R));
        const size_t ob = sizeof(u64) * (size_t)std::max<long>(total, 1);
        PYCI_CUDA(dev_malloc(&all.pay, ob));
        PYCI_CUDA(dev_malloc(&all.k0, ob));
        if (two)
            PYCI_CUDA(dev_malloc(&all.k1, ob));
        all.n = total;
        u64 *src[3] = {sp, s0, s1}, *dst[3] = {all.pay, all.k0, all.k1};
        for (int a = 0; a < (two ? 3 : 2); ++a) {
            PYCI_TRY(comm_allgather_f64(ctx, (const double *)src[a], (double *)g, stride)); // bit patterns: no arithmetic
            long off = 0;
            for (int p = 0; p < R; ++p) {
                if (rc[(size_t)p] > 0)
                    PYCI_CUDA(cudaMemcpyAsync(dst[a] + off, g + (size_t)p * stride, sizeof(u64) * (size_t)rc[(size_t)p],
                                              cudaMemcpyDeviceToDevice, st));
                off += rc[(size_t)p];
            }
        }
        return PYCI_OK;
    };
    const int rcode = body();
    dev_free(sp);
    dev_free(s0);
    dev_free(s1);
    dev_free(g);
    if (rcode != PYCI_OK)
        all.release();
    return rcode;
}

// The external space of the whole wave function as a duplicate-free list: on every rank for add_hci; for ENPT2, when
// row-sharded, each rank gets the determinants it OWNS (hash of the string mod ranks -- the share pt2_reduce_kernel
// reduces).  This rank's rows are walked in `split` consecutive chunks (PYCI_B200_EXT_SPLIT, default 1: bounds the size
// of one table; the chunks' lists are merged locally), then the per-rank lists are exchanged to their owners.
template<int MODE>
int collect_external(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, const double *coeffs_dev, double eps,
                     ExtList &out, double *seconds) {
    BuildParams P;
    PYCI_TRY(enum_params_init(P, wfn));
    OrderParams O;
    {
        const u32 na = (u32)wfn->nocc_up, nva = (u32)(wfn->nbasis - wfn->nocc_up);
        const bool fc = wfn->kind == PYCI_FULLCI;
        const u32 nb = fc ? (u32)wfn->nocc_dn : 0u, nvb = fc ? (u32)(wfn->nbasis - wfn->nocc_dn) : 0u;
        O.nva = nva;
        O.nvb = nvb;
        O.nSb = nb * nvb;
        O.Ma = 1u + O.nSb + na * nva;
        O.Mb = 1u + nb * nvb;
        O.offB = na * nva * O.Ma;
        if ((double)O.offB + (double)nb * nvb * O.Mb >= 16777216.0)
            PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "excitation space per determinant too large for the 24-bit order key");
        if (wfn->ndet >= (1L << 39))
            PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "too many determinants");
    }
    const long R = ctx->nranks, ndet = wfn->ndet;
    const long per = (ndet + R - 1) / R;
    const long my0 = std::min(ndet, per * ctx->rank), myn = std::min(ndet, per * (ctx->rank + 1)) - my0;
    P.ncol = ndet;
    P.one_mo = ham->one_mo;
    P.two_mo = ham->two_mo;
    P.h = ham->h;
    P.v = ham->v;
    P.w = ham->w;
    P.coeffs = coeffs_dev;
    const bool two = wfn->kind == PYCI_FULLCI;
    long split = 1;
    if (const char *e = getenv("PYCI_B200_EXT_SPLIT"))
        split = std::max(1L, std::min(64L, atol(e)));

    PYCI_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    std::vector<ExtList> parts((size_t)split);
    ExtList L;
    u64 *gp = nullptr, *g0 = nullptr, *g1 = nullptr, *sp = nullptr, *s0 = nullptr, *s1 = nullptr;
    auto body = [&]() -> int {
        const long chunk = (myn + split - 1) / split;
        std::vector<long> counts;
        long stride = 1;
        for (long p = 0; p < split; ++p) {
            const long lo = std::min(myn, chunk * p), hi = std::min(myn, chunk * (p + 1));
            PYCI_TRY(walk_rows<MODE>(ctx, wfn, P, O, eps, my0 + lo, hi - lo, parts[(size_t)p]));
            counts.push_back(parts[(size_t)p].n);
            stride = std::max(stride, parts[(size_t)p].n);
        }
        if (split == 1) {
            L = parts[0];
            parts[0] = ExtList();
        } else {
            PYCI_TRY(pack_lists(ctx, two, parts, stride, &sp, &s0, &s1));
            PYCI_TRY(merge_lists<MODE>(ctx, two, sp, s0, s1, stride, counts, L));
            dev_free(sp);
            dev_free(s0);
            dev_free(s1);
            sp = s0 = s1 = nullptr;
        }
        return PYCI_OK;
    };
    auto exchange = [&]() -> int {
        if (R > 1) {
            // owner-computes: duplicates found by different ranks meet at the determinant's owner.  ENPT2 stops there
            // (each rank reduces the determinants it owns); add_hci needs the whole list everywhere: the owners'
            // duplicate-free lists are all-gathered and concatenated.
            ExtList owned;
            PYCI_TRY(exchange_to_owners<MODE>(ctx, two, L, owned));
            L.release();
            if (MODE == MODE_PT2) {
                L = owned;
            } else {
                const int rcg = allgather_lists(ctx, two, owned, L);
                owned.release();
                PYCI_TRY(rcg);
            }
        }
        return PYCI_OK;
    };
    // PYCI_B200_HCI_TRACE: host wall time per phase on stderr (each phase closed by a stream synchronisation)
    const bool trace = getenv("PYCI_B200_HCI_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    int rc = body();
    if (trace) {
        cudaStreamSynchronize(ctx->stream);
        fprintf(stderr, "[pyci_b200 hci] rank %d: walk + local merge %.2f ms (%ld entries)\n", ctx->rank, 1e3 * (now() - t0), L.n);
        t0 = now();
    }
    if (R > 1) {
        // a rank-local failure of the walk (table growth out of memory, ...) must not leave the other ranks waiting
        // in the exchange: agree on the status first, every rank fails together
        long bad = rc != PYCI_OK;
        const int rca = comm_allreduce_sum_i64_host(ctx, &bad, 1);
        if (rc == PYCI_OK && rca != PYCI_OK)
            rc = rca;
        if (rc == PYCI_OK && bad) {
            pyci_set_error("another rank failed while walking its rows of the external space");
            rc = PYCI_ERR_RUNTIME;
        }
    }
    if (rc == PYCI_OK)
        rc = exchange();
    if (trace) {
        cudaStreamSynchronize(ctx->stream);
        fprintf(stderr, "[pyci_b200 hci] rank %d: exchange to owners%s %.2f ms (%ld entries)\n", ctx->rank,
                MODE == MODE_HCI ? " + all-gather" : "", 1e3 * (now() - t0), L.n);
    }
    for (ExtList &q : parts)
        q.release();
    dev_free(gp);
    dev_free(g0);
    dev_free(g1);
    dev_free(sp);
    dev_free(s0);
    dev_free(s1);
    if (rc != PYCI_OK) {
        L.release();
        return rc;
    }
    PYCI_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    if (seconds)
        *seconds = ms * 1e-3;
    out = L;
    return PYCI_OK;
}

int upload_coeffs(pyci_ctx *ctx, const double *coeffs, long ndet, double **dc) {
    PYCI_CUDA(dev_malloc(dc, sizeof(double) * (size_t)std::max<long>(ndet, 1)));
    PYCI_CUDA(cudaMemcpyAsync(*dc, coeffs, sizeof(double) * (size_t)ndet, cudaMemcpyHostToDevice, ctx->stream));
    return PYCI_OK;
}

} // namespace

// add_hci (hci.cpp:238-279): appends the selected determinants to the device wave function (first-encounter
// order), rebuilds its index, and reports how many were added.
int add_hci_impl(pyci_ctx *ctx, const pyci_ham *ham, pyci_wfn *wfn, const double *coeffs, double eps, long *n_new,
                 double *seconds) {
    PYCI_NVTX("pyci:add_hci");
    double *dc = nullptr;
    ExtList L;
    u32 *order_in = nullptr, *order_out = nullptr;
    u64 *keys_out = nullptr, *newdets = nullptr;
    void *tmp = nullptr;
    auto body = [&]() -> int {
        PYCI_TRY(upload_coeffs(ctx, coeffs, wfn->ndet, &dc));
        PYCI_TRY(collect_external<MODE_HCI>(ctx, ham, wfn, dc, eps, L, seconds));
        *n_new = L.n;
        if (L.n == 0)
            return PYCI_OK;
        if (wfn->ndet + L.n >= (1L << 31) - 1)
            PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "ndet = %ld out of range for int32 column indices", wfn->ndet + L.n);
        // order by first encounter: radix sort of (payload, list position)
        const long n = L.n;
        PYCI_CUDA(dev_malloc(&order_in, sizeof(u32) * (size_t)n));
        PYCI_CUDA(dev_malloc(&order_out, sizeof(u32) * (size_t)n));
        PYCI_CUDA(dev_malloc(&keys_out, sizeof(u64) * (size_t)n));
        iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(order_in, n);
        ctx->launches++;
        bool in_alt = false;
        int pbits = 25; // payload = row << 24 | position in the row's loop nest
        while ((1L << (pbits - 24)) < wfn->ndet)
            ++pbits;
        PYCI_TRY(radix_sort_pairs<u32>(ctx, L.pay, keys_out, order_in, order_out, n, pbits, &in_alt));
        if (!in_alt)
            std::swap(order_in, order_out); // ext_gather_dets_kernel reads order_out
        // grow the determinant array and append
        const int nw = wfn->nwords;
        PYCI_CUDA(dev_malloc(&newdets, sizeof(u64) * (size_t)((wfn->ndet + n) * nw)));
        PYCI_CUDA(cudaMemcpyAsync(newdets, wfn->dets, sizeof(u64) * (size_t)(wfn->ndet * nw), cudaMemcpyDeviceToDevice,
                                  ctx->stream));
        ext_gather_dets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(order_out, L.k0, L.k1, n, nw,
                                                                                    newdets + wfn->ndet * nw);
        ctx->launches++;
        PYCI_CUDA(cudaGetLastError());
        dev_free(wfn->dets);
        wfn->dets = newdets;
        newdets = nullptr;
        wfn->ndet += n;
        wfn->complete = false; // re-derived below
        dev_free(wfn->slots);
        wfn->slots = nullptr;
        wfn->index_valid = false;
        wfn->generated = false;
        return PYCI_OK;
    };
    int rc = body();
    dev_free(dc);
    dev_free(order_in);
    dev_free(order_out);
    dev_free(keys_out);
    dev_free(tmp);
    dev_free(newdets);
    L.release();
    return rc;
}

// compute_enpt2 (enpt2.cpp:344-374) for FullCI / GenCI wave functions
int enpt2_impl(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, const double *coeffs, double energy, double eps,
               double *out, long *nterms, double *seconds) {
    PYCI_NVTX("pyci:compute_enpt2");
    double *dc = nullptr, *acc = nullptr;
    ExtList L;
    auto body = [&]() -> int {
        PYCI_TRY(upload_coeffs(ctx, coeffs, wfn->ndet, &dc));
        PYCI_TRY(collect_external<MODE_PT2>(ctx, ham, wfn, dc, eps, L, seconds));
        PYCI_CUDA(dev_malloc(&acc, sizeof(double)));
        PYCI_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), ctx->stream));
        BuildParams P;
        PYCI_TRY(enum_params_init(P, wfn));
        P.one_mo = ham->one_mo;
        P.two_mo = ham->two_mo;
        // every rank holds the merged list: reduce the determinants this rank owns, then sum over ranks
        if (L.n > 0) {
            const unsigned blocks = (unsigned)std::min<long>((L.n + 255) / 256, (long)ctx->sm_count * 4);
            if (wfn->kind == PYCI_FULLCI)
                pt2_reduce_kernel<PYCI_FULLCI><<<blocks, 256, 0, ctx->stream>>>(P, L.pay, L.k0, L.k1, L.n, (u32)ctx->rank,
                                                                               (u32)ctx->nranks, energy - ham->ecore, acc);
            else
                pt2_reduce_kernel<PYCI_GENCI><<<blocks, 256, 0, ctx->stream>>>(P, L.pay, L.k0, L.k1, L.n, (u32)ctx->rank,
                                                                              (u32)ctx->nranks, energy - ham->ecore, acc);
            ctx->launches++;
            PYCI_CUDA(cudaGetLastError());
        }
        PYCI_TRY(comm_allreduce_sum_f64(ctx, acc, 1));
        double corr = 0.0;
        PYCI_CUDA(cudaMemcpyAsync(&corr, acc, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
        *out = energy + corr;
        if (nterms) { // row-sharded: L holds the determinants this rank owns
            long tot = L.n;
            PYCI_TRY(comm_allreduce_sum_i64_host(ctx, &tot, 1));
            *nterms = tot;
        }
        return PYCI_OK;
    };
    int rc = body();
    dev_free(dc);
    dev_free(acc);
    L.release();
    return rc;
}
