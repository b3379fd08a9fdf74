// Shared declarations of the pyci_b200 CUDA library (sm_100a only).
// Host-side structs behind the opaque C-ABI handles of include/pyci_b200.h, error plumbing, and the
// device-side determinant primitives (the device restatement of /root/reference/pyci/src/common.cpp).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/pyci_b200.h"

// NVTX ranges around the phases of the path (index / count / scan / fill / SpMV / all-gather / solve / RDM / HCI):
// header-only NVTX v3, a no-op unless a profiler (nsys, ncu --nvtx) injects its library.
#include <nvtx3/nvToolsExt.h>
struct PyciRange {
    explicit PyciRange(const char *name) { nvtxRangePushA(name); }
    ~PyciRange() { nvtxRangePop(); }
    PyciRange(const PyciRange &) = delete;
    PyciRange &operator=(const PyciRange &) = delete;
};
#define PYCI_CAT2(a, b) a##b
#define PYCI_CAT(a, b) PYCI_CAT2(a, b)
#define PYCI_NVTX(name) PyciRange PYCI_CAT(pyci_range_, __LINE__)(name)

typedef unsigned long long u64;
typedef unsigned int u32;

// ---------------------------------------------------------------------------------------------
// error plumbing

void pyci_set_error(const char *fmt, ...);

#define PYCI_CUDA(expr)                                                                            \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            pyci_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                           cudaGetErrorString(_e));                                                \
            return (_e == cudaErrorMemoryAllocation) ? PYCI_ERR_MEMORY : PYCI_ERR_CUDA;            \
        }                                                                                          \
    } while (0)

#define PYCI_TRY(expr)                                                                             \
    do {                                                                                           \
        int _s = (expr);                                                                           \
        if (_s != PYCI_OK)                                                                         \
            return _s;                                                                             \
    } while (0)

#define PYCI_FAIL(code, ...)                                                                       \
    do {                                                                                           \
        pyci_set_error(__VA_ARGS__);                                                               \
        return (code);                                                                             \
    } while (0)

// ---------------------------------------------------------------------------------------------
// handles

struct pyci_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int rank = 0, nranks = 1;
    void *comm = nullptr; // ncclComm_t
    long launches = 0;
    int sm_count = 148;
    int smem_optin = 0; // max dynamic shared memory per block (opt-in)
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // binomial table C(p, j), p < binom_n, j < binom_k1, of the last sorted-space build (the same for every build of
    // one problem size: uploaded once instead of once per construction)
    unsigned *binom_dev = nullptr;
    int binom_n = 0, binom_k1 = 0;
    // slots of the external-space table that held the last walk of ext_hint_rows rows (hci.cu: first guess of the next)
    long ext_hint_rows = -1, ext_hint_cap = 0;
};

struct pyci_ham {
    pyci_ctx *ctx = nullptr;
    long nbasis = 0;
    double ecore = 0.0;
    double *one_mo = nullptr, *two_mo = nullptr, *h = nullptr, *v = nullptr, *w = nullptr;
    bool kl_sym = false; // two_mo[i,k,a,l] == two_mo[i,l,a,k] bit for bit, all indices (checked at upload): the
                         // complete-space fill may then keep only k <= l of its shared-memory integral slices
};

// how determinant strings are packed into hash keys
enum KeyMode {
    KEY32 = 0,  // one u32: one-spin nbasis<=32, or two-spin nbasis<=16 as alpha | beta<<nbasis   ( 8 B slots)
    KEY64 = 1,  // one u64: one-spin nbasis<=64, or two-spin nbasis<=32 as alpha | beta<<nbasis   (16 B slots)
    KEY128 = 2, // two u64 (alpha, beta): two-spin 32<nbasis<=64                                  (32 B slots)
    KEY_MW = 3  // nbasis > 64: slots hold determinant indices, keys are compared in the determinant array (multiword.cu)
};

struct pyci_wfn {
    pyci_ctx *ctx = nullptr;
    int kind = 0;
    long nbasis = 0, nocc_up = 0, nocc_dn = 0, ndet = 0;
    int nwords = 1;       // u64 words per determinant in `dets` (1 one-spin, 2 two-spin)
    int keymode = KEY64;
    bool complete = false; // ndet equals the size of the full space => every excitation is present
    bool generated = false; // determinants unranked on the device (pyci_wfn_create_all_dets), never appended to
    bool sorted2 = false;  // two-spin determinants ascend in (alpha, beta) as integers (add_all_dets order)
    u64 *dets = nullptr;   // [ndet][nwords]
    void *slots = nullptr; // hash slots (layout by keymode)
    bool index_valid = false; // slots / bloom hold the current determinants (built lazily for complete sorted spaces)
    u32 mask = 0;          // capacity-1 (capacity is a power of two)
    u32 *bloom = nullptr;  // blocked Bloom filter over the keys (one 32-bit word per block, two bits per key), built
    u32 bmask = 0;         // only when the slot table outgrows L2: a miss then costs an L2 hit instead of an HBM sector
    double hash_seconds = 0.0;
    double ext_seconds = 0.0; // device seconds of the last add_hci / compute_enpt2 walk over this wfn
};

struct pyci_op {
    pyci_ctx *ctx = nullptr;
    long nrow = 0, ncol = 0; // global shape
    long row0 = 0, nloc = 0; // this rank's rows [row0, row0+nloc)
    long npad = 0;           // stride of a rank's vectors: rows per rank of the uniform partition, npad*nranks >= nrow
                             // (nnz-balanced partition: the largest row count of any rank)
    std::vector<long> bounds; // nnz-balanced partition (rebalance.cu): rank p holds rows [bounds[p], bounds[p+1]);
                              // empty = uniform blocks of npad rows
    int symmetric = 0;
    bool foreign = false;    // built by pyci_op_build_shard for a rank layout other than the context's: no collectives
    double ecore = 0.0;
    long nnz = 0;            // stored non-zeros (full rows)
    long size_ref = 0;       // SparseOp::size in the reference's storage for these rows
    long *indptr = nullptr;  // [nloc+1] device, int64
    int *cols = nullptr;     // [nnz] device, int32
    double *vals = nullptr;  // [nnz] device, fp64
    int *lowcnt = nullptr;   // [nloc] entries with col <= row (sorted rows => a prefix)
    double *diag = nullptr;  // [npad] diagonal H_ii of this rank's rows (0 where absent)
    double times[4] = {0, 0, 0, 0};
    bool joined = false;       // selected space: entries found by the segment-pair join (join.cuh), not by enumeration
    double join_tests = 0.0;   // predicted XOR/popcount tests of that join for this rank's rows
    double fill_seconds = 0.0; // device seconds of the fill kernel alone (CUDA events around its launch)
    const char *fill_kernel = "none"; // which fill path built this operator
    const char *count_kernel = "none"; // what found the stored entries: "analytic", "count_kernel", "join_rows_kernel"
    int spmv_tpr = 0, spmv_ctas = 4; // SpMV launch shape (threads per row, CTAs per SM), chosen at first use;
                                     // spmv_tpr < 0: bulk-copy stream kernel with -spmv_tpr warps per CTA, ring depth spmv_ctas
    int spmv_depth = 2;              // trips of stream loads in flight per thread in spmv_rows (2, 3 or 4)
    int spmv_block = 256;            // threads per CTA of spmv_rows (256, 512 or 1024)
    long *spmv_part = nullptr;       // stream kernel: first row of every warp's range [spmv_part_n + 1]
    int spmv_part_n = 0;
    // scratch for host-facing matvec
    double *xbuf = nullptr, *ybuf = nullptr;
    // pyci_op_get_element: host copy of the row asked for last (callers walk a row element by element,
    // pyci/test/test_routines.py:75-78); -1 = none.  Every site that replaces cols/vals resets it.
    long ge_row = -1;
    std::vector<int> ge_cols;
    std::vector<double> ge_vals;
    // nnz-balanced partition: staging of the all-gather of unequal shards [nranks][npad], the partition on the device
    double *gather_stage = nullptr;
    long *bounds_dev = nullptr;
};

// Makes ctx's device current and its stream the target of dev_malloc / dev_free on this thread.
int ctx_activate(const pyci_ctx *ctx);

// Stream-ordered device allocation from the device's default memory pool (cudaMallocAsync on the active
// context's stream).  pyci_ctx_create raises the pool's release threshold, so the multi-GB CSR buffers of
// a destroyed operator are reused by the next build instead of going back to the driver.
cudaStream_t dev_current_stream();
template<class T>
inline cudaError_t dev_malloc(T **p, size_t bytes) {
    return cudaMallocAsync(reinterpret_cast<void **>(p), bytes, dev_current_stream());
}
inline cudaError_t dev_free(void *p) { return p ? cudaFreeAsync(p, dev_current_stream()) : cudaSuccess; }

// nccl (loaded lazily with dlopen; see comm.cpp)
int comm_unique_id(void *out128);
int comm_init(pyci_ctx *ctx, int rank, int nranks, const void *id128);
void comm_destroy(pyci_ctx *ctx);
int comm_allgather_f64(pyci_ctx *ctx, const double *send_dev, double *recv_dev, long count_per_rank);
int comm_allgatherv_f64(pyci_ctx *ctx, const double *send_dev, double *recv_dev, const long *bounds);
int comm_alltoallv_bytes(pyci_ctx *ctx, const void *send, const long *scount, const long *soff, void *recv,
                         const long *rcount, const long *roff);
int comm_allreduce_sum_f64(pyci_ctx *ctx, double *buf_dev, long count);
int comm_allreduce_sum_i64_host(pyci_ctx *ctx, long *vals, int count);
int comm_alltoallv_u64(pyci_ctx *ctx, const unsigned long long *send, const long *scount, const long *soff,
                       unsigned long long *recv, const long *rcount, const long *roff);

// multiword.cu: nbasis > 64
int mw_index_build(pyci_wfn *wfn);
int mw_op_build(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, pyci_op *op);
// dets.cu
int wfn_generate_all_dets(pyci_wfn *wfn, long na, long nb);
// build.cu
int wfn_build_index(pyci_wfn *wfn);
int wfn_ensure_index(const pyci_wfn *wfn); // builds the hash index if it was deferred
long op_size_ref(pyci_op *op);             // SparseOp::size of this rank's rows (lazily summed)
int op_build_impl(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, pyci_op *op);
int scan_counts(pyci_ctx *ctx, const int *cnt, long n, long *indptr, int *maxcnt);
// update.cu
int op_update_impl(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, pyci_op *op);
// rebalance.cu: moves rows between neighbouring ranks so that every rank stores the same number of entries (collective)
int op_rebalance(pyci_ctx *ctx, pyci_op *op);
// every rank's shard of a row-distributed vector, concatenated in row order, on every rank
int op_allgather_rows(pyci_ctx *ctx, pyci_op *op, const double *send_dev, double *recv_dev);
// spmv.cu
int spmv_launch(pyci_op *op, const double *x_dev, double *y_dev);
// solver.cu
int solve_impl(pyci_op *op, long n, const double *c0, long ncv, long maxiter, double tol,
               double *evals, double *evecs, pyci_solve_stats *stats);
// rdm.cu
int rdms_impl(pyci_ctx *ctx, const pyci_wfn *wfn, const double *coeffs, double *rdm1, double *rdm2);

int trdms_impl(pyci_ctx *ctx, const pyci_wfn *wfn1, const pyci_wfn *wfn2, const double *coeffs1, const double *coeffs2,
               double *rdm1, double *rdm2);
int overlap_impl(pyci_ctx *ctx, const pyci_wfn *wfn1, const pyci_wfn *wfn2, const double *coeffs1, const double *coeffs2,
                 double *out);
// hci.cu
int add_hci_impl(pyci_ctx *ctx, const pyci_ham *ham, pyci_wfn *wfn, const double *coeffs, double eps, long *n_new,
                 double *seconds);
int enpt2_impl(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, const double *coeffs, double energy, double eps,
               double *out, long *nterms, double *seconds);

// ---------------------------------------------------------------------------------------------
// device primitives

#ifdef __CUDACC__

// 64-bit finaliser of MurmurHash3 / 32-bit finaliser: the hash of the GPU determinant index
__device__ __forceinline__ u32 mix32(u32 h) {
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
__device__ __forceinline__ u32 mix64(u64 h) {
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return (u32)h;
}

// Bloom probe of the determinant index: word and the two bits of a key from its 32-bit hash
__device__ __forceinline__ bool bloom_pass(const u32 *__restrict__ bloom, u32 bmask, u32 h) {
    const u32 g = h * 0x9e3779b1u;
    const u32 w = __ldg(bloom + ((g >> 10) & bmask));
    return ((w >> (g & 31u)) & (w >> ((g >> 5) & 31u)) & 1u) != 0u;
}
__device__ __forceinline__ void bloom_set(u32 *bloom, u32 bmask, u32 h) {
    const u32 g = h * 0x9e3779b1u;
    atomicOr(bloom + ((g >> 10) & bmask), (1u << (g & 31u)) | (1u << ((g >> 5) & 31u)));
}

struct __align__(8) Slot32 {
    u32 key;
    int val;
};
struct __align__(16) Slot64 {
    u64 key;
    int val;
    int pad;
};
struct __align__(32) Slot128 {
    u64 k0, k1;
    int val;
    int pad[3];
};

// Determinant index: the device form of Wfn::index_det (onespinwfn.cpp:123-126, twospinwfn.cpp:129-132).
// a = alpha string (or the only string), b = beta string (0 for one-spin), shift = nbasis (two-spin packing).
template<int KM>
struct DetIndex;

template<>
struct DetIndex<KEY32> {
    const Slot32 *slots;
    u32 mask;
    int shift;
    const u32 *bloom;
    u32 bmask;
    __device__ __forceinline__ u32 key(u64 a, u64 b) const { return (u32)a | ((u32)b << shift); }
    __device__ __forceinline__ u32 hash(u64 a, u64 b) const { return mix32(key(a, b)); }
    __device__ __forceinline__ int find(u64 a, u64 b) const {
        const u32 h = hash(a, b);
        if (bloom && !bloom_pass(bloom, bmask, h))
            return -1;
        return probe(a, b, h);
    }
    __device__ __forceinline__ int probe(u64 a, u64 b, u32 h) const {
        const u32 k = key(a, b);
        u32 p = h & mask;
        for (;;) {
            const uint2 s = __ldg(reinterpret_cast<const uint2 *>(slots + p));
            if ((int)s.y < 0)
                return -1;
            if (s.x == k)
                return (int)s.y;
            p = (p + 1) & mask;
        }
    }
};

template<>
struct DetIndex<KEY64> {
    const Slot64 *slots;
    u32 mask;
    int shift;
    const u32 *bloom;
    u32 bmask;
    __device__ __forceinline__ u64 key(u64 a, u64 b) const { return a | (b << shift); }
    // 32-bit mixing only (a 64-bit multiply is four to six instructions; the enumeration hashes 10^5 keys per row)
    __device__ __forceinline__ u32 hash(u64 a, u64 b) const {
        const u64 k = key(a, b);
#ifdef PYCI_HASH_MIX64
        return mix64(k);
#else
        return mix32((u32)k ^ ((u32)(k >> 32) * 0x9e3779b1u));
#endif
    }
    __device__ __forceinline__ int find(u64 a, u64 b) const {
        const u32 h = hash(a, b);
        if (bloom && !bloom_pass(bloom, bmask, h))
            return -1;
        return probe(a, b, h);
    }
    __device__ __forceinline__ int probe(u64 a, u64 b, u32 h) const {
        const u64 k = key(a, b);
        u32 p = h & mask;
        for (;;) {
            const uint4 s = __ldg(reinterpret_cast<const uint4 *>(slots + p));
            if ((int)s.z < 0)
                return -1;
            if ((((u64)s.y << 32) | s.x) == k)
                return (int)s.z;
            p = (p + 1) & mask;
        }
    }
};

template<>
struct DetIndex<KEY128> {
    const Slot128 *slots;
    u32 mask;
    int shift;
    const u32 *bloom;
    u32 bmask;
    __device__ __forceinline__ u32 hash(u64 a, u64 b) const {
        return mix32((u32)a ^ ((u32)(a >> 32) * 0x9e3779b1u) ^ ((u32)b * 0x85ebca6bu) ^ ((u32)(b >> 32) * 0xc2b2ae35u));
    }
    __device__ __forceinline__ int find(u64 a, u64 b) const {
        const u32 h = hash(a, b);
        if (bloom && !bloom_pass(bloom, bmask, h))
            return -1;
        return probe(a, b, h);
    }
    __device__ __forceinline__ int probe(u64 a, u64 b, u32 h) const {
        u32 p = h & mask;
        for (;;) {
            const uint4 s0 = __ldg(reinterpret_cast<const uint4 *>(slots + p));
            const int v = __ldg(reinterpret_cast<const int *>(slots + p) + 4);
            if (v < 0)
                return -1;
            if ((((u64)s0.y << 32) | s0.x) == a && (((u64)s0.w << 32) | s0.z) == b)
                return v;
            p = (p + 1) & mask;
        }
    }
};

// U lookups at once.  With the Bloom filter in place the U filter words are loaded back to back (independent loads
// in flight) before any slot is probed: the enumeration of a selected space is bound by the latency of exactly that
// load.  want[u] = false skips candidate u (out = -1).
template<int U, class Index>
__device__ __forceinline__ void find_batch(const Index &ix, const u64 (&A)[U], const u64 (&B)[U], const bool (&want)[U],
                                           int (&out)[U]) {
    if (ix.bloom) {
        u32 h[U], w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            h[u] = ix.hash(A[u], B[u]);
            const u32 g = h[u] * 0x9e3779b1u;
            w[u] = want[u] ? __ldg(ix.bloom + ((g >> 10) & ix.bmask)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const u32 g = h[u] * 0x9e3779b1u;
            out[u] = (((w[u] >> (g & 31u)) & (w[u] >> ((g >> 5) & 31u)) & 1u) != 0u) ? ix.probe(A[u], B[u], h[u]) : -1;
        }
    } else {
#pragma unroll
        for (int u = 0; u < U; ++u)
            out[u] = want[u] ? ix.find(A[u], B[u]) : -1;
    }
}

// bits strictly between positions lo < hi of one 64-bit string
__device__ __forceinline__ u64 between_mask(int lo, int hi) {
    return ((1ULL << hi) - 1ULL) & ~((2ULL << lo) - 1ULL);
}

// phase_single_det (common.cpp:163-194): parity of the occupied orbitals strictly between i and a,
// evaluated on the un-excited string.  Returns 0 for +1, 1 for -1.
__device__ __forceinline__ int parity_single(u64 det, int i, int a) {
    const int lo = min(i, a), hi = max(i, a);
    return __popcll(det & between_mask(lo, hi)) & 1;
}

// phase_double_det (common.cpp:196-261): both single parities on the original string plus one flip
// when (i2 < a1) or (i1 > a2) (:258-259).  Callers pass i1 < i2 and a1 < a2.
__device__ __forceinline__ int parity_double(u64 det, int i1, int i2, int a1, int a2) {
    return (parity_single(det, i1, a1) + parity_single(det, i2, a2) + ((i2 < a1) || (i1 > a2))) & 1;
}

__device__ __forceinline__ double apply_sign(double x, int parity) { return parity ? -x : x; }

#endif // __CUDACC__
