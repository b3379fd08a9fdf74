// C-ABI entry points of libpyci_b200.so (include/pyci_b200.h): argument checking with the
// reference's error behaviour, HBM allocation, host<->device staging, and dispatch to the kernels.
#include <stdarg.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

int wfn_index_dets_impl(pyci_wfn *wfn, long n, const u64 *dets_dev, long *out_dev);

static thread_local char g_error[1024] = "";

void pyci_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

static thread_local cudaStream_t g_stream = nullptr;

cudaStream_t dev_current_stream() { return g_stream; }

int ctx_activate(const pyci_ctx *ctx) {
    PYCI_CUDA(cudaSetDevice(ctx->device));
    g_stream = ctx->stream;
    return PYCI_OK;
}

namespace {

long binom_l(long n, long k) {
    if (k < 0 || k > n)
        return 0;
    if (k > n - k)
        k = n - k;
    __int128 b = 1;
    for (long d = 1; d <= k; ++d) {
        b = b * (n - k + d) / d;
        if (b > (__int128)1 << 62)
            return INT64_MAX;
    }
    return (long)b;
}

long full_space_size(int kind, long nbasis, long nocc_up, long nocc_dn) {
    if (kind != PYCI_FULLCI)
        return binom_l(nbasis, nocc_up);
    return (binom_l(nbasis, nocc_up) < (1L << 31) && binom_l(nbasis, nocc_dn) < (1L << 31))
               ? binom_l(nbasis, nocc_up) * binom_l(nbasis, nocc_dn)
               : INT64_MAX;
}

template<class T>
int upload(T **dst, const T *src, size_t count, cudaStream_t st) {
    PYCI_CUDA(dev_malloc(dst, sizeof(T) * std::max<size_t>(count, 1)));
    if (count)
        PYCI_CUDA(cudaMemcpyAsync(*dst, src, sizeof(T) * count, cudaMemcpyHostToDevice, st));
    return PYCI_OK;
}

// flag |= 1 when two_mo[i,k,a,l] != two_mo[i,l,a,k] somewhere (bit patterns compared)
__global__ void kl_symmetry_kernel(const double *__restrict__ two_mo, long n, int *flag) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long n2 = n * n, n3 = n2 * n;
    if (idx >= n2 * n2)
        return;
    const long i = idx / n3, k = (idx / n2) % n, a = (idx / n) % n, l = idx % n;
    if (k < l && __double_as_longlong(two_mo[idx]) != __double_as_longlong(two_mo[i * n3 + l * n2 + a * n + k]))
        atomicOr(flag, 1);
}

__global__ void export_lower_kernel(const long *__restrict__ indptr, const int *__restrict__ cols,
                                    const double *__restrict__ vals, const int *__restrict__ take,
                                    const long *__restrict__ outptr, long *__restrict__ out_idx,
                                    double *__restrict__ out_val, long nrows) {
    // one warp per row copies the first take[r] entries (all of them when take == nullptr)
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long r = warp; r < nrows; r += nwarps) {
        const long src = indptr[r], dst = outptr[r];
        const long cnt = outptr[r + 1] - dst;
        for (long e = lane; e < cnt; e += 32) {
            out_idx[dst + e] = cols[src + e];
            out_val[dst + e] = vals[src + e];
        }
    }
}

// rows[k] (global) -> first stored entry and number of exported entries (the col <= row prefix when `take` is given)
__global__ void row_extent_kernel(const long *__restrict__ indptr, const int *__restrict__ take,
                                  const long *__restrict__ rows, long row0, long nrows, long *__restrict__ src,
                                  long *__restrict__ cnt) {
    const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nrows)
        return;
    const long r = rows[k] - row0;
    src[k] = indptr[r];
    cnt[k] = take ? (long)take[r] : indptr[r + 1] - indptr[r];
}

__global__ void export_rows_kernel(const long *__restrict__ src, const int *__restrict__ cols,
                                   const double *__restrict__ vals, const long *__restrict__ outptr,
                                   long *__restrict__ out_idx, double *__restrict__ out_val, long nrows) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long k = warp; k < nrows; k += nwarps) {
        const long s0 = src[k], dst = outptr[k], cnt = outptr[k + 1] - dst;
        for (long e = lane; e < cnt; e += 32) {
            out_idx[dst + e] = cols[s0 + e];
            out_val[dst + e] = vals[s0 + e];
        }
    }
}

} // namespace

extern "C" {

const char *pyci_last_error(void) { return g_error; }

int pyci_abi_version(void) { return PYCI_B200_ABI_VERSION; }

int pyci_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int pyci_ctx_create(int device, void *stream, pyci_ctx **out) {
    if (!out)
        PYCI_FAIL(PYCI_ERR_VALUE, "null output pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        PYCI_FAIL(PYCI_ERR_CUDA, "no CUDA device available (%s); pyci_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= ndev)
        PYCI_FAIL(PYCI_ERR_VALUE, "device %d out of range (%d visible)", device, ndev);
    PYCI_CUDA(cudaSetDevice(device));
    pyci_ctx *ctx = new pyci_ctx();
    ctx->device = device;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
        ctx->own_stream = false;
    } else {
        cudaError_t se = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (se != cudaSuccess) {
            delete ctx;
            PYCI_CUDA(se);
        }
        ctx->own_stream = true;
    }
    cudaDeviceProp prop;
    PYCI_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    {
        // keep freed blocks in the pool (see dev_malloc)
        cudaMemPool_t pool = nullptr;
        PYCI_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ULL;
        PYCI_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    for (int i = 0; i < 6; ++i)
        PYCI_CUDA(cudaEventCreate(&ctx->ev[i]));
    *out = ctx;
    return PYCI_OK;
}

void pyci_ctx_destroy(pyci_ctx *ctx) {
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    comm_destroy(ctx);
    for (int i = 0; i < 6; ++i)
        if (ctx->ev[i])
            cudaEventDestroy(ctx->ev[i]);
    if (ctx->binom_dev)
        cudaFree(ctx->binom_dev);
    if (ctx->own_stream && ctx->stream)
        cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int pyci_ctx_synchronize(pyci_ctx *ctx) {
    PYCI_TRY(ctx_activate(ctx));
    PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
    return PYCI_OK;
}

int pyci_host_alloc(void **ptr, size_t bytes) {
    if (!ptr)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    *ptr = nullptr;
    PYCI_CUDA(cudaHostAlloc(ptr, std::max<size_t>(bytes, 1), cudaHostAllocPortable));
    return PYCI_OK;
}

void pyci_host_free(void *ptr) {
    if (ptr)
        cudaFreeHost(ptr);
}

int pyci_ctx_release_memory(pyci_ctx *ctx) {
    if (!ctx)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    PYCI_TRY(ctx_activate(ctx));
    PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaMemPool_t pool = nullptr;
    PYCI_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
    PYCI_CUDA(cudaMemPoolTrimTo(pool, 0));
    return PYCI_OK;
}

int pyci_nccl_unique_id(void *unique_id_128) { return comm_unique_id(unique_id_128); }

int pyci_ctx_init_comm(pyci_ctx *ctx, int rank, int nranks, const void *unique_id_128) {
    PYCI_TRY(ctx_activate(ctx));
    if (nranks < 1 || rank < 0 || rank >= nranks)
        PYCI_FAIL(PYCI_ERR_VALUE, "bad rank %d of %d", rank, nranks);
    if (nranks == 1) {
        ctx->rank = 0;
        ctx->nranks = 1;
        return PYCI_OK;
    }
    PYCI_TRY(comm_init(ctx, rank, nranks, unique_id_128));
    // NCCL sets its channels up lazily, on the first collective of each kind (seconds on an 8-GPU box): pay for it
    // here, not inside the first solve
    double *warm = nullptr;
    const long each = 1L << 17; // 1 MB per rank
    PYCI_CUDA(dev_malloc(&warm, sizeof(double) * (size_t)each * (size_t)(nranks + 1)));
    PYCI_CUDA(cudaMemsetAsync(warm, 0, sizeof(double) * (size_t)each * (size_t)(nranks + 1), ctx->stream));
    int rc = comm_allgather_f64(ctx, warm, warm + each, each);
    if (rc == PYCI_OK)
        rc = comm_allreduce_sum_f64(ctx, warm, 64);
    if (rc == PYCI_OK)
        rc = comm_allreduce_sum_f64(ctx, warm, each);
    cudaStreamSynchronize(ctx->stream);
    dev_free(warm);
    return rc;
}

int pyci_ctx_rank(const pyci_ctx *ctx) { return ctx->rank; }
int pyci_ctx_nranks(const pyci_ctx *ctx) { return ctx->nranks; }
long pyci_ctx_launch_count(const pyci_ctx *ctx) { return ctx->launches; }
void pyci_ctx_reset_launch_count(pyci_ctx *ctx) { ctx->launches = 0; }

// ---- Hamiltonian -------------------------------------------------------------------------------

int pyci_ham_upload(pyci_ctx *ctx, long nbasis, double ecore, const double *one_mo, const double *two_mo,
                    const double *h, const double *v, const double *w, pyci_ham **out) {
    if (!ctx || !out)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    *out = nullptr;
    if (nbasis < 1)
        PYCI_FAIL(PYCI_ERR_VALUE, "nbasis must be positive");
    PYCI_TRY(ctx_activate(ctx));
    pyci_ham *ham = new pyci_ham();
    ham->ctx = ctx;
    ham->nbasis = nbasis;
    ham->ecore = ecore;
    const size_t n1 = (size_t)nbasis, n2 = n1 * n1;
    int rc = PYCI_OK;
    if (one_mo && rc == PYCI_OK)
        rc = upload(&ham->one_mo, one_mo, n2, ctx->stream);
    if (two_mo && rc == PYCI_OK)
        rc = upload(&ham->two_mo, two_mo, n2 * n2, ctx->stream);
    if (h && rc == PYCI_OK)
        rc = upload(&ham->h, h, n1, ctx->stream);
    if (v && rc == PYCI_OK)
        rc = upload(&ham->v, v, n2, ctx->stream);
    if (w && rc == PYCI_OK)
        rc = upload(&ham->w, w, n2, ctx->stream);
    int *dflag = nullptr, hflag = 1;
    if (rc == PYCI_OK && two_mo && dev_malloc(&dflag, sizeof(int)) == cudaSuccess) {
        cudaMemsetAsync(dflag, 0, sizeof(int), ctx->stream);
        const long total = (long)(n2 * n2);
        kl_symmetry_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(ham->two_mo, nbasis, dflag);
        ctx->launches++;
        cudaMemcpyAsync(&hflag, dflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (rc == PYCI_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        pyci_set_error("CUDA error while uploading integrals: %s", cudaGetErrorString(cudaGetLastError()));
        rc = PYCI_ERR_CUDA;
    }
    ham->kl_sym = two_mo && dflag && hflag == 0;
    dev_free(dflag);
    if (rc != PYCI_OK) {
        pyci_ham_destroy(ham);
        return rc;
    }
    *out = ham;
    return PYCI_OK;
}

void pyci_ham_destroy(pyci_ham *ham) {
    if (!ham)
        return;
    ctx_activate(ham->ctx);
    dev_free(ham->one_mo);
    dev_free(ham->two_mo);
    dev_free(ham->h);
    dev_free(ham->v);
    dev_free(ham->w);
    delete ham;
}

// ---- wave function -----------------------------------------------------------------------------

static int wfn_check_args(pyci_ctx *ctx, int kind, long nbasis, long nocc_up, long nocc_dn) {
    if (kind != PYCI_DOCI && kind != PYCI_FULLCI && kind != PYCI_GENCI)
        PYCI_FAIL(PYCI_ERR_VALUE, "unknown wave-function kind %d", kind);
    // Wfn::init checks (wfn.cpp:52-57) and the per-class ones (dociwfn.cpp:28, genciwfn.cpp:50)
    if (nocc_dn < 0)
        PYCI_FAIL(PYCI_ERR_VALUE, "nocc_dn is < 0");
    if (nocc_up < nocc_dn)
        PYCI_FAIL(PYCI_ERR_VALUE, "nocc_up is < nocc_dn");
    if (nbasis < nocc_up)
        PYCI_FAIL(PYCI_ERR_VALUE, "nbasis is < nocc_up");
    if (kind == PYCI_DOCI && nocc_up != nocc_dn)
        PYCI_FAIL(PYCI_ERR_VALUE, "nocc_up != nocc_dn");
    if (kind == PYCI_GENCI && nocc_dn != 0)
        PYCI_FAIL(PYCI_ERR_VALUE, "nocc_dn != 0");
    if (nbasis > 256)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED,
                  "nbasis = %ld: the device kernels handle nbasis <= 64 (fast paths) and <= 256 (multi-word path)", nbasis);
    return ctx_activate(ctx);
}

static pyci_wfn *wfn_new(pyci_ctx *ctx, int kind, long nbasis, long nocc_up, long nocc_dn, long ndet) {
    pyci_wfn *wfn = new pyci_wfn();
    wfn->ctx = ctx;
    wfn->kind = kind;
    wfn->nbasis = nbasis;
    wfn->nocc_up = nocc_up;
    wfn->nocc_dn = nocc_dn;
    wfn->ndet = ndet;
    wfn->nwords = ((kind == PYCI_FULLCI) ? 2 : 1) * (int)((nbasis + 63) / 64); // common.cpp:280-282
    if (nbasis > 64)
        wfn->keymode = KEY_MW; // multi-word strings: the generic slow path (multiword.cu)
    else if (kind == PYCI_FULLCI)
        wfn->keymode = (nbasis <= 16) ? KEY32 : (nbasis <= 32) ? KEY64 : KEY128;
    else
        wfn->keymode = (nbasis <= 32) ? KEY32 : KEY64;
    wfn->complete = (ndet == full_space_size(kind, nbasis, nocc_up, nocc_dn)); // uniqueness: verified by the index build
    return wfn;
}

int pyci_wfn_upload(pyci_ctx *ctx, int kind, long nbasis, long nocc_up, long nocc_dn, long ndet,
                    const uint64_t *dets, pyci_wfn **out) {
    if (!ctx || !out || (ndet > 0 && !dets))
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    *out = nullptr;
    PYCI_TRY(wfn_check_args(ctx, kind, nbasis, nocc_up, nocc_dn));
    if (ndet < 0 || ndet >= (1L << 31) - 1)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "ndet = %ld out of range for int32 column indices", ndet);
    // that every string holds the declared number of electrons inside nbasis orbitals is checked on the device,
    // together with uniqueness, by the verify pass of the index build
    pyci_wfn *wfn = wfn_new(ctx, kind, nbasis, nocc_up, nocc_dn, ndet);
    int rc = upload(&wfn->dets, (const u64 *)dets, (size_t)ndet * wfn->nwords, ctx->stream);
    if (rc == PYCI_OK)
        rc = wfn_build_index(wfn);
    if (rc != PYCI_OK) {
        pyci_wfn_destroy(wfn);
        return rc;
    }
    *out = wfn;
    return PYCI_OK;
}

int pyci_wfn_create_all_dets(pyci_ctx *ctx, int kind, long nbasis, long nocc_up, long nocc_dn, pyci_wfn **out) {
    if (!ctx || !out)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    *out = nullptr;
    PYCI_TRY(wfn_check_args(ctx, kind, nbasis, nocc_up, nocc_dn));
    if (nbasis > 64)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "determinants of more than 64 orbitals are not generated on the device: upload them");
    const long ndet = full_space_size(kind, nbasis, nocc_up, nocc_dn);
    if (ndet >= (1L << 31) - 1)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "ndet = %ld out of range for int32 column indices", ndet);
    pyci_wfn *wfn = wfn_new(ctx, kind, nbasis, nocc_up, nocc_dn, ndet);
    int rc = wfn_generate_all_dets(wfn, binom_l(nbasis, nocc_up), kind == PYCI_FULLCI ? binom_l(nbasis, nocc_dn) : 1);
    if (rc == PYCI_OK) {
        wfn->generated = true; // add_all_dets order and valid occupations by construction (wfn_build_index)
        rc = wfn_build_index(wfn);
    }
    if (rc != PYCI_OK) {
        pyci_wfn_destroy(wfn);
        return rc;
    }
    *out = wfn;
    return PYCI_OK;
}

void pyci_wfn_destroy(pyci_wfn *wfn) {
    if (!wfn)
        return;
    ctx_activate(wfn->ctx);
    dev_free(wfn->dets);
    dev_free(wfn->slots);
    dev_free(wfn->bloom);
    delete wfn;
}

int pyci_wfn_reindex(pyci_wfn *wfn) {
    if (!wfn)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    PYCI_TRY(ctx_activate(wfn->ctx));
    return wfn_build_index(wfn);
}

double pyci_wfn_index_seconds(const pyci_wfn *wfn) { return wfn ? wfn->hash_seconds : 0.0; }

long pyci_wfn_ndet(const pyci_wfn *wfn) { return wfn ? wfn->ndet : 0; }

int pyci_wfn_download_dets(const pyci_wfn *wfn, long start, long n, uint64_t *out) {
    if (!wfn || (n > 0 && !out))
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    if (start < 0 || n < 0 || start + n > wfn->ndet)
        PYCI_FAIL(PYCI_ERR_VALUE, "determinant range [%ld, %ld) out of bounds (ndet = %ld)", start, start + n, wfn->ndet);
    if (n == 0)
        return PYCI_OK;
    PYCI_TRY(ctx_activate(wfn->ctx));
    PYCI_CUDA(cudaMemcpyAsync(out, wfn->dets + start * wfn->nwords, sizeof(u64) * (size_t)(n * wfn->nwords),
                              cudaMemcpyDeviceToHost, wfn->ctx->stream));
    PYCI_CUDA(cudaStreamSynchronize(wfn->ctx->stream));
    return PYCI_OK;
}

int pyci_wfn_add_hci(pyci_ctx *ctx, const pyci_ham *ham, pyci_wfn *wfn, const double *coeffs, double eps, long *n_new) {
    if (!ctx || !ham || !wfn || !coeffs || !n_new)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    if (ham->nbasis != wfn->nbasis)
        PYCI_FAIL(PYCI_ERR_VALUE, "Hamiltonian has %ld basis functions, wave function %ld", ham->nbasis, wfn->nbasis);
    *n_new = 0;
    if (wfn->nbasis > 64)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "add_hci of wave functions with more than 64 orbitals is not on the device "
                                        "(multi-word determinants: sparse_op / matvec / solve only)");
    if (wfn->ndet == 0)
        return PYCI_OK;
    PYCI_TRY(ctx_activate(ctx));
    PYCI_TRY(wfn_ensure_index(wfn));
    PYCI_TRY(add_hci_impl(ctx, ham, wfn, coeffs, eps, n_new, &wfn->ext_seconds));
    if (*n_new > 0) {
        wfn->complete = (wfn->ndet == full_space_size(wfn->kind, wfn->nbasis, wfn->nocc_up, wfn->nocc_dn));
        PYCI_TRY(wfn_build_index(wfn));
    }
    return PYCI_OK;
}

double pyci_wfn_ext_seconds(const pyci_wfn *wfn) { return wfn ? wfn->ext_seconds : 0.0; }

int pyci_compute_enpt2(pyci_ctx *ctx, const pyci_ham *ham, pyci_wfn *wfn, const double *coeffs, double energy,
                       double eps, double *out, long *nterms) {
    if (!ctx || !ham || !wfn || !coeffs || !out)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    if (ham->nbasis != wfn->nbasis)
        PYCI_FAIL(PYCI_ERR_VALUE, "Hamiltonian has %ld basis functions, wave function %ld", ham->nbasis, wfn->nbasis);
    if (wfn->kind == PYCI_DOCI)
        PYCI_FAIL(PYCI_ERR_VALUE, "compute_enpt2 of a DOCI wave function runs on its FullCI image (enpt2.cpp:376-380): "
                                  "upload the determinants as (d, d) FullCI strings");
    if (wfn->nbasis > 64)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "compute_enpt2 of wave functions with more than 64 orbitals is not on the device "
                                        "(multi-word determinants: sparse_op / matvec / solve only)");
    PYCI_TRY(ctx_activate(ctx));
    PYCI_TRY(wfn_ensure_index(wfn));
    return enpt2_impl(ctx, ham, wfn, coeffs, energy, eps, out, nterms, &wfn->ext_seconds);
}

int pyci_wfn_index_dets(pyci_wfn *wfn, long n, const uint64_t *dets, long *out) {
    if (!wfn || (n > 0 && (!dets || !out)))
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    if (n <= 0)
        return PYCI_OK;
    pyci_ctx *ctx = wfn->ctx;
    if (wfn->nbasis > 64)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "index_det of wave functions with more than 64 orbitals is not on the device "
                                        "(multi-word determinants: sparse_op / matvec / solve only)");
    PYCI_TRY(ctx_activate(ctx));
    u64 *d = nullptr;
    long *o = nullptr;
    PYCI_TRY(upload(&d, (const u64 *)dets, (size_t)n * wfn->nwords, ctx->stream));
    cudaError_t e = dev_malloc(&o, sizeof(long) * n);
    if (e != cudaSuccess) {
        dev_free(d);
        PYCI_CUDA(e);
    }
    int rc = wfn_index_dets_impl(wfn, n, d, o);
    if (rc == PYCI_OK) {
        e = cudaMemcpyAsync(out, o, sizeof(long) * n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            pyci_set_error("CUDA error: %s", cudaGetErrorString(e));
            rc = PYCI_ERR_CUDA;
        }
    }
    dev_free(d);
    dev_free(o);
    return rc;
}

// ---- sparse operator -----------------------------------------------------------------------------

static int op_build_common(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, long nrow, long ncol, int symmetric,
                           int rank, int nranks, bool foreign, pyci_op **out) {
    if (!ctx || !ham || !wfn || !out)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    *out = nullptr;
    if (ham->nbasis != wfn->nbasis)
        PYCI_FAIL(PYCI_ERR_VALUE, "ham.nbasis (%ld) != wfn.nbasis (%ld)", ham->nbasis, wfn->nbasis);
    if (nrow < 0)
        nrow = wfn->ndet; // sparseop.cpp:53-54
    if (ncol < 0)
        ncol = wfn->ndet;
    if (nrow > wfn->ndet || ncol > wfn->ndet)
        PYCI_FAIL(PYCI_ERR_VALUE, "nrow/ncol (%ld, %ld) exceed the number of determinants (%ld)", nrow, ncol,
                  wfn->ndet);
    if (wfn->kind == PYCI_DOCI ? (!ham->h || !ham->v || !ham->w) : (!ham->one_mo || !ham->two_mo))
        PYCI_FAIL(PYCI_ERR_VALUE, "Hamiltonian lacks the integrals this wave-function kind needs");
    PYCI_TRY(ctx_activate(ctx));
    pyci_op *op = new pyci_op();
    op->ctx = ctx;
    op->nrow = nrow;
    op->ncol = ncol;
    op->symmetric = symmetric ? 1 : 0;
    op->ecore = ham->ecore;
    op->foreign = foreign;
    const long R = nranks;
    op->npad = std::max<long>(1, (std::max(nrow, ncol) + R - 1) / R);
    op->row0 = std::min(nrow, op->npad * rank);
    op->nloc = std::min(nrow, op->npad * (rank + 1)) - op->row0;
    int rc = PYCI_OK;
    auto alloc = [&](void **p, size_t bytes) {
        if (rc == PYCI_OK && dev_malloc(p, std::max<size_t>(bytes, 8)) != cudaSuccess) {
            pyci_set_error("device allocation of %zu bytes failed", bytes);
            cudaGetLastError();
            rc = PYCI_ERR_MEMORY;
        }
    };
    alloc((void **)&op->indptr, sizeof(long) * (size_t)(op->nloc + 1));
    alloc((void **)&op->lowcnt, sizeof(int) * (size_t)(op->nloc + 1));
    alloc((void **)&op->diag, sizeof(double) * (size_t)op->npad);
    if (rc == PYCI_OK) {
        cudaMemsetAsync(op->lowcnt, 0, sizeof(int) * (size_t)(op->nloc + 1), ctx->stream);
        cudaMemsetAsync(op->diag, 0, sizeof(double) * (size_t)op->npad, ctx->stream);
        rc = op_build_impl(ctx, ham, wfn, op);
    }
    // selected spaces: equal stored entries per rank instead of equal rows (rebalance.cu; collective)
    if (rc == PYCI_OK && !foreign && ctx->nranks > 1)
        rc = op_rebalance(ctx, op);
    if (rc != PYCI_OK) {
        pyci_op_destroy(op);
        return rc;
    }
    *out = op;
    return PYCI_OK;
}

int pyci_op_build(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, long nrow, long ncol, int symmetric,
                  pyci_op **out) {
    if (!ctx)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    return op_build_common(ctx, ham, wfn, nrow, ncol, symmetric, ctx->rank, ctx->nranks, false, out);
}

int pyci_op_build_shard(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, long nrow, long ncol, int symmetric,
                        int rank, int nranks, pyci_op **out) {
    if (!ctx)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks)
        PYCI_FAIL(PYCI_ERR_VALUE, "bad rank %d of %d", rank, nranks);
    const bool foreign = !(rank == ctx->rank && nranks == ctx->nranks);
    return op_build_common(ctx, ham, wfn, nrow, ncol, symmetric, rank, nranks, foreign, out);
}

int pyci_op_update(pyci_op *op, const pyci_ham *ham, const pyci_wfn *wfn) {
    if (!op || !ham || !wfn)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    pyci_ctx *ctx = op->ctx;
    if (ham->nbasis != wfn->nbasis)
        PYCI_FAIL(PYCI_ERR_VALUE, "ham.nbasis (%ld) != wfn.nbasis (%ld)", ham->nbasis, wfn->nbasis);
    if (wfn->ndet < op->nrow)
        PYCI_FAIL(PYCI_ERR_VALUE, "the wave function holds fewer determinants (%ld) than the operator has rows (%ld)",
                  wfn->ndet, op->nrow);
    // (row-sharded: the uniform row partition of the grown operator moves rows between ranks -- a rebuild; symmetric
    // with nrow != ncol: the device rows hold columns the transposed entries cannot be told from)
    if ((op->symmetric && op->nrow != op->ncol) || ctx->nranks != 1 || op->foreign || wfn->nbasis > 64)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "incremental update needs an operator on one rank (square if symmetric): rebuild instead");
    if (wfn->kind == PYCI_DOCI ? (!ham->h || !ham->v || !ham->w) : (!ham->one_mo || !ham->two_mo))
        PYCI_FAIL(PYCI_ERR_VALUE, "Hamiltonian lacks the integrals this wave-function kind needs");
    PYCI_TRY(ctx_activate(ctx));
    return op_update_impl(ctx, ham, wfn, op);
}

void pyci_op_destroy(pyci_op *op) {
    if (!op)
        return;
    ctx_activate(op->ctx);
    dev_free(op->indptr);
    dev_free(op->cols);
    dev_free(op->vals);
    dev_free(op->lowcnt);
    dev_free(op->diag);
    dev_free(op->xbuf);
    dev_free(op->ybuf);
    dev_free(op->spmv_part);
    dev_free(op->gather_stage);
    dev_free(op->bounds_dev);
    delete op;
}

long pyci_op_nrow(const pyci_op *op) { return op->nrow; }
long pyci_op_ncol(const pyci_op *op) { return op->ncol; }
long pyci_op_row_begin(const pyci_op *op) { return op->row0; }
long pyci_op_row_count(const pyci_op *op) { return op->nloc; }
long pyci_op_size(const pyci_op *op) { return op_size_ref(const_cast<pyci_op *>(op)); }
double pyci_op_fill_seconds(const pyci_op *op) { return op ? op->fill_seconds : 0.0; }
long pyci_op_stored_nnz(const pyci_op *op) { return op->nnz; }
double pyci_op_ecore(const pyci_op *op) { return op->ecore; }

const char *pyci_op_fill_kernel(const pyci_op *op) { return op ? op->fill_kernel : "none"; }
const char *pyci_op_count_kernel(const pyci_op *op) { return op ? op->count_kernel : "none"; }

int pyci_op_build_times(const pyci_op *op, double *seconds4) {
    for (int i = 0; i < 4; ++i)
        seconds4[i] = op->times[i];
    return PYCI_OK;
}

int pyci_op_export_csr(pyci_op *op, long *indptr, long *indices, double *data) {
    if (!op || !indptr)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    pyci_ctx *ctx = op->ctx;
    PYCI_TRY(ctx_activate(ctx));
    cudaStream_t st = ctx->stream;
    const long nloc = op->nloc;
    if (!op->symmetric && !indices && !data) {
        PYCI_CUDA(cudaMemcpyAsync(indptr, op->indptr, sizeof(long) * (nloc + 1), cudaMemcpyDeviceToHost, st));
        PYCI_CUDA(cudaStreamSynchronize(st));
        return PYCI_OK;
    }
    // output row pointer: the full row (general) or its col <= row prefix (symmetric, sparseop.cpp:223,262,431)
    long *outptr = nullptr, *didx = nullptr;
    double *dval = nullptr;
    auto body = [&]() -> int { // (every early return below leaves through the frees after the lambda)
    PYCI_CUDA(dev_malloc(&outptr, sizeof(long) * (size_t)(nloc + 1)));
    if (op->symmetric) {
        PYCI_TRY(scan_counts(ctx, op->lowcnt, nloc, outptr, nullptr)); // multi-block scan of the prefix counts
    } else {
        PYCI_CUDA(cudaMemcpyAsync(outptr, op->indptr, sizeof(long) * (nloc + 1), cudaMemcpyDeviceToDevice, st));
    }
    PYCI_CUDA(cudaMemcpyAsync(indptr, outptr, sizeof(long) * (nloc + 1), cudaMemcpyDeviceToHost, st));
    PYCI_CUDA(cudaStreamSynchronize(st));
    const long total = indptr[nloc];
    int rc = PYCI_OK;
    if ((indices || data) && total > 0) {
        // widen to the reference's int64 indices on the device in bounded chunks of rows
        const long chunk_entries = 1L << 26; // 64 Mi entries -> 1 GiB of staging
        const long cap = std::min(total, chunk_entries + (long)INT32_MAX / 2);
        long r0 = 0;
        while (r0 < nloc && rc == PYCI_OK) {
            long r1 = r0;
            while (r1 < nloc && indptr[r1 + 1] - indptr[r0] <= chunk_entries)
                ++r1;
            if (r1 == r0)
                r1 = r0 + 1; // a single very long row
            const long cnt = indptr[r1] - indptr[r0];
            if (cnt > 0) {
                if (!didx) {
                    const long c = std::max(cnt, std::min(cap, chunk_entries));
                    PYCI_CUDA(dev_malloc(&didx, sizeof(long) * (size_t)c));
                    PYCI_CUDA(dev_malloc(&dval, sizeof(double) * (size_t)c));
                }
                // outptr shifted so that this chunk starts at 0 of the staging buffers
                export_lower_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(op->indptr + r0, op->cols, op->vals, nullptr,
                                                                       outptr + r0, didx - indptr[r0],
                                                                       dval - indptr[r0], r1 - r0);
                ctx->launches++;
                if (indices)
                    PYCI_CUDA(cudaMemcpyAsync(indices + indptr[r0], didx, sizeof(long) * cnt, cudaMemcpyDeviceToHost, st));
                if (data)
                    PYCI_CUDA(cudaMemcpyAsync(data + indptr[r0], dval, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st));
                PYCI_CUDA(cudaStreamSynchronize(st));
            }
            r0 = r1;
        }
    }
    return rc;
    };
    const int rc_all = body();
    dev_free(didx);
    dev_free(dval);
    dev_free(outptr);
    return rc_all;
}

int pyci_op_export_rows(pyci_op *op, long nrows, const long *rows, long cap, long *indptr, long *indices, double *data) {
    if (!op || !indptr || (nrows > 0 && !rows) || nrows < 0)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    pyci_ctx *ctx = op->ctx;
    PYCI_TRY(ctx_activate(ctx));
    cudaStream_t st = ctx->stream;
    indptr[0] = 0;
    if (nrows == 0)
        return PYCI_OK;
    for (long k = 0; k < nrows; ++k)
        if (rows[k] < op->row0 || rows[k] >= op->row0 + op->nloc)
            PYCI_FAIL(PYCI_ERR_VALUE, "row %ld is not held by this rank (rows [%ld, %ld))", rows[k], op->row0,
                      op->row0 + op->nloc);
    long *drows = nullptr, *dsrc = nullptr, *dout = nullptr, *didx = nullptr;
    double *dval = nullptr;
    std::vector<long> hsrc((size_t)nrows), hcnt((size_t)nrows);
    auto body = [&]() -> int {
        PYCI_CUDA(dev_malloc(&drows, sizeof(long) * (size_t)nrows));
        PYCI_CUDA(dev_malloc(&dsrc, sizeof(long) * (size_t)nrows));
        PYCI_CUDA(dev_malloc(&dout, sizeof(long) * (size_t)(nrows + 1)));
        PYCI_CUDA(cudaMemcpyAsync(drows, rows, sizeof(long) * (size_t)nrows, cudaMemcpyHostToDevice, st));
        row_extent_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(op->indptr, op->symmetric ? op->lowcnt : nullptr,
                                                                          drows, op->row0, nrows, dsrc, dout);
        ctx->launches++;
        PYCI_CUDA(cudaMemcpyAsync(hsrc.data(), dsrc, sizeof(long) * (size_t)nrows, cudaMemcpyDeviceToHost, st));
        PYCI_CUDA(cudaMemcpyAsync(hcnt.data(), dout, sizeof(long) * (size_t)nrows, cudaMemcpyDeviceToHost, st));
        PYCI_CUDA(cudaStreamSynchronize(st));
        for (long k = 0; k < nrows; ++k)
            indptr[k + 1] = indptr[k] + hcnt[(size_t)k];
        const long total = indptr[nrows];
        if (!indices && !data)
            return PYCI_OK;
        if (total > cap)
            PYCI_FAIL(PYCI_ERR_VALUE, "the requested rows hold %ld entries, the output buffers %ld", total, cap);
        if (total == 0)
            return PYCI_OK;
        PYCI_CUDA(cudaMemcpyAsync(dout, indptr, sizeof(long) * (size_t)(nrows + 1), cudaMemcpyHostToDevice, st));
        PYCI_CUDA(dev_malloc(&didx, sizeof(long) * (size_t)total));
        PYCI_CUDA(dev_malloc(&dval, sizeof(double) * (size_t)total));
        export_rows_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(dsrc, op->cols, op->vals, dout, didx, dval, nrows);
        ctx->launches++;
        if (indices)
            PYCI_CUDA(cudaMemcpyAsync(indices, didx, sizeof(long) * (size_t)total, cudaMemcpyDeviceToHost, st));
        if (data)
            PYCI_CUDA(cudaMemcpyAsync(data, dval, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost, st));
        PYCI_CUDA(cudaStreamSynchronize(st));
        return PYCI_OK;
    };
    const int rc = body();
    dev_free(drows);
    dev_free(dsrc);
    dev_free(dout);
    dev_free(didx);
    dev_free(dval);
    return rc;
}

int pyci_op_matvec_dev(pyci_op *op, const double *x_dev, double *y_dev) {
    if (!op || !x_dev || !y_dev)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    PYCI_TRY(ctx_activate(op->ctx));
    if (op->symmetric && op->nrow != op->ncol)
        PYCI_FAIL(PYCI_ERR_TYPE, "symmetric operator must be square for matvec");
    return spmv_launch(op, x_dev, y_dev);
}

int pyci_op_matvec(pyci_op *op, const double *x, double *y) {
    if (!op || !x || !y)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    pyci_ctx *ctx = op->ctx;
    PYCI_TRY(ctx_activate(ctx));
    if (op->symmetric && op->nrow != op->ncol)
        PYCI_FAIL(PYCI_ERR_TYPE, "symmetric operator must be square for matvec");
    if (op->foreign)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "operator was built with pyci_op_build_shard for another rank layout: use "
                                        "pyci_op_matvec_dev on its rows");
    const long R = ctx->nranks;
    if (!op->xbuf)
        PYCI_CUDA(dev_malloc(&op->xbuf, sizeof(double) * (size_t)std::max<long>(op->ncol, 1)));
    if (!op->ybuf) {
        PYCI_CUDA(dev_malloc(&op->ybuf, sizeof(double) * (size_t)(op->npad * (R + 1))));
        PYCI_CUDA(cudaMemsetAsync(op->ybuf, 0, sizeof(double) * (size_t)(op->npad * (R + 1)), ctx->stream));
    }
    PYCI_CUDA(cudaMemcpyAsync(op->xbuf, x, sizeof(double) * op->ncol, cudaMemcpyHostToDevice, ctx->stream));
    double *yloc = op->ybuf + op->npad * R; // local shard, then gathered into ybuf[0 .. npad*R)
    PYCI_TRY(spmv_launch(op, op->xbuf, yloc));
    const double *ysrc = yloc;
    if (R > 1) {
        PYCI_TRY(op_allgather_rows(ctx, op, yloc, op->ybuf));
        ysrc = op->ybuf;
    }
    PYCI_CUDA(cudaMemcpyAsync(y, ysrc, sizeof(double) * op->nrow, cudaMemcpyDeviceToHost, ctx->stream));
    PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
    return PYCI_OK;
}

int pyci_op_set_spmv_shape(pyci_op *op, int threads_per_row, int ctas_per_sm) {
    if (!op)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    if (threads_per_row != 0 && threads_per_row != 1 && threads_per_row != 32 && threads_per_row != 64 &&
        threads_per_row != 128 && threads_per_row != 256 && threads_per_row != -8 && threads_per_row != -16 &&
        threads_per_row != -24 && threads_per_row != -32)
        PYCI_FAIL(PYCI_ERR_VALUE, "threads_per_row must be 0 (automatic), 1 (short-row kernel), 32, 64, 128, 256, or "
                                  "-8/-16/-24/-32 (bulk-copy stream kernel with that many warps per CTA)");
    if (ctas_per_sm < 1 || ctas_per_sm > 32)
        PYCI_FAIL(PYCI_ERR_VALUE, "ctas_per_sm must be in [1, 32]");
    op->spmv_tpr = threads_per_row;
    op->spmv_ctas = ctas_per_sm;
    return PYCI_OK;
}

int pyci_op_set_spmv_block(pyci_op *op, int block_threads, int depth) {
    if (!op)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    if (block_threads != 256 && block_threads != 512 && block_threads != 1024)
        PYCI_FAIL(PYCI_ERR_VALUE, "block_threads must be 256, 512 or 1024");
    if (depth < 2 || depth > 4)
        PYCI_FAIL(PYCI_ERR_VALUE, "depth must be 2, 3 or 4");
    op->spmv_block = block_threads;
    op->spmv_depth = depth;
    return PYCI_OK;
}

int pyci_op_time_spmv(pyci_op *op, int warmup, int reps, long flush_bytes, double *ms) {
    if (!op || !ms || reps < 1)
        PYCI_FAIL(PYCI_ERR_VALUE, "bad argument");
    pyci_ctx *ctx = op->ctx;
    PYCI_TRY(ctx_activate(ctx));
    cudaStream_t st = ctx->stream;
    double *x = nullptr, *y = nullptr;
    void *flush = nullptr;
    const long nx = std::max<long>(op->ncol, 1);
    std::vector<double> hx((size_t)nx);
    std::vector<cudaEvent_t> ev(2 * (size_t)reps, nullptr);
    auto setup = [&]() -> int { // (a failure here leaves through the clean-up at the end)
        PYCI_CUDA(dev_malloc(&x, sizeof(double) * nx));
        PYCI_CUDA(dev_malloc(&y, sizeof(double) * std::max<long>(op->nloc, 1)));
        if (flush_bytes > 0)
            PYCI_CUDA(dev_malloc(&flush, (size_t)flush_bytes));
        u64 sdd = 0x9e3779b97f4a7c15ULL;
        for (long i = 0; i < nx; ++i) {
            sdd ^= sdd << 13; sdd ^= sdd >> 7; sdd ^= sdd << 17;
            hx[i] = (double)(sdd >> 11) * (1.0 / 9007199254740992.0) - 0.5;
        }
        PYCI_CUDA(cudaMemcpyAsync(x, hx.data(), sizeof(double) * nx, cudaMemcpyHostToDevice, st));
        for (auto &e : ev)
            PYCI_CUDA(cudaEventCreate(&e));
        return PYCI_OK;
    };
    int rc = setup();
    for (int it = -warmup; it < reps && rc == PYCI_OK; ++it) {
        if (flush)
            cudaMemsetAsync(flush, it & 0xff, (size_t)flush_bytes, st);
        if (it >= 0)
            cudaEventRecord(ev[2 * it], st);
        rc = spmv_launch(op, x, y);
        if (it >= 0)
            cudaEventRecord(ev[2 * it + 1], st);
    }
    if (rc == PYCI_OK && cudaStreamSynchronize(st) != cudaSuccess) {
        pyci_set_error("CUDA error in SpMV timing: %s", cudaGetErrorString(cudaGetLastError()));
        rc = PYCI_ERR_CUDA;
    }
    for (int it = 0; it < reps && rc == PYCI_OK; ++it) {
        float t = 0;
        cudaEventElapsedTime(&t, ev[2 * it], ev[2 * it + 1]);
        ms[it] = t;
    }
    if (rc != PYCI_OK)
        cudaStreamSynchronize(st); // hx is read by the upload
    for (auto &e : ev)
        if (e)
            cudaEventDestroy(e);
    dev_free(x);
    dev_free(y);
    dev_free(flush);
    return rc;
}

int pyci_op_get_element(pyci_op *op, long i, long j, double *out) {
    if (!op || !out)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    pyci_ctx *ctx = op->ctx;
    PYCI_TRY(ctx_activate(ctx));
    if (i < op->row0 || i >= op->row0 + op->nloc)
        PYCI_FAIL(PYCI_ERR_VALUE, "row %ld is not held by this rank", i);
    *out = 0.0;
    // symmetric storage of the reference only holds j <= i (sparseop.cpp:89-94 searches the stored row)
    if (op->symmetric && j > i)
        return PYCI_OK;
    if (op->ge_row != i) { // one trip to the device per row, not per element
        op->ge_row = -1;
        long ptr[2];
        PYCI_CUDA(cudaMemcpyAsync(ptr, op->indptr + (i - op->row0), sizeof(long) * 2, cudaMemcpyDeviceToHost, ctx->stream));
        PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
        const long m = std::max<long>(ptr[1] - ptr[0], 0);
        op->ge_cols.resize((size_t)m);
        op->ge_vals.resize((size_t)m);
        if (m > 0) {
            PYCI_CUDA(cudaMemcpyAsync(op->ge_cols.data(), op->cols + ptr[0], sizeof(int) * m, cudaMemcpyDeviceToHost, ctx->stream));
            PYCI_CUDA(cudaMemcpyAsync(op->ge_vals.data(), op->vals + ptr[0], sizeof(double) * m, cudaMemcpyDeviceToHost, ctx->stream));
            PYCI_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        op->ge_row = i;
    }
    const std::vector<int> &cols = op->ge_cols;
    auto it = std::lower_bound(cols.begin(), cols.end(), (int)std::min<long>(j, INT32_MAX));
    if (it != cols.end() && *it == j)
        *out = op->ge_vals[(size_t)(it - cols.begin())];
    return PYCI_OK;
}

int pyci_op_solve(pyci_op *op, long n, const double *c0, long ncv, long maxiter, double tol, double *evals,
                  double *evecs, pyci_solve_stats *stats) {
    if (!op || !evals || !evecs)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    pyci_ctx *ctx = op->ctx;
    PYCI_TRY(ctx_activate(ctx));
    if (stats)
        memset(stats, 0, sizeof(*stats));
    if (op->foreign)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "operator was built with pyci_op_build_shard for another rank layout");
    const long nrow = op->nrow;
    // guards of SparseOp::solve_ci, sparseop.cpp:116-124
    if (n < 1 || (nrow > 1 && n >= nrow) || (nrow == 1 && n > 1))
        PYCI_FAIL(PYCI_ERR_VALUE, "cannot find >=n eigenpairs for sparse operator with n rows");
    if (nrow != op->ncol)
        PYCI_FAIL(PYCI_ERR_TYPE, "Can only solve sparse symmetric matrix operators");
    if (nrow == 1) {
        double h00 = 0.0;
        cudaStream_t st = ctx->stream;
        if (ctx->rank == 0) {
            PYCI_CUDA(cudaMemcpyAsync(&h00, op->diag, sizeof(double), cudaMemcpyDeviceToHost, st));
            PYCI_CUDA(cudaStreamSynchronize(st));
        }
        if (ctx->nranks > 1) {
            // every rank must return the same value; rank 0 holds row 0
            double *d = nullptr;
            PYCI_CUDA(dev_malloc(&d, sizeof(double)));
            PYCI_CUDA(cudaMemcpyAsync(d, &h00, sizeof(double), cudaMemcpyHostToDevice, st));
            PYCI_TRY(comm_allreduce_sum_f64(ctx, d, 1));
            PYCI_CUDA(cudaMemcpyAsync(&h00, d, sizeof(double), cudaMemcpyDeviceToHost, st));
            PYCI_CUDA(cudaStreamSynchronize(st));
            dev_free(d);
        }
        evals[0] = h00 + op->ecore;
        evecs[0] = 1.0;
        return PYCI_OK;
    }
    return solve_impl(op, n, c0, ncv, maxiter, tol, evals, evecs, stats);
}

int pyci_compute_rdms(pyci_ctx *ctx, const pyci_wfn *wfn, const double *coeffs, double *rdm1, double *rdm2) {
    if (!ctx || !wfn || !coeffs || !rdm1 || !rdm2)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    if (wfn->nbasis > 64)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "compute_rdms of wave functions with more than 64 orbitals is not on the device "
                                        "(multi-word determinants: sparse_op / matvec / solve only)");
    PYCI_TRY(ctx_activate(ctx));
    PYCI_TRY(wfn_ensure_index(wfn));
    return rdms_impl(ctx, wfn, coeffs, rdm1, rdm2);
}

int pyci_compute_transition_rdms(pyci_ctx *ctx, const pyci_wfn *wfn1, const pyci_wfn *wfn2, const double *coeffs1,
                                 const double *coeffs2, double *rdm1, double *rdm2) {
    if (!ctx || !wfn1 || !wfn2 || !coeffs1 || !coeffs2 || !rdm1 || !rdm2)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    if (wfn1->nbasis > 64)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "compute_transition_rdms of wave functions with more than 64 orbitals is not on the device "
                                        "(multi-word determinants: sparse_op / matvec / solve only)");
    PYCI_TRY(ctx_activate(ctx));
    PYCI_TRY(wfn_ensure_index(wfn2));
    return trdms_impl(ctx, wfn1, wfn2, coeffs1, coeffs2, rdm1, rdm2);
}

int pyci_compute_overlap(pyci_ctx *ctx, const pyci_wfn *wfn1, const pyci_wfn *wfn2, const double *coeffs1,
                         const double *coeffs2, double *out) {
    if (!ctx || !wfn1 || !wfn2 || !coeffs1 || !coeffs2 || !out)
        PYCI_FAIL(PYCI_ERR_VALUE, "null argument");
    if (wfn1->nbasis > 64)
        PYCI_FAIL(PYCI_ERR_UNSUPPORTED, "compute_overlap of wave functions with more than 64 orbitals is not on the device "
                                        "(multi-word determinants: sparse_op / matvec / solve only)");
    PYCI_TRY(ctx_activate(ctx));
    PYCI_TRY(wfn_ensure_index(wfn2));
    return overlap_impl(ctx, wfn1, wfn2, coeffs1, coeffs2, out);
}

} // extern "C"
