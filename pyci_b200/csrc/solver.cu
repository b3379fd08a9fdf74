// Lowest eigenpairs of the CI matrix on the device: the B200 replacement for SparseOp::solve_ci
// (/root/reference/pyci/src/sparseop.cpp:114-146), which hands the matrix to Spectra's
// implicitly-restarted Lanczos.  Here: block Davidson with the diagonal preconditioner, the subspace
// vectors row-sharded like the matrix, one SpMV (spmv.cu) per new vector, and -- when the rows are
// sharded over several GPUs -- one NCCL all-gather of the trial vector per SpMV plus small
// all-reduces of the projected quantities.  The small projected eigenproblem (subspace dimension
// <= ncv) is diagonalised on the host with cyclic Jacobi.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "common.cuh"

namespace {

constexpr int RB = 256; // threads per block of the vector kernels
constexpr int KCHUNK = 8;

// out[j] += sum_i V[j*ld + i] * w[i], j < k <= KCHUNK
__global__ void __launch_bounds__(RB) multi_dot_kernel(const double *__restrict__ V, long ld, int k,
                                                       const double *__restrict__ w, long n,
                                                       double *__restrict__ out) {
    double acc[KCHUNK];
#pragma unroll
    for (int j = 0; j < KCHUNK; ++j)
        acc[j] = 0.0;
    for (long i = (long)blockIdx.x * RB + threadIdx.x; i < n; i += (long)gridDim.x * RB) {
        const double wi = w[i];
#pragma unroll
        for (int j = 0; j < KCHUNK; ++j)
            if (j < k)
                acc[j] = fma(V[j * ld + i], wi, acc[j]);
    }
    __shared__ double ws[KCHUNK][RB / 32];
#pragma unroll
    for (int j = 0; j < KCHUNK; ++j) {
        double a = acc[j];
        for (int o = 16; o > 0; o >>= 1)
            a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((threadIdx.x & 31) == 0)
            ws[j][threadIdx.x >> 5] = a;
    }
    __syncthreads();
    if (threadIdx.x < k) {
        double a = 0.0;
        for (int q = 0; q < RB / 32; ++q)
            a += ws[threadIdx.x][q];
        atomicAdd(out + threadIdx.x, a);
    }
}

// y[i] = beta*x[i] + alpha * sum_j s[j] V[j*ld+i]   (y may be x)
__global__ void __launch_bounds__(RB) combine_kernel(const double *__restrict__ V, long ld, int k,
                                                     const double *__restrict__ s, double alpha, double beta,
                                                     const double *x, double *y, long n) {
    extern __shared__ double sh[];
    for (int j = threadIdx.x; j < k; j += RB)
        sh[j] = s[j];
    __syncthreads();
    for (long i = (long)blockIdx.x * RB + threadIdx.x; i < n; i += (long)gridDim.x * RB) {
        double a = 0.0;
        for (int j = 0; j < k; ++j)
            a = fma(sh[j], V[j * ld + i], a);
        y[i] = (beta == 0.0 ? 0.0 : beta * x[i]) + alpha * a;
    }
}

// Ritz vector, its image, residual and Davidson correction in one pass:
//   x = V s, ax = W s, r = ax - theta x, t = r / (theta - d)
//   sums[0] += |r|^2, sums[1] += x . (theta - D)^-1 r, sums[2] += x . (theta - D)^-1 x   (the last two for the Olsen
//   form of the correction, olsen_kernel)
__device__ __forceinline__ double davidson_den(double theta, double d) {
    double den = theta - d;
    if (fabs(den) < 1.0e-8)
        den = (den < 0.0) ? -1.0e-8 : 1.0e-8;
    return den;
}

__global__ void __launch_bounds__(RB) ritz_residual_kernel(const double *__restrict__ V, const double *__restrict__ W,
                                                           long ld, int k, const double *__restrict__ s, double theta,
                                                           const double *__restrict__ diag, double *__restrict__ x,
                                                           double *__restrict__ ax, double *__restrict__ t, long n,
                                                           double *__restrict__ sums) {
    extern __shared__ double sh[];
    for (int j = threadIdx.x; j < k; j += RB)
        sh[j] = s[j];
    __syncthreads();
    double acc[3] = {0.0, 0.0, 0.0};
    for (long i = (long)blockIdx.x * RB + threadIdx.x; i < n; i += (long)gridDim.x * RB) {
        double xv = 0.0, av = 0.0;
        for (int j = 0; j < k; ++j) {
            xv = fma(sh[j], V[j * ld + i], xv);
            av = fma(sh[j], W[j * ld + i], av);
        }
        const double r = av - theta * xv;
        const double den = davidson_den(theta, diag[i]);
        x[i] = xv;
        ax[i] = av;
        t[i] = r / den;
        acc[0] = fma(r, r, acc[0]);
        acc[1] = fma(xv, r / den, acc[1]);
        acc[2] = fma(xv, xv / den, acc[2]);
    }
    __shared__ double ws[3][RB / 32];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double a = acc[q];
        for (int o = 16; o > 0; o >>= 1)
            a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((threadIdx.x & 31) == 0)
            ws[q][threadIdx.x >> 5] = a;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double a = 0.0;
        for (int q = 0; q < RB / 32; ++q)
            a += ws[threadIdx.x][q];
        atomicAdd(sums + threadIdx.x, a);
    }
}

// Olsen's correction: t = (theta - D)^-1 (r - eps x) with eps = [x.(theta-D)^-1 r] / [x.(theta-D)^-1 x], i.e. the
// diagonal-preconditioned residual made orthogonal to the Ritz vector: t -= eps x / (theta - d)
__global__ void __launch_bounds__(RB) olsen_kernel(double *__restrict__ t, const double *__restrict__ x,
                                                   const double *__restrict__ diag, double theta, double eps, long n) {
    for (long i = (long)blockIdx.x * RB + threadIdx.x; i < n; i += (long)gridDim.x * RB)
        t[i] -= eps * x[i] / davidson_den(theta, diag[i]);
}

// Thick restart, in place: V_r <- sum_j Z[j][r] V_j for r < kk, vectors V_j = V + j * ld, j < m <= ROT_MAXM.  Every
// thread reads the m values of its element before it writes the kk <= m new ones.
constexpr int ROT_MAXM = 64;
__global__ void __launch_bounds__(RB) rotate_kernel(double *__restrict__ V, long ld, int m, int kk,
                                                    const double *__restrict__ Z, long n) {
    extern __shared__ double sh[]; // Z[m][m], row j = coefficients of V_j
    for (int t = threadIdx.x; t < m * m; t += RB)
        sh[t] = Z[t];
    __syncthreads();
    double v[ROT_MAXM];
    for (long i = (long)blockIdx.x * RB + threadIdx.x; i < n; i += (long)gridDim.x * RB) {
        for (int j = 0; j < m; ++j)
            v[j] = V[j * ld + i];
        for (int r = 0; r < kk; ++r) {
            double a = 0.0;
            for (int j = 0; j < m; ++j)
                a = fma(sh[j * m + r], v[j], a);
            V[r * ld + i] = a;
        }
    }
}

__global__ void scale_kernel(double *x, double alpha, long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        x[i] *= alpha;
}

// deterministic start vector: unit vector at `hot` plus small hash noise so that no symmetry sector
// of the Hamiltonian is excluded from the search (Spectra starts from a random vector)
__global__ void guess_kernel(double *v, long row0, long nloc, long nrow, long hot, double eps, u32 seed) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += (long)gridDim.x * blockDim.x) {
        const long g = row0 + i;
        double val = 0.0;
        if (g < nrow) {
            const u32 hsh = mix64((u64)g * 0x9e3779b97f4a7c15ULL + seed);
            val = eps * ((double)hsh * (2.0 / 4294967296.0) - 1.0);
            if (g == hot)
                val += 1.0;
        }
        v[i] = val;
    }
}

// cyclic Jacobi for a dense symmetric m x m matrix (row-major a, destroyed); eigenvalues ascending in w,
// eigenvectors in the COLUMNS of z (row-major, z[i*m + j] = component i of vector j)
void jacobi_eigh(std::vector<double> &a, int m, std::vector<double> &w, std::vector<double> &z) {
    z.assign((size_t)m * m, 0.0);
    for (int i = 0; i < m; ++i)
        z[(size_t)i * m + i] = 1.0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0, dsum = 0.0;
        for (int i = 0; i < m; ++i) {
            dsum += a[(size_t)i * m + i] * a[(size_t)i * m + i];
            for (int j = i + 1; j < m; ++j)
                off += a[(size_t)i * m + j] * a[(size_t)i * m + j];
        }
        if (off <= 1.0e-60 || off <= 1.0e-34 * dsum)
            break;
        for (int p = 0; p < m - 1; ++p)
            for (int q = p + 1; q < m; ++q) {
                const double apq = a[(size_t)p * m + q];
                if (apq == 0.0)
                    continue;
                const double app = a[(size_t)p * m + p], aqq = a[(size_t)q * m + q];
                const double tau = (aqq - app) / (2.0 * apq);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
                for (int k = 0; k < m; ++k) {
                    const double akp = a[(size_t)k * m + p], akq = a[(size_t)k * m + q];
                    a[(size_t)k * m + p] = c * akp - s * akq;
                    a[(size_t)k * m + q] = s * akp + c * akq;
                }
                for (int k = 0; k < m; ++k) {
                    const double apk = a[(size_t)p * m + k], aqk = a[(size_t)q * m + k];
                    a[(size_t)p * m + k] = c * apk - s * aqk;
                    a[(size_t)q * m + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < m; ++k) {
                    const double zkp = z[(size_t)k * m + p], zkq = z[(size_t)k * m + q];
                    z[(size_t)k * m + p] = c * zkp - s * zkq;
                    z[(size_t)k * m + q] = s * zkp + c * zkq;
                }
            }
    }
    std::vector<int> order(m);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int x, int y) { return a[(size_t)x * m + x] < a[(size_t)y * m + y]; });
    w.resize(m);
    std::vector<double> zs((size_t)m * m);
    for (int j = 0; j < m; ++j) {
        w[j] = a[(size_t)order[j] * m + order[j]];
        for (int i = 0; i < m; ++i)
            zs[(size_t)i * m + j] = z[(size_t)i * m + order[j]];
    }
    z.swap(zs);
}

struct Solver {
    pyci_op *op;
    pyci_ctx *ctx;
    cudaStream_t st;
    long nrow, nloc, ld; // ld = npad
    int R;
    int grid;
    double *V = nullptr, *W = nullptr, *X = nullptr, *AX = nullptr, *T = nullptr;
    double *xfull = nullptr; // all-gather target (R > 1)
    double *dsmall = nullptr; // device scratch for dots / coefficients
    double *dZ = nullptr, *hZ = nullptr; // projected eigenvectors of a thick restart
    double *hsmall = nullptr; // pinned host mirror (results of dots)
    double *hring = nullptr, *dring = nullptr; // coefficient staging ring: RING slots of small_cap doubles, host pinned
                                               // + device, so that uploads need no synchronisation before reuse
    int ring_pos = 0, ring_pending = 0; // uploads since the last stream synchronisation
    int small_cap = 0;
    static constexpr int RING = 32;
    pyci_solve_stats stats;
    cudaEvent_t e0 = nullptr, e1 = nullptr;

    ~Solver() {
        dev_free(V);
        dev_free(W);
        dev_free(X);
        dev_free(AX);
        dev_free(T);
        dev_free(xfull);
        dev_free(dsmall);
        dev_free(dZ);
        if (hZ)
            cudaFreeHost(hZ);
        if (hsmall)
            cudaFreeHost(hsmall);
        if (hring)
            cudaFreeHost(hring);
        dev_free(dring);
        if (e0)
            cudaEventDestroy(e0);
        if (e1)
            cudaEventDestroy(e1);
    }

    // w = A v   (v, w: local shards)
    int apply(const double *v, double *w) {
        PYCI_CUDA(cudaEventRecord(e0, st));
        const double *xin = v;
        if (R > 1) {
            PYCI_NVTX("pyci:allgather(trial vector)");
            PYCI_TRY(op_allgather_rows(ctx, op, v, xfull));
            xin = xfull;
        }
        PYCI_TRY(spmv_launch(op, xin, w));
        PYCI_CUDA(cudaEventRecord(e1, st));
        stats.matvecs++;
        return PYCI_OK;
    }

    // host[j] = V_j . w for j < k (all-reduced over ranks)
    int dots(const double *Vb, int k, const double *w, double *host) {
        for (int j0 = 0; j0 < k; j0 += small_cap) {
            const int kk = std::min(k - j0, small_cap);
            PYCI_CUDA(cudaMemsetAsync(dsmall, 0, sizeof(double) * kk, st));
            for (int c0 = 0; c0 < kk; c0 += KCHUNK) {
                const int kc = std::min(KCHUNK, kk - c0);
                multi_dot_kernel<<<grid, RB, 0, st>>>(Vb + (long)(j0 + c0) * ld, ld, kc, w, nloc, dsmall + c0);
                ctx->launches++;
            }
            if (R > 1)
                PYCI_TRY(comm_allreduce_sum_f64(ctx, dsmall, kk));
            PYCI_CUDA(cudaMemcpyAsync(hsmall, dsmall, sizeof(double) * kk, cudaMemcpyDeviceToHost, st));
            PYCI_CUDA(cudaStreamSynchronize(st));
            ring_pending = 0;
            std::memcpy(host + j0, hsmall, sizeof(double) * kk);
        }
        return PYCI_OK;
    }

    // host[0..k) = Vb_j . w, host[k] = extra . w: one reduction, one synchronisation
    int dots2(const double *Vb, int k, const double *extra, const double *w, double *host) {
        if (k + 1 > small_cap)
            PYCI_FAIL(PYCI_ERR_RUNTIME, "subspace larger than scratch");
        PYCI_CUDA(cudaMemsetAsync(dsmall, 0, sizeof(double) * (k + 1), st));
        for (int c0 = 0; c0 < k; c0 += KCHUNK) {
            const int kc = std::min(KCHUNK, k - c0);
            multi_dot_kernel<<<grid, RB, 0, st>>>(Vb + (long)c0 * ld, ld, kc, w, nloc, dsmall + c0);
            ctx->launches++;
        }
        multi_dot_kernel<<<grid, RB, 0, st>>>(extra, ld, 1, w, nloc, dsmall + k);
        ctx->launches++;
        if (R > 1)
            PYCI_TRY(comm_allreduce_sum_f64(ctx, dsmall, k + 1));
        PYCI_CUDA(cudaMemcpyAsync(hsmall, dsmall, sizeof(double) * (k + 1), cudaMemcpyDeviceToHost, st));
        PYCI_CUDA(cudaStreamSynchronize(st));
        ring_pending = 0;
        std::memcpy(host, hsmall, sizeof(double) * (k + 1));
        return PYCI_OK;
    }

    // coefficients to the device through the staging ring; a slot is reused after RING uploads, and every Davidson
    // iteration synchronises the stream at least twice, so a slot is never overwritten while its copy is pending
    const double *stage(const double *s_host, int k, int *err) {
        *err = PYCI_OK;
        if (++ring_pending >= RING) { // many roots: do not lap a slot whose copy may still be pending
            cudaStreamSynchronize(st);
            ring_pending = 0;
        }
        double *h = hring + (size_t)ring_pos * small_cap, *d = dring + (size_t)ring_pos * small_cap;
        ring_pos = (ring_pos + 1) % RING;
        std::memcpy(h, s_host, sizeof(double) * k);
        if (cudaMemcpyAsync(d, h, sizeof(double) * k, cudaMemcpyHostToDevice, st) != cudaSuccess) {
            pyci_set_error("CUDA error while staging coefficients");
            *err = PYCI_ERR_CUDA;
        }
        return d;
    }

    // y = beta x + alpha * sum_j s_j Vb_j   (asynchronous; y = nullptr: in place)
    int combine(const double *Vb, int k, const double *s_host, double alpha, double beta, double *x, double *y = nullptr) {
        if (k > small_cap)
            PYCI_FAIL(PYCI_ERR_RUNTIME, "subspace larger than scratch");
        int err;
        const double *d = stage(s_host, k, &err);
        PYCI_TRY(err);
        combine_kernel<<<grid, RB, sizeof(double) * (size_t)(k + 2), st>>>(Vb, ld, k, d, alpha, beta, x, y ? y : x, nloc);
        ctx->launches++;
        return PYCI_OK;
    }

    // Orthogonalise t against V[0..m) (classical Gram-Schmidt, twice) and normalise.  Each pass takes ONE batch
    // of dot products, [V_0..V_{m-1}, t] . t, i.e. one reduction and one synchronisation; V is orthonormal, so
    // the norm after the second pass is |t|^2 - sum_j (V_j . t)^2 and the scaling rides on the second update.
    // *rel = norm of the orthogonal component relative to the input norm.
    // "Twice is enough", and once is when little was removed (Daniel-Gragg-Kaufman-Stewart): if the first pass leaves
    // more than 1/sqrt(2) of the vector, what rounding puts back along V is at the level of the unit round-off and the
    // second pass -- a second reading of the whole basis -- is skipped; the scaling then rides on the first update.
    // (On short-row operators the basis passes of a Davidson step are a third of its time: config 5.)
    // dst != nullptr: the orthonormal vector is written there (the next basis vector's place) by the last update
    // instead of being copied afterwards; t is then scratch.
    int orthonormalize(double *t, int m, std::vector<double> &tmp, double *rel, double *dst = nullptr) {
        tmp.resize((size_t)m + 1);
        double n0 = 0.0, n1 = 0.0;
        static const bool dgks = getenv("PYCI_B200_SOLVER_GS2") == nullptr;
        for (int pass = 0; pass < 2; ++pass) {
            // t sits at V + m * ld when it is the next basis vector; in general it is a separate buffer: two batches
            if (m > 0)
                PYCI_TRY(dots2(V, m, t, t, tmp.data()));
            else
                PYCI_TRY(dots(t, 1, t, tmp.data()));
            const double tt = tmp[(size_t)m];
            if (pass == 0) {
                n0 = tt;
                if (!(n0 > 0.0)) {
                    *rel = 0.0;
                    return PYCI_OK;
                }
                if (m > 0) {
                    double proj = 0.0;
                    for (int j = 0; j < m; ++j)
                        proj += tmp[(size_t)j] * tmp[(size_t)j];
                    const double left = n0 - proj;
                    if (dgks && left > 0.5 * n0) {
                        const double inv = 1.0 / std::sqrt(left);
                        PYCI_TRY(combine(V, m, tmp.data(), -inv, inv, t, dst));
                        *rel = std::sqrt(left / n0);
                        return PYCI_OK;
                    }
                    PYCI_TRY(combine(V, m, tmp.data(), -1.0, 1.0, t));
                }
            } else {
                double proj = 0.0;
                for (int j = 0; j < m; ++j)
                    proj += tmp[(size_t)j] * tmp[(size_t)j];
                n1 = tt - proj;
                if (!(n1 > 0.0)) {
                    *rel = 0.0;
                    return PYCI_OK;
                }
                const double inv = 1.0 / std::sqrt(n1);
                if (m > 0) {
                    PYCI_TRY(combine(V, m, tmp.data(), -inv, inv, t, dst));
                } else {
                    scale_kernel<<<grid, RB, 0, st>>>(t, inv, nloc);
                    ctx->launches++;
                    if (dst)
                        PYCI_CUDA(cudaMemcpyAsync(dst, t, sizeof(double) * (size_t)ld, cudaMemcpyDeviceToDevice, st));
                }
            }
        }
        *rel = std::sqrt(n1 / n0);
        return PYCI_OK;
    }
};

} // namespace

int solve_impl(pyci_op *op, long n, const double *c0, long ncv, long maxiter, double tol, double *evals,
               double *evecs, pyci_solve_stats *stats_out) {
    PYCI_NVTX("pyci:solve(davidson)");
    pyci_ctx *ctx = op->ctx;
    const long nrow = op->nrow;
    const int R = ctx->nranks;
    // ncv = -1: the reference's default is max(2 n + 1, 20) Lanczos vectors (sparseop.cpp:129-130).  A Davidson step
    // reads its basis about five times (projection, Ritz vector + residual, orthogonalisation), which is free beside
    // a product over rows of thousands of entries and a third of the step over rows of ~200 (config 5): there the
    // thick restart converges in the same number of products with half the basis (5 M determinants, rows of 158:
    // 516 / 479 / 489 / 482 / 533 products and 1.63 / 1.46 / 1.43 / 1.38 / 1.48 s at 20 / 16 / 12 / 10 / 8 vectors).
    if (ncv == -1) {
        long tot[2] = {op->nnz, op->nloc}; // (the same decision on every rank: the ranks' rows differ in length)
        if (R > 1)
            PYCI_TRY(comm_allreduce_sum_i64_host(ctx, tot, 2));
        const long avg_row = tot[0] / std::max<long>(tot[1], 1);
        ncv = std::min(nrow, avg_row < 512 ? std::max(3 * n + 7, 10L) : std::max(2 * n + 1, 20L));
    }
    if (maxiter == -1)
        maxiter = n * nrow * 10;
    if (ncv <= n || ncv > nrow)
        PYCI_FAIL(PYCI_ERR_VALUE, "ncv must satisfy n < ncv <= nrow (n=%ld, ncv=%ld, nrow=%ld)", n, ncv, nrow);
    // Davidson needs room for the n Ritz vectors plus at least one correction per root
    const int mmax = (int)std::min<long>(nrow, std::max<long>(ncv, 2 * n));
    const int nroot = (int)n;

    Solver S;
    S.op = op;
    S.ctx = ctx;
    S.st = ctx->stream;
    S.nrow = nrow;
    S.nloc = op->nloc;
    S.ld = op->npad;
    S.R = R;
    S.grid = (int)std::max<long>(1, std::min<long>((S.nloc + RB - 1) / RB, (long)ctx->sm_count * 8));
    S.small_cap = std::max(mmax + 3 * nroot, 64);
    std::memset(&S.stats, 0, sizeof(S.stats));
    const size_t vec = sizeof(double) * (size_t)S.ld;
    PYCI_CUDA(dev_malloc(&S.V, vec * mmax));
    PYCI_CUDA(dev_malloc(&S.W, vec * mmax));
    PYCI_CUDA(dev_malloc(&S.X, vec * nroot));
    PYCI_CUDA(dev_malloc(&S.AX, vec * nroot));
    PYCI_CUDA(dev_malloc(&S.T, vec * nroot));
    PYCI_CUDA(cudaMemsetAsync(S.V, 0, vec * mmax, S.st));
    PYCI_CUDA(cudaMemsetAsync(S.W, 0, vec * mmax, S.st));
    PYCI_CUDA(cudaMemsetAsync(S.X, 0, vec * nroot, S.st));
    PYCI_CUDA(cudaMemsetAsync(S.AX, 0, vec * nroot, S.st));
    PYCI_CUDA(cudaMemsetAsync(S.T, 0, vec * nroot, S.st));
    PYCI_CUDA(dev_malloc(&S.xfull, vec * R));
    PYCI_CUDA(dev_malloc(&S.dsmall, sizeof(double) * S.small_cap));
    PYCI_CUDA(cudaMallocHost(&S.hsmall, sizeof(double) * S.small_cap));
    PYCI_CUDA(dev_malloc(&S.dZ, sizeof(double) * (size_t)mmax * mmax));
    PYCI_CUDA(cudaMallocHost(&S.hZ, sizeof(double) * (size_t)mmax * mmax));
    PYCI_CUDA(cudaMallocHost(&S.hring, sizeof(double) * S.small_cap * Solver::RING));
    PYCI_CUDA(dev_malloc(&S.dring, sizeof(double) * S.small_cap * Solver::RING));
    PYCI_CUDA(cudaEventCreate(&S.e0));
    PYCI_CUDA(cudaEventCreate(&S.e1));
    struct EventGuard { // destroyed on every return path
        cudaEvent_t e = nullptr;
        ~EventGuard() {
            if (e)
                cudaEventDestroy(e);
        }
    } guard_begin, guard_end;
    PYCI_CUDA(cudaEventCreate(&guard_begin.e));
    PYCI_CUDA(cudaEventCreate(&guard_end.e));
    const cudaEvent_t t_begin = guard_begin.e, t_end = guard_end.e;
    PYCI_CUDA(cudaEventRecord(t_begin, S.st));

    // ---- start vectors
    std::vector<double> hdiag((size_t)S.ld * R);
    {
        const double *dsrc = op->diag;
        if (R > 1) {
            PYCI_TRY(op_allgather_rows(ctx, op, op->diag, S.xfull));
            dsrc = S.xfull;
        }
        PYCI_CUDA(cudaMemcpyAsync(hdiag.data(), dsrc, vec * R, cudaMemcpyDeviceToHost, S.st));
        PYCI_CUDA(cudaStreamSynchronize(S.st));
    }
    std::vector<long> order(nrow);
    std::iota(order.begin(), order.end(), 0L);
    const long nguess = std::min<long>(nrow, nroot);
    std::partial_sort(order.begin(), order.begin() + nguess, order.end(), [&](long a, long b) {
        return hdiag[a] < hdiag[b] || (hdiag[a] == hdiag[b] && a < b);
    });

    int m = 0; // subspace dimension
    std::vector<double> tmp;
    for (int j = 0; j < nroot; ++j) {
        double *t = S.T;
        if (j == 0 && c0 != nullptr) {
            PYCI_CUDA(cudaMemsetAsync(t, 0, vec, S.st));
            PYCI_CUDA(cudaMemcpyAsync(t, c0 + op->row0, sizeof(double) * S.nloc, cudaMemcpyHostToDevice, S.st));
        } else {
            guess_kernel<<<S.grid, RB, 0, S.st>>>(t, op->row0, S.nloc, nrow, order[j], 1.0e-3, 0x5eedu + 77u * j);
            ctx->launches++;
        }
        double rel = 0.0;
        PYCI_TRY(S.orthonormalize(t, m, tmp, &rel));
        if (rel < 1.0e-12) { // degenerate start vector (e.g. c0 == 0): fall back to the unit-vector guess
            guess_kernel<<<S.grid, RB, 0, S.st>>>(t, op->row0, S.nloc, nrow, order[j], 1.0e-3, 0xabcdu + 131u * j);
            ctx->launches++;
            PYCI_TRY(S.orthonormalize(t, m, tmp, &rel));
        }
        PYCI_CUDA(cudaMemcpyAsync(S.V + (size_t)m * S.ld, t, vec, cudaMemcpyDeviceToDevice, S.st));
        ++m;
    }
    int nnew = m;

    std::vector<double> G((size_t)mmax * mmax, 0.0); // projected matrix V^T A V (row-major, leading dim mmax)
    std::vector<double> theta, Z, Gw, col(mmax);
    std::vector<double> rn(nroot, 0.0);
    std::vector<char> conv(nroot, 0);
    const double eps23 = std::pow(2.220446049250313e-16, 2.0 / 3.0);
    const bool olsen = !getenv("PYCI_B200_NO_OLSEN");
    // Ritz vectors kept by a restart: half the subspace (PYCI_B200_SOLVER_KEEP overrides; 0 = the wanted ones only)
    int keep_target = std::max(nroot, mmax / 2);
    if (const char *e = getenv("PYCI_B200_SOLVER_KEEP"))
        keep_target = std::max(nroot, std::min(mmax - nroot, atoi(e)));
    keep_target = std::min(keep_target, mmax - nroot);
    bool done = false;
    double spmv_ms = 0.0;
    long iter = 0;
    // PYCI_B200_SOLVER_TRACE: host wall time per phase (every phase ends in a stream synchronisation)
    const bool trace = getenv("PYCI_B200_SOLVER_TRACE") != nullptr;
    double tr[4] = {0, 0, 0, 0};
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    for (; iter < maxiter; ++iter) {
        double tp = now();
        // images of the new basis vectors and the new rows/columns of G
        for (int j = m - nnew; j < m; ++j) {
            PYCI_TRY(S.apply(S.V + (size_t)j * S.ld, S.W + (size_t)j * S.ld));
            PYCI_TRY(S.dots(S.V, m, S.W + (size_t)j * S.ld, col.data()));
            float ms = 0;
            cudaEventElapsedTime(&ms, S.e0, S.e1);
            spmv_ms += ms;
            for (int i = 0; i < m; ++i)
                G[(size_t)i * mmax + j] = G[(size_t)j * mmax + i] = col[i];
        }
        tr[0] += now() - tp;
        tp = now();
        // Rayleigh-Ritz
        Gw.assign((size_t)m * m, 0.0);
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j)
                Gw[(size_t)i * m + j] = 0.5 * (G[(size_t)i * mmax + j] + G[(size_t)j * mmax + i]);
        jacobi_eigh(Gw, m, theta, Z);
        tr[1] += now() - tp;
        tp = now();
        // residuals and corrections
        const int nr = std::min(nroot, m);
        double *dsums = S.dsmall + S.small_cap - 3 * nroot; // per root: |r|^2, x.(theta-D)^-1 r, x.(theta-D)^-1 x
        PYCI_CUDA(cudaMemsetAsync(dsums, 0, sizeof(double) * 3 * nroot, S.st));
        std::vector<double> sj(m);
        for (int r = 0; r < nr; ++r) {
            for (int i = 0; i < m; ++i)
                sj[i] = Z[(size_t)i * m + r];
            int err;
            const double *dsj = S.stage(sj.data(), m, &err);
            PYCI_TRY(err);
            ritz_residual_kernel<<<S.grid, RB, sizeof(double) * (size_t)(m + 2), S.st>>>(
                S.V, S.W, S.ld, m, dsj, theta[r], op->diag, S.X + (size_t)r * S.ld, S.AX + (size_t)r * S.ld,
                S.T + (size_t)r * S.ld, S.nloc, dsums + 3 * r);
            ctx->launches++;
        }
        if (R > 1)
            PYCI_TRY(comm_allreduce_sum_f64(ctx, dsums, 3 * nroot));
        PYCI_CUDA(cudaMemcpyAsync(S.hsmall, dsums, sizeof(double) * 3 * nroot, cudaMemcpyDeviceToHost, S.st));
        PYCI_CUDA(cudaStreamSynchronize(S.st));
        const std::vector<double> sums(S.hsmall, S.hsmall + 3 * nroot); // (hsmall is reused by the dot products below)
        bool all = (nr == nroot);
        double worst = 0.0;
        for (int r = 0; r < nr; ++r) {
            rn[r] = std::sqrt(std::max(sums[3 * r], 0.0));
            conv[r] = rn[r] <= tol * std::max(eps23, std::fabs(theta[r]));
            all = all && conv[r];
            worst = std::max(worst, rn[r]);
        }
        S.stats.residual = worst;
        S.stats.iterations = iter + 1;
        tr[2] += now() - tp;
        tp = now();
        if (all || m == nrow) {
            // m == nrow: the subspace is the whole space, the Ritz pairs are exact
            done = (nr == nroot);
            break;
        }
        // restart when the corrections would not fit
        int nunconv = 0;
        for (int r = 0; r < nr; ++r)
            nunconv += !conv[r];
        if (m + nunconv > mmax) {
            // (Measured and not kept: GD+k, i.e. also keeping the part of the previous step's Ritz vector that the kept
            // ones do not span -- 497 against 503 products at the default subspace of the 5 M-determinant config-5-style
            // operator, 492 / 527 at 8 vectors, 500 / 494 at 20: with half the basis kept the momentum direction is
            // already in it.)
            // Thick restart: keep the lowest `keep` Ritz vectors (the wanted ones first), not only the wanted ones --
            // collapsing to the current approximation throws the Krylov information of the subspace away and costs
            // two to three times the matvecs on dense spectra (config 5).  V <- V Z, W <- W Z in place, G = diag(theta).
            int keep = std::max(nr, std::min(m - 1, keep_target));
            if (m > ROT_MAXM)
                keep = nr;
            if (keep > nr) {
                std::memcpy(S.hZ, Z.data(), sizeof(double) * (size_t)m * m);
                PYCI_CUDA(cudaMemcpyAsync(S.dZ, S.hZ, sizeof(double) * (size_t)m * m, cudaMemcpyHostToDevice, S.st));
                rotate_kernel<<<S.grid, RB, sizeof(double) * ((size_t)m * m + 2), S.st>>>(S.V, S.ld, m, keep, S.dZ, S.nloc);
                rotate_kernel<<<S.grid, RB, sizeof(double) * ((size_t)m * m + 2), S.st>>>(S.W, S.ld, m, keep, S.dZ, S.nloc);
                ctx->launches += 2;
                PYCI_CUDA(cudaStreamSynchronize(S.st)); // hZ is reused by the next restart
            } else {
                PYCI_CUDA(cudaMemcpyAsync(S.V, S.X, vec * nr, cudaMemcpyDeviceToDevice, S.st));
                PYCI_CUDA(cudaMemcpyAsync(S.W, S.AX, vec * nr, cudaMemcpyDeviceToDevice, S.st));
            }
            std::fill(G.begin(), G.end(), 0.0);
            for (int r = 0; r < keep; ++r)
                G[(size_t)r * mmax + r] = theta[r];
            m = keep;
            S.stats.restarts++;
        }
        // expand
        nnew = 0;
        for (int r = 0; r < nr && m < mmax; ++r) {
            if (conv[r])
                continue;
            double *t = S.T + (size_t)r * S.ld;
            if (olsen && std::fabs(sums[3 * r + 2]) > 1.0e-300) {
                const double eps = sums[3 * r + 1] / sums[3 * r + 2];
                olsen_kernel<<<S.grid, RB, 0, S.st>>>(t, S.X + (size_t)r * S.ld, op->diag, theta[r], eps, S.nloc);
                ctx->launches++;
            }
            double rel = 0.0;
            PYCI_TRY(S.orthonormalize(t, m, tmp, &rel, S.V + (size_t)m * S.ld)); // lands in the next basis slot
            if (rel < 1.0e-10)
                continue; // correction already in the subspace
            ++m;
            ++nnew;
        }
        if (nnew == 0) {
            // cannot expand: add a fresh pseudo-random direction to escape stagnation
            double *t = S.T;
            guess_kernel<<<S.grid, RB, 0, S.st>>>(t, op->row0, S.nloc, nrow, -1, 1.0, 0x1234u + (u32)iter);
            ctx->launches++;
            double rel = 0.0;
            PYCI_TRY(S.orthonormalize(t, m, tmp, &rel));
            if (rel < 1.0e-10 || m >= mmax)
                break;
            PYCI_CUDA(cudaMemcpyAsync(S.V + (size_t)m * S.ld, t, vec, cudaMemcpyDeviceToDevice, S.st));
            ++m;
            nnew = 1;
        }
        tr[3] += now() - tp;
    }
    if (trace)
        fprintf(stderr, "[pyci_b200 solver] rank %d: %ld iterations, apply+dots %.3f s, jacobi %.3f s, ritz/residual %.3f s, "
                        "orthogonalise/expand %.3f s\n", ctx->rank, (long)S.stats.iterations, tr[0], tr[1], tr[2], tr[3]);
    PYCI_CUDA(cudaEventRecord(t_end, S.st));
    PYCI_CUDA(cudaStreamSynchronize(S.st));
    float total_ms = 0;
    cudaEventElapsedTime(&total_ms, t_begin, t_end);
    S.stats.seconds = total_ms * 1e-3;
    S.stats.spmv_seconds = spmv_ms * 1e-3;
    if (stats_out)
        *stats_out = S.stats;
    if (!done)
        PYCI_FAIL(PYCI_ERR_RUNTIME, "did not converge");

    // results: evals (+ecore, sparseop.cpp:140-141), evecs[n][nrow] row-major (:142-144).  Order of the n lowest pairs:
    // the reference calls Spectra's compute(SmallestAlge, maxit, tol) with the default `sorting` argument
    // (sparseop.cpp:134), LargestAlge - the n smallest eigenvalues come back LARGEST FIRST.  The reference's
    // pyci/test/test_odometer.py:38-40,47-60 depends on it (costs = -evals of the len-1 lowest roots).
    for (int r = 0; r < nroot; ++r)
        evals[nroot - 1 - r] = theta[r] + op->ecore;
    for (int r = 0; r < nroot; ++r) {
        const double *src = S.X + (size_t)r * S.ld;
        if (R > 1) {
            PYCI_TRY(op_allgather_rows(ctx, op, src, S.xfull));
            src = S.xfull;
        }
        PYCI_CUDA(cudaMemcpyAsync(evecs + (size_t)(nroot - 1 - r) * nrow, src, sizeof(double) * nrow, cudaMemcpyDeviceToHost,
                                  S.st));
    }
    PYCI_CUDA(cudaStreamSynchronize(S.st));
    return PYCI_OK;
}
