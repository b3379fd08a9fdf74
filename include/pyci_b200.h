/*
 * pyci_b200 -- C ABI of the B200 (sm_100a) CI-Hamiltonian hot path.
 *
 * This is the drop-in boundary beneath the reference's pybind11 module `pyci._pyci`
 * (/root/reference/pyci/src/binding.cpp:29): plain pointers and sizes, opaque handles, int status,
 * no exceptions and no torch/pybind types.  The host module `pyci_b200._pyci` (C++/pybind11, same
 * class and method names as the reference) binds exactly these entry points; INTEGRATION.md shows
 * the equivalent stub a PyCI maintainer would add.  Citations are relative to /root/reference.
 *
 * All `long` are int64.  Host pointers unless a name ends in `_dev`.  Every function returns
 * PYCI_OK or a negative status; pyci_last_error() gives the message of the calling thread's last
 * failure.  There is NO CPU fallback: without a usable CUDA device every compute entry point fails
 * with PYCI_ERR_CUDA.
 */
#ifndef PYCI_B200_H
#define PYCI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PYCI_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define PYCI_API __attribute__((visibility("default")))
#else
#define PYCI_API
#endif

/* status codes; the pybind11 layer maps them to the reference's exception types
 * (wfn.cpp:52-57, sparseop.cpp:117-119,136) */
enum {
    PYCI_OK = 0,
    PYCI_ERR_VALUE = -1,   /* std::invalid_argument / domain_error -> ValueError   */
    PYCI_ERR_TYPE = -2,    /* pybind11::type_error                 -> TypeError    */
    PYCI_ERR_RUNTIME = -3, /* std::runtime_error ("did not converge") -> RuntimeError */
    PYCI_ERR_CUDA = -4,    /* CUDA / NCCL failure, no device       -> RuntimeError */
    PYCI_ERR_MEMORY = -5,  /* device or host allocation failed     -> MemoryError  */
    PYCI_ERR_UNSUPPORTED = -6 /* valid for the reference, not on device (nbasis > 64) -> RuntimeError */
};

/* wave-function kinds: DOCIWfn / FullCIWfn / GenCIWfn (pyci.h:496-619) */
enum { PYCI_DOCI = 0, PYCI_FULLCI = 1, PYCI_GENCI = 2 };

typedef struct pyci_ctx pyci_ctx; /* one device + stream (+ NCCL communicator when row-sharded) */
typedef struct pyci_ham pyci_ham; /* SQuantOp integrals resident in HBM          (pyci.h:285-302) */
typedef struct pyci_wfn pyci_wfn; /* determinant array + GPU hash index          (pyci.h:306-330) */
typedef struct pyci_op pyci_op;   /* SparseOp: CSR row shard resident in HBM     (pyci.h:621-693) */

PYCI_API const char *pyci_last_error(void);
PYCI_API int pyci_abi_version(void);
/* number of visible CUDA devices (0 when there is none); never fails */
PYCI_API int pyci_device_count(void);

/* ---- context ------------------------------------------------------------------------------- */

/* device: CUDA ordinal.  stream: a cudaStream_t to launch on (e.g. torch's current stream), or
 * NULL to let the context create its own non-blocking stream. */
PYCI_API int pyci_ctx_create(int device, void *stream, pyci_ctx **out);
PYCI_API void pyci_ctx_destroy(pyci_ctx *ctx);
PYCI_API int pyci_ctx_synchronize(pyci_ctx *ctx);
/* Device allocations come from the device's stream-ordered memory pool, whose release threshold the library raises so
 * that a rebuilt operator reuses the blocks of the destroyed one (10-100 GB) instead of paying the driver again.
 * Memory the pool holds is invisible to other allocators of the process (torch's caching allocator, cudaMalloc):
 * this hands everything that is not in use back to the driver. */
PYCI_API int pyci_ctx_release_memory(pyci_ctx *ctx);
/* Page-locked host memory for the buffers a binding hands to the export calls (pyci_op_export_csr, ...): every entry
 * point accepts any host pointer, but a device-to-host copy into pageable memory goes through the driver's bounce
 * buffers at about a quarter of the PCIe rate.  The host module keeps a small pool of these for the row pointer it
 * returns (the reference returns a numpy array over its own vector: pyci/src/binding.cpp:520-528). */
PYCI_API int pyci_host_alloc(void **ptr, size_t bytes);
PYCI_API void pyci_host_free(void *ptr);
/* Row-sharding across `nranks` processes (one per GPU of one box).  unique_id: the 128 bytes of an
 * ncclUniqueId made by pyci_nccl_unique_id() on rank 0 and handed to the other ranks by the caller
 * (torch.distributed broadcast, MPI, a file ...).  Collective: every rank must call it. */
PYCI_API int pyci_nccl_unique_id(void *unique_id_128);
PYCI_API int pyci_ctx_init_comm(pyci_ctx *ctx, int rank, int nranks, const void *unique_id_128);
PYCI_API int pyci_ctx_rank(const pyci_ctx *ctx);
PYCI_API int pyci_ctx_nranks(const pyci_ctx *ctx);
/* Number of this library's kernels launched through ctx since creation / since the last reset. */
PYCI_API long pyci_ctx_launch_count(const pyci_ctx *ctx);
PYCI_API void pyci_ctx_reset_launch_count(pyci_ctx *ctx);

/* ---- Hamiltonian: SQuantOp (squantop.cpp:163-183) --------------------------------------------- */

/* one_mo[n*n], two_mo[n^4] (physicist order <ik|jl> at i*n^3+k*n^2+j*n+l), h[n], v[n*n], w[n*n]. */
PYCI_API int pyci_ham_upload(pyci_ctx *ctx, long nbasis, double ecore, const double *one_mo,
                    const double *two_mo, const double *h, const double *v, const double *w,
                    pyci_ham **out);
PYCI_API void pyci_ham_destroy(pyci_ham *ham);

/* ---- wave function: determinant storage + index_det (onespinwfn.cpp:123-126, twospinwfn.cpp:129-132) */

/* dets: [ndet][nword] (DOCI, GenCI) or [ndet][2][nword] (FullCI) uint64 bit-strings, nword =
 * ceil(nbasis/64) (common.cpp:280-282).  Builds the open-addressing GPU hash keyed by the bit-string.
 * Device kernels handle nword == 1; larger nbasis returns PYCI_ERR_UNSUPPORTED.  Duplicate
 * determinants return PYCI_ERR_VALUE. */
PYCI_API int pyci_wfn_upload(pyci_ctx *ctx, int kind, long nbasis, long nocc_up, long nocc_dn, long ndet,
                    const uint64_t *dets, pyci_wfn **out);
/* Wfn::add_all_dets (onespinwfn.cpp:173-217, twospinwfn.cpp:181-245): the complete space of (nbasis, nocc_up,
 * nocc_dn), unranked in HBM in the reference's order -- colex for DOCI / GenCI, colex(alpha) * C(nbasis, nocc_dn) +
 * colex(beta) for FullCI -- so that a wave function filled by add_all_dets needs no determinant upload. */
PYCI_API int pyci_wfn_create_all_dets(pyci_ctx *ctx, int kind, long nbasis, long nocc_up, long nocc_dn, pyci_wfn **out);
PYCI_API void pyci_wfn_destroy(pyci_wfn *wfn);
/* Rebuild the hash index from the determinants already resident in HBM (what SparseOp::update's
 * callers get from Wfn::add_det, onespinwfn.cpp:141-149: the index is part of the construction path).
 * pyci_wfn_index_seconds: device seconds of the last index build. */
PYCI_API int pyci_wfn_reindex(pyci_wfn *wfn);
PYCI_API double pyci_wfn_index_seconds(const pyci_wfn *wfn);
/* index_det for a batch of determinants (same layout as dets); out[i] = row index or -1 */
PYCI_API int pyci_wfn_index_dets(pyci_wfn *wfn, long n, const uint64_t *dets, long *out);

/* number of determinants now in the device wave function, and a copy of determinants [start, start+n) */
PYCI_API long pyci_wfn_ndet(const pyci_wfn *wfn);
PYCI_API int pyci_wfn_download_dets(const pyci_wfn *wfn, long start, long n, uint64_t *out);

/* ---- selected CI: add_hci (hci.cpp:238-279) and compute_enpt2 (enpt2.cpp:344-400) ------------------- */

/* One heat-bath iteration: every excitation j of every determinant i with |H_ji| > eps / |coeffs[i]| that is not
 * yet in wfn is appended to wfn (once) and the index is rebuilt; *n_new = number appended (the return value of
 * the reference's add_hci).  coeffs[ndet].  The reference appends in the iteration order of its hash map
 * (twospinwfn.cpp:267-270: unspecified); here the new determinants come in first-encounter order of the
 * reference's serial loop nest.  Read them back with pyci_wfn_download_dets(wfn, old_ndet, *n_new, ...).
 * The reference's nthread argument has no meaning on the device.  Collective when row-sharded. */
PYCI_API int pyci_wfn_add_hci(pyci_ctx *ctx, const pyci_ham *ham, pyci_wfn *wfn, const double *coeffs, double eps,
                     long *n_new);
/* Epstein-Nesbet second-order energy: *out = energy + sum_j (sum_i H_ji c_i)^2 / (energy - ecore - H_jj) over the
 * external determinants j reached with |H_ji| > eps / |c_i| (enpt2.cpp:344-374).  FullCI and GenCI wave
 * functions; for DOCI the reference converts to FullCI first (enpt2.cpp:376-380) and so must the caller.
 * nterms (may be NULL) receives the number of external determinants.  Collective when row-sharded. */
PYCI_API int pyci_compute_enpt2(pyci_ctx *ctx, const pyci_ham *ham, pyci_wfn *wfn, const double *coeffs, double energy,
                       double eps, double *out, long *nterms);
/* device seconds of the last add_hci / compute_enpt2 walk over wfn */
PYCI_API double pyci_wfn_ext_seconds(const pyci_wfn *wfn);

/* ---- sparse operator: SparseOp (sparseop.cpp) --------------------------------------------------- */

/* SparseOp::SparseOp + update (sparseop.cpp:49-71,186-201): rows [0,nrow) x columns [0,ncol) of H
 * in the determinant basis of wfn; nrow/ncol < 0 mean ndet.  symmetric != 0 gives the operator the
 * reference's lower-triangular export; on the device the full rows are kept for a gather SpMV.
 * With a communicator, rank r builds the contiguous row block r of ceil(nrow/nranks) rows.  The rows of a selected
 * space are then re-partitioned (collectively) so that every rank stores the same number of entries -- contiguous
 * ranges balanced by nnz; pyci_op_row_begin / pyci_op_row_count report what the rank holds (PYCI_B200_NO_REBALANCE=1
 * keeps the uniform blocks; complete spaces are uniform by construction). */
PYCI_API int pyci_op_build(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, long nrow, long ncol,
                  int symmetric, pyci_op **out);
/* The row block that rank `rank` of `nranks` owns, built WITHOUT a communicator: construction has no collective
 * (the reference's rows are independent, sparseop.cpp:196-199), so one process can build -- and check -- any shard
 * of a row-sharded operator.  Export, pyci_op_get_element and pyci_op_matvec_dev (full x in, the shard's rows out)
 * work on it; pyci_op_matvec / pyci_op_solve / pyci_op_update need the context's own layout (pyci_op_build). */
PYCI_API int pyci_op_build_shard(pyci_ctx *ctx, const pyci_ham *ham, const pyci_wfn *wfn, long nrow, long ncol,
                        int symmetric, int rank, int nranks, pyci_op **out);
PYCI_API void pyci_op_destroy(pyci_op *op);
/* SparseOp::update (sparseop.cpp:175-201): grow an operator built for the first op.nrow determinants of wfn to all
 * wfn.ndet of them (the caller appended determinants, e.g. with pyci_wfn_add_hci; the first op.nrow determinants
 * must be the ones the operator was built from, as for the reference).  Only the new determinants are enumerated.
 * Symmetric (square): their rows are built, and their entries with an old column are transposed into the ends of
 * the old rows (the device keeps full rows); the exported CSR equals that of a fresh build.  Non-symmetric: the new
 * rows [op.nrow, ndet) x [0, ndet) are appended and the rows the operator has keep the columns they were built with,
 * exactly what the reference's update leaves (NOT the fresh-build matrix).
 * PYCI_ERR_UNSUPPORTED for symmetric operators with nrow != ncol and when row-sharded: rebuild instead. */
PYCI_API int pyci_op_update(pyci_op *op, const pyci_ham *ham, const pyci_wfn *wfn);

PYCI_API long pyci_op_nrow(const pyci_op *op);
PYCI_API long pyci_op_ncol(const pyci_op *op);
/* first row and number of rows held by this rank */
PYCI_API long pyci_op_row_begin(const pyci_op *op);
PYCI_API long pyci_op_row_count(const pyci_op *op);
/* SparseOp::size of this rank's rows in the REFERENCE's storage (lower triangle + diagonal when symmetric) */
PYCI_API long pyci_op_size(const pyci_op *op);
/* non-zeros this rank streams per SpMV (full rows) */
PYCI_API long pyci_op_stored_nnz(const pyci_op *op);
PYCI_API double pyci_op_ecore(const pyci_op *op);
/* device seconds of the last build on this rank: [0] hash index, [1] count+scan, [2] fill+sort, [3] total */
PYCI_API int pyci_op_build_times(const pyci_op *op, double *seconds4);
/* device seconds of the fill kernel alone in the last build (CUDA events around its launch on the context's stream:
 * the duration the roofline fraction of bench.py is computed from) */
PYCI_API double pyci_op_fill_seconds(const pyci_op *op);
/* name of the CUDA kernel that filled this operator (the dominant kernel of a construction; profiling aid --
 * the reference has one code path, SparseOp::add_row, sparseop.cpp:220-502) */
PYCI_API const char *pyci_op_fill_kernel(const pyci_op *op);
/* what found the stored entries of the rows: "analytic" (complete space: the count is a formula), "count_kernel"
 * (one index probe per candidate excitation, the reference's algorithm) or "join_rows_kernel" (selected spaces:
 * determinants bucketed by segment pairs and compared by XOR / popcount, join.cuh) */
PYCI_API const char *pyci_op_count_kernel(const pyci_op *op);

/* py_indptr / py_indices / py_data (sparseop.cpp:504-514) for this rank's rows, in the reference's
 * layout: indptr[row_count+1] starting at 0, indices int64, data fp64, each row sorted by column
 * (sparseop.cpp:214-218). */
PYCI_API int pyci_op_export_csr(pyci_op *op, long *indptr, long *indices, double *data);

/* The same export for a list of rows (global indices, all held by this rank; any order, repeats allowed): what
 * py_indptr / py_indices / py_data would hold for those rows.  indptr[nrows+1] starts at 0.  indices / data may
 * both be NULL to size the buffers first (indptr[nrows] = entries needed); cap = their capacity in entries.
 * For operators whose whole export does not fit the host (configs 3-5) -- sampled-row parity checks. */
PYCI_API int pyci_op_export_rows(pyci_op *op, long nrows, const long *rows, long cap, long *indptr, long *indices,
                        double *data);

/* SparseOp::perform_op (sparseop.cpp:96-112): y[nrow] = A x[ncol]; ecore is not applied.
 * Host buffers; with a communicator every rank passes the full x and receives the full y. */
PYCI_API int pyci_op_matvec(pyci_op *op, const double *x, double *y);
/* Same on device buffers: x_dev[ncol] full, y_dev[row_count] this rank's rows; asynchronous on the
 * context's stream; no collective. */
PYCI_API int pyci_op_matvec_dev(pyci_op *op, const double *x_dev, double *y_dev);
/* Measurement helper: `warmup` untimed + `reps` timed launches of the SpMV kernel on a device-resident
 * pseudo-random x, each timed with CUDA events on the context's stream; ms[reps] receives the per-launch
 * device times.  flush_bytes > 0 overwrites a scratch buffer of that size before every launch (evicts L2). */
PYCI_API int pyci_op_time_spmv(pyci_op *op, int warmup, int reps, long flush_bytes, double *ms);
/* Launch shape of the SpMV kernel: threads cooperating on one row (32, 64, 128, 256, or 0 = chosen from
 * the mean row length) and resident CTAs of 256 threads per SM (default 4).  A tuning knob only: results
 * are identical for every shape up to the summation order inside a row. */
PYCI_API int pyci_op_set_spmv_shape(pyci_op *op, int threads_per_row, int ctas_per_sm);
/* Threads per CTA of the row kernel (256, 512 or 1024): block_threads / threads_per_row consecutive rows are
 * streamed in lockstep by one CTA and share their gathers of x in L1; depth (2..4) = trips of (value, column)
 * loads every thread keeps in flight.  Tuning knobs only. */
PYCI_API int pyci_op_set_spmv_block(pyci_op *op, int block_threads, int depth);
/* SparseOp::get_element (sparseop.cpp:89-94); i must be a row of this rank.  The row asked for last is kept on the
 * host, so a walk along a row costs one device round trip (not thread-safe on one handle, like the reference's GIL-held
 * call) */
PYCI_API int pyci_op_get_element(pyci_op *op, long i, long j, double *out);

typedef struct pyci_solve_stats {
    long matvecs;        /* SpMV launches */
    long iterations;     /* Davidson iterations */
    long restarts;
    double residual;     /* max residual norm of the returned pairs */
    double seconds;      /* device seconds inside the solver */
    double spmv_seconds; /* of which SpMV (+ all-gather) */
} pyci_solve_stats;

/* SparseOp::solve_ci (sparseop.cpp:114-146): n lowest eigenpairs, evals[n] (ecore added),
 * evecs[n][nrow] row-major, in the reference's order for n > 1: largest of the n first (Spectra's default
 * `sorting` of the selected pairs, sparseop.cpp:134; pyci/test/test_odometer.py depends on it).  c0: nrow doubles or NULL; ncv: subspace size or -1 for
 * min(nrow, max(2n+1, 20)); maxiter: -1 for 10*n*nrow; tol: residual tolerance relative to
 * max(eps^(2/3), |theta|) as in Spectra.  Errors follow sparseop.cpp:116-124,136.
 * Collective when row-sharded (NCCL all-gather of the trial vector every iteration). */
PYCI_API int pyci_op_solve(pyci_op *op, long n, const double *c0, long ncv, long maxiter, double tol,
                  double *evals, double *evecs, pyci_solve_stats *stats);

/* ---- reduced density matrices: compute_rdms (rdm.cpp:20-65, 269-530, 532-632) ------------------- */

/* DOCI: rdm1 = d0[n*n], rdm2 = d2[n*n].  FullCI: rdm1[2*n*n] (aa,bb), rdm2[3*n^4] (aaaa,bbbb,abab).
 * GenCI: rdm1[n*n], rdm2[n^4] (fully antisymmetric; the reference routine is defective).
 * Collective when row-sharded (all-reduce of the tensors). */
PYCI_API int pyci_compute_rdms(pyci_ctx *ctx, const pyci_wfn *wfn, const double *coeffs, double *rdm1,
                      double *rdm2);

/* compute_transition_rdms (rdm.cpp:634-1009): <Psi1| ... |Psi2> with the shapes of pyci_compute_rdms; rows run over
 * wfn1, excited determinants are looked up in wfn2, every connected ordered pair contributes in one direction
 * (T(wfn, wfn, c, c) = compute_rdms(wfn, c)).  Both wave functions must share kind, nbasis and occupations.
 * GenCI: intended semantics (the reference routine is defective).  Collective when row-sharded. */
PYCI_API int pyci_compute_transition_rdms(pyci_ctx *ctx, const pyci_wfn *wfn1, const pyci_wfn *wfn2,
                                 const double *coeffs1, const double *coeffs2, double *rdm1, double *rdm2);
/* compute_overlap (overlap.cpp:17-58): sum over the determinants common to both wave functions of c1_i c2_j */
PYCI_API int pyci_compute_overlap(pyci_ctx *ctx, const pyci_wfn *wfn1, const pyci_wfn *wfn2, const double *coeffs1,
                         const double *coeffs2, double *out);

#ifdef __cplusplus
}
#endif
#endif /* PYCI_B200_H */
