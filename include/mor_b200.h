/* mor_b200.h -- C ABI of libmor_b200.so: the B200-native numeric hot path behind
 * SciML/ModelOrderReduction.jl's Julia API (POD / reduce! / deim_interpolation_indices).
 *
 * The reference has no FFI of its own; the seams this ABI replaces are (paths relative
 * to the reference tree):
 *   src/DataReduction/POD.jl:8-13   _svd   -> LinearAlgebra.svd        (reduce!(pod, SVD()),  POD.jl:156-166)
 *   src/DataReduction/POD.jl:15-20  _tsvd  -> TSVD.tsvd                (reduce!(pod, TSVD()), POD.jl:195-202)
 *   src/DataReduction/POD.jl:22-27  _rsvd  -> RandomizedLinAlg.rsvd    (reduce!(pod, RSVD()), POD.jl:231-238)
 *   src/DataReduction/POD.jl:4-6    matricize (hcat of a Vector{Vector}) -> column-pointer ingest
 *   src/deim.jl:9-28                deim_interpolation_indices(basis)
 * INTEGRATION.md shows the Julia `ccall` shim that binds every entry point below.
 *
 * Conventions
 *   - plain C, no C++/torch types; every call returns an int32 status (0 = OK).
 *   - matrices are column-major FP64 ("each column is a state snapshot", POD.jl:38-39);
 *     ld* is the leading dimension in elements.  Indices are Int64, 1-based on output.
 *   - by default one process drives one GPU ("rank").  With mor_b200_comm_init() the ranks of a
 *     job hold contiguous ROW shards of every matrix, rank order = row order; small
 *     (n x n) results are replicated, tall results (U) stay row-sharded.  Alternatively ONE process
 *     drives all GPUs of the box and the library does the sharding (mor_b200_init_multi).
 *   - host matrices may be pageable (a Julia Matrix) or page-locked (mor_b200_host_alloc): pageable ones
 *     travel through a ring of pinned staging buffers filled by a few host threads.
 *   - a device matrix handed to mor_b200_pod_factor_dev must stay alive and unmodified until the handle
 *     is freed (the handle borrows it unless a second pass had to copy it).
 *   - host buffers are read/written only during the call; nothing is retained.
 *   - calls are synchronous (they return after the work is done) unless named *_async.
 *   - no CPU fallback exists: without a usable CUDA device every compute entry point
 *     returns MOR_ERR_CUDA.
 */
#ifndef MOR_B200_H
#define MOR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOR_B200_VERSION 100 /* 0.1.0 */

/* status codes (shim: 1,2 -> ArgumentError; 6 -> SingularException; else ErrorException) */
#define MOR_OK 0
#define MOR_ERR_ARG 1         /* null pointer, k outside 1..min(m,n), ld < m, unsupported size */
#define MOR_ERR_NONFINITE 2   /* Inf/NaN in a POD input */
#define MOR_ERR_CUDA 3        /* CUDA / NCCL failure, or no device */
#define MOR_ERR_NOMEM 4       /* out of device memory */
#define MOR_ERR_NOCONV 5      /* Jacobi sweeps / CholeskyQR passes exhausted */
#define MOR_ERR_SINGULAR 6    /* DEIM: exactly singular P'U (Julia: SingularException) */

/* POD backends -- the dispatch tags of src/Types.jl:61-124 */
#define MOR_SVD 0
#define MOR_TSVD 1
#define MOR_RSVD 2

/* synthetic snapshot generators (bench / parity inputs) */
#define MOR_SYNTH_GRADED 0  /* N(0,1) * 10^(-param * j/(n-1)) per column j            */
#define MOR_SYNTH_GAUSS 1   /* N(0,1) * param                                          */
#define MOR_SYNTH_LOWRANK 2 /* graded, but the first `iparam` columns have scale 1     */
#define MOR_SYNTH_KAT 3     /* known-answer matrix H_u [S V'; 0]: singular values exactly 10^(-param j/(n-1)),
                               left vectors the columns of the Householder reflector H_u.  COLLECTIVE over
                               the ranks (one all-reduce); every rank passes its own global_row0            */
#define MOR_SYNTH_BURGERS 4 /* u component of a 2-D viscous Burgers solution (Cole-Hopf closed form: three states,
                               travelling fronts merging in a triple point) on a grid `iparam` nodes wide, row =
                               ix + iparam * iy, column j = time j/(n-1), viscosity `param` (BASELINE config 5) */

typedef struct mor_pod_s* mor_pod_t; /* factorisation handle (owns device memory)  */
typedef struct mor_mat_s* mor_mat_t; /* device-resident column-major FP64 matrix   */

/* ---- lifecycle ---------------------------------------------------------------------- */
int32_t mor_b200_version(void);
/* Bind this process to CUDA device `device` (>= 0) and create the library stream. */
int32_t mor_b200_init(int32_t device);
int32_t mor_b200_shutdown(void);
/* Thread-local message of the last failing call on this thread ("" if none). */
const char* mor_b200_last_error(void);
/* Run on the caller's CUDA stream (cudaStream_t as void*); NULL restores the own stream. */
int32_t mor_b200_set_stream(void* cuda_stream);
int32_t mor_b200_sync(void);
/* Page-locked host staging memory (cudaHostAlloc): snapshots handed to the host-pointer entry
 * points from such a buffer move at full PCIe speed.  Optional -- any host pointer works. */
/* Give cached device memory back to the driver: the stream-ordered work pool and the parked copy of
 * the last large (> 8 GB) host-ingested matrix, kept so that the next call skips a ~0.4 s
 * cudaMalloc/cudaFree pair.  Done automatically when an allocation fails and at shutdown. */
int32_t mor_b200_trim(void);
int32_t mor_b200_host_alloc(uint64_t bytes, void** out);
int32_t mor_b200_host_free(void* p);

/* ---- multi-GPU, form 1: ONE process drives the whole box ------------------------------------
 * mor_b200_init_multi(ngpus) (or MOR_B200_NGPUS=ngpus in the environment of a process that never calls
 * mor_b200_init) starts one worker thread, context, stream set and NCCL rank per GPU.  The host-pointer entry points
 * below (pod_factor[_cols], pod_basis, deim_indices, deim_projector, project) then take the WHOLE matrix: the library
 * shards its rows over the GPUs, replicated results (spectrum, indices, projector) come back once and the tall result
 * is written into the caller's U_out by all GPUs -- reduce!(pod, SVD()) from a single Julia session uses 8 GPUs.
 * Wide or tiny inputs, mor_b200_galerkin_csc and the device-resident API (mor_b200_mat_*, *_dev) run on GPU 0. */
int32_t mor_b200_init_multi(int32_t ngpus);
int32_t mor_b200_multi_info(int32_t* ngpus); /* 0 when the mode is off */

/* ---- multi-GPU, form 2: one process per GPU, NCCL over NVLink ------------------------------- */
/* Rank 0 creates a 128-byte NCCL unique id; the host runtime (Julia Distributed / MPI,
 * torch.distributed, ...) broadcasts it; every rank then calls comm_init. */
int32_t mor_b200_comm_unique_id(void* id128);
int32_t mor_b200_comm_init(int32_t rank, int32_t nranks, const void* id128);
int32_t mor_b200_comm_info(int32_t* rank, int32_t* nranks);
int32_t mor_b200_comm_destroy(void);

/* ---- POD: two-phase because nmodes may depend on the spectrum ---------------------------
 * replaces POD.jl:157/196/232 (factor) and POD.jl:163/199/235 (basis).
 *   X       m x n host matrix (this rank's row shard), ld = ldx >= m.  Not modified.
 *   method  MOR_SVD: S_out gets all min(m,n) singular values (descending) -- POD.jl:164.
 *           MOR_TSVD / MOR_RSVD: S_out gets exactly k values -- POD.jl:200,236.
 *   k       modes requested (TSVD/RSVD; for SVD an upper bound hint, may be 0).
 *   p       RSVD oversampling (Types.jl:119-124); ignored otherwise.
 *   Omega   RSVD test matrix n x (k+p) column-major (ld = n), or NULL to draw it from
 *           `seed` with the library's Philox generator.
 *   nS      out: number of values written to S_out (capacity must be min(M,n), M = global rows).
 */
int32_t mor_b200_pod_factor(const double* X, int64_t m, int64_t n, int64_t ldx, int32_t method,
                            int64_t k, int64_t p, const double* Omega, uint64_t seed,
                            mor_pod_t* handle, double* S_out, int64_t* nS);
/* Same, but the snapshots arrive as n separate length-m columns (a Julia
 * Vector{Vector{Float64}}): replaces matricize (POD.jl:4-6) without the hcat copy. */
int32_t mor_b200_pod_factor_cols(const double* const* cols, int64_t m, int64_t n, int32_t method,
                                 int64_t k, int64_t p, const double* Omega, uint64_t seed,
                                 mor_pod_t* handle, double* S_out, int64_t* nS);
/* Leading k left singular vectors of this rank's rows -> U_out (m x k, ld = ldu); optional
 * right vectors -> V_out (n x k, ld = ldv) or NULL.  Signs: the largest-|entry| of each
 * right vector is positive (deterministic and shard-count invariant). */
int32_t mor_b200_pod_basis(mor_pod_t handle, int64_t k, double* U_out, int64_t ldu, double* V_out,
                           int64_t ldv);
int32_t mor_b200_pod_free(mor_pod_t handle);
/* passes = CholeskyQR passes over X that the factorisation took (1, 2, 3..);
 * sweeps = Jacobi sweeps of the small core SVD. */
int32_t mor_b200_pod_info(mor_pod_t handle, int32_t* passes, int32_t* sweeps);

/* ---- DEIM: replaces the body of deim_interpolation_indices, deim.jl:9-28 ----------------
 * B is this rank's row shard (m x k, ld = ldb) of the basis; idx_out receives k global
 * 1-based row indices (identical on every rank).  First-maximal-index tie-break, NaN
 * counts as maximal (Julia argmax). */
int32_t mor_b200_deim_indices(const double* B, int64_t m, int64_t k, int64_t ldb, int64_t* idx_out);
/* How firmly the last DEIM call on this rank decided each greedy step: values[l] = r_best, the winning residual of
 * deim.jl:23 at step l+1, margins[l] = (r_best - r_second) / r_best (0 = exact tie).  Writes min(count, cap) entries;
 * *min_margin / *min_step (1-based) are taken over steps >= 2 (step 1 is the argmax of |u_1| itself: no arithmetic,
 * identical in every formulation).  The default blocked path rounds differently from deim.jl:21-23; when a step >= 2
 * is decided by less than 32 k eps (MOR_B200_DEIM_TIE_TOL overrides) the library repeats the call with the streaming
 * path, the reference's own formulation (MOR_B200_DEIM_TIE_FALLBACK=0 disables; MOR_T_DEIM_FALLBACK tells).
 * Any pointer may be NULL. */
int32_t mor_b200_deim_last_margins(double* margins, double* values, int64_t cap, int64_t* count, double* min_margin,
                                   int64_t* min_step);

/* DEIM projector, deim.jl:150: P = ((U[idx, :])' \ (U' * V))'  (pod_dim x deim_dim, ld = ldp),
 * U (m x kd) the DEIM basis, V (m x kp) the POD basis (this rank's row shards), idx the kd global
 * 1-based interpolation indices.  P is replicated on every rank. */
int32_t mor_b200_deim_projector(const double* U, int64_t m, int64_t kd, int64_t ldu, const double* V, int64_t kp,
                                int64_t ldv, const int64_t* idx, double* P_out, int64_t ldp);

/* Galerkin projection of the linear part, deim.jl:154-155,261.
 * A_hat (k x k, ld = lda) = V' * A * V with A an m x m Julia SparseMatrixCSC given by its own arrays
 * (colptr m+1 entries, rowval / nzval nnz entries, 1-based Int64 indices).  One process per GPU: single rank.
 * Single-process multi-GPU mode: the columns of A are sharded over the GPUs (each takes the rows of V its columns
 * touch straight from the caller's host matrix), the k x k partial products are all-reduced. */
int32_t mor_b200_galerkin_csc(int64_t m, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                              const double* V, int64_t k, int64_t ldv, double* Ahat, int64_t lda);
/* out (k x nvec, ld = ldo) = V' * Xv: g_hat = V' * g and y_hat(0) = V' * snapshot[:, 1].  V (m x k) and
 * Xv (m x nvec) are this rank's row shards; the result is replicated. */
int32_t mor_b200_project(const double* V, int64_t m, int64_t k, int64_t ldv, const double* Xv, int64_t nvec,
                         int64_t ldx, double* out, int64_t ldo);

/* ---- device-resident API (inputs already in HBM; used by bench and GPU-side callers) --- */
int32_t mor_b200_mat_alloc(int64_t m, int64_t n, mor_mat_t* out);
/* Borrow caller-owned device memory (e.g. a torch tensor); never freed by the library. */
int32_t mor_b200_mat_wrap(void* device_ptr, int64_t m, int64_t n, int64_t ld, mor_mat_t* out);
int32_t mor_b200_mat_free(mor_mat_t a);
int32_t mor_b200_mat_info(mor_mat_t a, int64_t* m, int64_t* n, int64_t* ld, void** device_ptr);
int32_t mor_b200_mat_upload(mor_mat_t a, const double* host, int64_t ldh);
int32_t mor_b200_mat_download(mor_mat_t a, double* host, int64_t ldh);
/* Fill `a` with rows [global_row0, global_row0 + m) of the synthetic matrix `kind`;
 * the value of an entry depends only on (seed, global row, column): shard invariant. */
int32_t mor_b200_synth(mor_mat_t a, int32_t kind, uint64_t seed, int64_t global_row0,
                       double param, int64_t iparam);
/* Known-answer check for MOR_SYNTH_KAT: max over all ranks, rows and the first k columns of
 * |U[i][j] -+ (H_u e_j)[i]| for this rank's row shard U of the computed left vectors (collective). */
int32_t mor_b200_kat_check(mor_mat_t U, int64_t k, int64_t n, uint64_t seed, int64_t global_row0, double param,
                           double* max_err);
/* flags for pod_factor_dev */
#define MOR_POD_MAY_OVERWRITE 1 /* X may be overwritten by X*inv(R) when >1 pass is needed */
#define MOR_POD_FORCE_TWO_PASS 2 /* always run CholeskyQR2 (never stop after one Gram pass) */
int32_t mor_b200_pod_factor_dev(mor_mat_t X, int32_t method, int64_t k, int64_t p, mor_mat_t Omega,
                                uint64_t seed, int32_t flags, mor_pod_t* handle, double* S_out,
                                int64_t* nS);
/* U (m x k) is written into the device matrix U_out (allocated by the caller, m x >=k). */
int32_t mor_b200_pod_basis_dev(mor_pod_t handle, int64_t k, mor_mat_t U_out);
/* Leading k right singular vectors (n x k, replicated on every rank) into the device matrix V_out. */
int32_t mor_b200_pod_right_dev(mor_pod_t handle, int64_t k, mor_mat_t V_out);
/* Orthonormalise the columns of A in place (CholeskyQR passes until ||A'A - I|| is at
 * rounding level); used to build the DEIM benchmark basis. */
int32_t mor_b200_orthonormalize_dev(mor_mat_t A, int32_t* passes);
int32_t mor_b200_deim_indices_dev(mor_mat_t B, int64_t* idx_out);
int32_t mor_b200_deim_projector_dev(mor_mat_t U, mor_mat_t V, const int64_t* idx, mor_mat_t P_out);

/* ---- instrumentation ----------------------------------------------------------------------
 * CUDA-event timings (ms) of the last POD/DEIM call on this rank, by slot: */
#define MOR_T_TOTAL 0     /* whole call, device side                               */
#define MOR_T_H2D 1       /* host -> device staging                                */
#define MOR_T_GRAM 2      /* all Gram (SYRK) launches incl. split-K reduction      */
#define MOR_T_APPLY 3     /* X <- X*inv(R) launches                                */
#define MOR_T_CORE 4      /* n x n work: Cholesky, inverse, Jacobi SVD             */
#define MOR_T_BASIS 5     /* U_k = X*W                                             */
#define MOR_T_COMM 6      /* NCCL collectives                                      */
#define MOR_T_DEIM_STEP 7 /* the fused residual+argmax(+tail) kernels: total - panel - update */
#define MOR_T_DEIM_TAIL 8 /* part of DEIM_STEP spent in the per-step tails (reduction, exchange, small LU updates) */
#define MOR_T_D2H 9
#define MOR_T_SKETCH 10   /* RSVD: Y = X*Omega and B = Q'X                        */
#define MOR_T_DEIM_PANEL 11 /* DEIM blocked path: panel residual kernels (one per 16 columns) */
#define MOR_T_DEIM_BLOCKED 12 /* not a time: 1 if the last DEIM call took the blocked path, else 0 */
#define MOR_T_DEIM_UPDATE 13 /* DEIM blocked path: per-panel coefficient and inverse-update kernels */
#define MOR_T_DEIM_FALLBACK 14 /* not a time: 1 if a near-tie made the last DEIM call repeat on the streaming path */
#define MOR_T_DEIM_EXCHANGE 15 /* not a time: per-step candidate exchange of the last DEIM call: 0 single rank,
                                  1 peer mailboxes over NVLink (stores + flags inside the step kernel), 2 ncclAllGather */
#define MOR_T_SPMM 16      /* Galerkin projection: the sparse A'V kernel                */
#define MOR_T_PROJ 17      /* Galerkin projection / projector: the dense W'V contraction */
#define MOR_T_NSLOTS 24
int32_t mor_b200_last_timings(double* ms, int32_t n);
/* number of kernel launches issued by the library since the last reset (this rank) */
int64_t mor_b200_launch_count(int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* MOR_B200_H */
