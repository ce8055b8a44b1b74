// Internal declarations shared by the translation units of libmor_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mor_b200.h"

namespace mor {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const std::string& msg);
int32_t cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define MOR_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return ::mor::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)
#define MOR_TRY(call)                    \
    do {                                 \
        int32_t s__ = (call);            \
        if (s__ != MOR_OK) return s__;   \
    } while (0)
#define MOR_ARG(cond, msg)               \
    do {                                 \
        if (!(cond)) {                   \
            ::mor::set_error(msg);       \
            return MOR_ERR_ARG;          \
        }                                \
    } while (0)

// ---- context: one per process (= per GPU rank) ------------------------------------------
struct Ctx {
    bool ready = false;
    int device = -1;
    int num_sms = 148;
    cudaStream_t stream = nullptr;      // the stream all work is enqueued on
    cudaStream_t own_stream = nullptr;  // created by init
    cudaStream_t copy_stream = nullptr; // host->device ingest, overlapped with the first Gram pass; DEIM side stream
    cudaStream_t aux_stream = nullptr;  // DEIM host ingest: column groups upload while earlier panels are processed
    void* comm = nullptr;               // ncclComm_t
    int rank = 0, nranks = 1;
    double timings[MOR_T_NSLOTS] = {0};
    int64_t launches = 0;
    void* scratch_p = nullptr;          // grow-only workspace (partial tiles, packed W)
    size_t scratch_bytes = 0;
    void* ingest_p = nullptr;           // parked device copy of the last large host-ingested matrix:
    size_t ingest_bytes = 0;            // cudaMalloc/cudaFree of 137 GB cost ~0.4 s per call otherwise
    // DEIM exchange mailbox (deim.cu, nccl.cpp): per greedy step every rank stores its candidate record into every
    // peer's mailbox over NVLink and raises a flag there -- no NCCL call on the 256-step latency chain
    void* box_base = nullptr;           // one cudaMalloc: header (ticket) | plain records (NCCL fallback) | tagged words
    double* box = nullptr;
    uint4* box_ll = nullptr;
    unsigned* box_ticket = nullptr;
    void* peer_base[8] = {nullptr};     // peers' mailboxes mapped into this process (CUDA IPC) or this device (P2P)
    uint4* peer_ll[8] = {nullptr};
    bool peers_tried = false, peers_ok = false, peers_ipc = false;
    unsigned long long deim_seq = 0;    // greedy steps taken since comm_init (the same on every rank): slot / flag value
    std::vector<double> deim_margins;   // relative gap to the runner-up at every greedy step of the last call
    std::vector<double> deim_values;    // the winning residual of every greedy step of the last call
    // pageable-host ingest (ingest.cpp): ring of pinned staging buffers filled by a few host copy threads
    static constexpr int NSTAGE = 4;
    void* stage_buf[NSTAGE] = {nullptr};
    cudaEvent_t stage_ev[NSTAGE] = {nullptr};
    bool stage_busy[NSTAGE] = {false};
    int stage_next = 0;
    void* copy_pool = nullptr;
    bool last_ingest_staged = false;    // the last host -> device transfer went through the staging ring
};
// mailbox geometry (shared by deim.cu and nccl.cpp)
constexpr int MBOX_SLOTS = 4, MBOX_MAXRANKS = 8, MBOX_REC_MAX = 1024 + 16 + 4 + 3;
constexpr size_t MBOX_HEADER_BYTES = 1024;   // the ticket at byte 512
constexpr size_t MBOX_PLAIN_BYTES = (sizeof(double) * (size_t)(MBOX_SLOTS * MBOX_MAXRANKS + 1) * MBOX_REC_MAX + 255) & ~size_t(255);
constexpr size_t MBOX_LL_BYTES = 16 * (size_t)(MBOX_SLOTS * MBOX_MAXRANKS) * MBOX_REC_MAX;
constexpr size_t MBOX_BYTES = MBOX_HEADER_BYTES + MBOX_PLAIN_BYTES + MBOX_LL_BYTES;
int32_t deim_mailbox();   // allocate the own mailbox; with several ranks, map the peers' (all ranks or none)
void deim_mailbox_release();
void stage_release();       // pinned staging ring of the pageable-host ingest (ingest.cpp)
bool host_is_pinned(const void* p);
// column-major 2-D transfers; a page-locked host pointer goes to the DMA engine directly, a pageable one through the
// staging ring.  h2d is asynchronous on `st` (the host thread is busy while it fills staging buffers); d2h returns
// when the data is in place.
int32_t h2d_2d(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols, cudaStream_t st);
int32_t d2h_2d(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols, cudaStream_t st,
               bool sync = true);
void multi_skip_env();
const std::string& last_error_string();
// The context of the calling thread: the process-wide one (one process per GPU), or -- inside a worker thread of the
// single-process multi-GPU mode (multi.cpp) -- the context of that worker's GPU.
Ctx& ctx();
Ctx& ctx_of(int r);          // multi mode: the context of rank r
void bind_ctx(int r);        // worker threads: make ctx() mean rank r's context
int32_t init_device(Ctx& c, int device);
void teardown_device(Ctx& c);
int32_t require_ctx();
// Every public entry point holds this (recursive: entry points call each other) for its duration:
// calls from several host threads are safe and simply serialise.  Worker threads of the multi-GPU mode only run
// while the dispatching thread holds the lock, so for them it is a no-op.
std::recursive_mutex& api_mutex();
bool is_worker();
struct ApiLock {
    bool held;
    ApiLock() : held(!is_worker()) {
        if (held) api_mutex().lock();
    }
    ~ApiLock() {
        if (held) api_mutex().unlock();
    }
    ApiLock(const ApiLock&) = delete;
    ApiLock& operator=(const ApiLock&) = delete;
};
#define MOR_API_LOCK ::mor::ApiLock api_lock__

// ---- single-process multi-GPU mode (multi.cpp): mor_b200_init_multi / MOR_B200_NGPUS ----------------------------
bool multi_on();             // true once init_multi succeeded (also performs the lazy MOR_B200_NGPUS start-up)
int multi_n();
int32_t multi_shutdown();
int32_t multi_solo(const std::function<int32_t()>& fn);          // run on GPU 0 as a single-rank job
int32_t multi_all(const std::function<int32_t(int)>& fn);        // run on every GPU's worker thread, first error wins
void multi_shard(int64_t m, int r, int64_t* row0, int64_t* rows);  // contiguous row ranges, starts on multiples of 4
}  // namespace mor
struct mor_pod_s;
namespace mor {
int32_t multi_pod_factor(const double* X, const double* const* cols, int64_t m, int64_t n, int64_t ldx, int32_t method,
                         int64_t k, int64_t p, const double* Omega, uint64_t seed, mor_pod_s** handle, double* S_out,
                         int64_t* nS);
int32_t multi_pod_basis(mor_pod_s* h, int64_t k, double* U_out, int64_t ldu, double* V_out, int64_t ldv);
int32_t multi_pod_free(mor_pod_s* h);
int32_t multi_deim_indices(const double* B, int64_t m, int64_t k, int64_t ldb, int64_t* idx_out);
int32_t multi_deim_projector(const double* U, int64_t m, int64_t kd, int64_t ldu, const double* V, int64_t kp, int64_t ldv,
                             const int64_t* idx, double* P_out, int64_t ldp);
int32_t multi_galerkin_csc(int64_t m, const int64_t* colptr, const int64_t* rowval, const double* nzval, const double* V,
                           int64_t k, int64_t ldv, double* Ahat, int64_t lda);
int32_t galerkin_csc_slice(int64_t ncols, const int64_t* colptr, const int64_t* rowval, const double* nzval, int64_t row_lo,
                           int64_t nrows, int64_t c_off, const double* V, int64_t k, int64_t ldv, double* Ahat_dev);
int32_t multi_project(const double* V, int64_t m, int64_t k, int64_t ldv, const double* Xv, int64_t nvec, int64_t ldx,
                      double* out, int64_t ldo);
// entry points that only make sense on one device (device-resident API, single-rank helpers): in multi mode GPU 0
#define MOR_ROUTE_SOLO(call)                                                                  \
    do {                                                                                      \
        if (::mor::multi_on() && !::mor::is_worker())                                          \
            return ::mor::multi_solo([&]() -> int32_t { return call; });                      \
    } while (0)
inline void count_launch(int n = 1) { ctx().launches += n; }
// Grow-only device workspace of at least `bytes`; contents are only valid until the next
// scratch() call on this context (all users are ordered on ctx().stream).
int32_t scratch(size_t bytes, void** out);
// Device buffer for a host-ingested snapshot matrix: reuses the parked one when it is large enough.
struct DevBuf;
int32_t ingest_alloc(DevBuf& buf, size_t bytes);

// Debug stopwatch: when MOR_B200_DEBUG is set, lap("label") synchronises the stream and prints the
// wall time since the previous lap to stderr (serialises the pipeline: for diagnosis only).
struct DebugLaps {
    bool on;
    double t0;
    explicit DebugLaps(const char* what);
    void lap(const char* label);
};

// RAII device buffer (freed with cudaFree on destruction)
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    bool pooled = false;   // from the stream-ordered pool (freed with cudaFreeAsync)
    bool cacheable = false; // a large host-ingest copy: parked in the context for the next call on release
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    int32_t alloc(size_t nbytes);
    void release();
    void take(DevBuf& o) {  // move ownership
        release();
        p = o.p; bytes = o.bytes; pooled = o.pooled; cacheable = o.cacheable;
        o.p = nullptr; o.bytes = 0; o.pooled = false; o.cacheable = false;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// CUDA-event stopwatch accumulating into ctx().timings[slot]
struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    int slot;
    bool running = false;
    explicit Timer(int slot_);
    ~Timer();
    void start();
    void stop();  // records; elapsed is collected by collect()
    void collect();  // synchronises on b and adds to the slot
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans;
};

}  // namespace mor

// ---- matrices ---------------------------------------------------------------------------
struct mor_mat_s {
    double* p = nullptr;
    int64_t m = 0, n = 0, ld = 0;
    bool owned = false;
};

// ---- factorisation handle ---------------------------------------------------------------------
struct mor_pod_s {
    int64_t m = 0, n = 0;       // this rank's shard of the input X is m x n
    int64_t m_global = 0;
    bool wide = false;          // T = X' was factored (single rank)
    int method = MOR_SVD;
    // tall side: left vectors of T are T * M(:, 0:k)
    double* T = nullptr;
    int64_t trows = 0, ldt = 0;
    int tn = 0;                 // columns of T = rows of M
    mor::DevBuf T_owned;        // set when the library owns T
    mor::DevBuf M;              // tn x r, ld = tn
    mor::DevBuf Vs;             // vs_rows x r, ld = vs_rows: small-side (right) vectors of T
    int vs_rows = 0, r = 0;     // r = singular triplets available
    std::vector<double> sigma;  // r values, descending
    int passes = 0, sweeps = 0;
    // single-process multi-GPU mode: one per-GPU handle per rank and the row range each one holds
    std::vector<mor_pod_s*> parts;
    std::vector<int64_t> part_row0, part_rows;
};

namespace mor {

// ---- collectives (nccl.cpp; NCCL is dlopen'ed on first use) -------------------------------
int32_t comm_allreduce_sum(double* buf, size_t count);
int32_t comm_allgather(const void* send, void* recv, size_t bytes_per_rank);
int32_t comm_allgather_i64(int64_t mine, std::vector<int64_t>& all);  // host-visible

// ---- synthetic inputs (synth.cu) ----------------------------------------------------------
int32_t synth_fill(double* X, int64_t m, int64_t n, int64_t ld, int32_t kind, uint64_t seed,
                   int64_t row0, double param, int64_t iparam);
int32_t fill_normal(double* X, int64_t m, int64_t n, int64_t ld, uint64_t seed, int64_t row0,
                    double scale);
int32_t kat_check_left(const double* U, int64_t m, int64_t ld, int k, int64_t n, uint64_t seed, int64_t row0,
                       double decades, double* max_err);

// ---- DEIM (deim.cu) -------------------------------------------------------------------------
// col_ready (optional): event g is recorded once columns [g * group_cols, (g + 1) * group_cols) of B are in place
// (pipelined host ingest); the greedy loop waits for group g right before it first touches those columns.
int32_t deim_indices_device(const double* B, int64_t m, int64_t k, int64_t ld, int64_t* idx_host,
                            const cudaEvent_t* col_ready = nullptr, int group_cols = 16);
// R (m x bw, ldr) = U[:, l0 : l0+bw] - U[:, 0 : l0] * C with C packed in 16-row slabs (deim_panel.cu)
int32_t deim_panel(const double* U, int64_t ld, int l0, int bw, const double* Cp, int64_t m, double* R, int64_t ldr);

// ---- tall-skinny GEMMs on DMMA (gemm_tn.cu / gemm_nn.cu) -----------------------------------
// C (p x q, ldc, column-major, both triangles when A == B) = A' * B, A: m x p, B: m x q.
// accumulate != 0: C += A' * B (used to build the Gram matrix chunk by chunk while rows still arrive)
int32_t gemm_tn(const double* A, int64_t lda, int p, const double* B, int64_t ldb, int q, int64_t m,
                double* C, int ldc, int accumulate = 0);
// Y (m x q) = X (m x n) * W (n x q).  Wt is W TRANSPOSED, i.e. q x n column-major ... see .cu
// upper != 0: W is upper triangular (W[i][j] = 0 for i > j) and Y may alias X (in place).
int32_t gemm_nn(const double* X, int64_t ldx, int n, const double* W, int ldw, int q, int64_t m,
                double* Y, int64_t ldy, int upper);

// ---- small dense kernels (small.cu) and Jacobi SVD (jacobi.cu) -------------------------------
struct SmallInfo {
    double delta2;   // ||D^-1 G D^-1 - I||_F^2
    int nonfinite;   // Inf/NaN seen in G
    int nzero;       // all-zero columns (G_jj == 0)
    int chol_fail;   // 1-based index of the first non-positive pivot, 0 = ok
    int rotations;   // Jacobi rotations applied in the last sweep
    double max_cos;  // largest |cosine| between two columns met in the last sweep
    int sync_fail;   // Jacobi: a block-version wait timed out (never expected; reported instead of hanging)
    int pad_;
};
int32_t small_info_reset(SmallInfo* d_info);
int32_t small_info_read(const SmallInfo* d_info, SmallInfo* h);
int32_t small_diag(const double* G, int n, int ld, double* D, SmallInfo* d_info);
int32_t small_scale(const double* G, int n, int ld, const double* D, double shift, double* Gs, int lds,
                    SmallInfo* d_info);
int32_t small_chol_upper(double* A, int n, int ld, SmallInfo* d_info);   // A = R'R, lower part zeroed
int32_t small_trinv_upper(const double* R, int n, int ld, double* Rinv, int ldi);
int32_t small_gemm(bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int lda,
                   const double* B, int ldb, double beta, double* C, int ldc);
// A[i][j] *= s with s = d[i] (rows) or d[j]; inv: s = 1/d; a zero entry of d acts as 1
int32_t small_scale_rc(double* A, int m, int n, int ld, const double* d, bool rows, bool inv);
int32_t small_identity(double* A, int n, int ld);
// A X = B (or A' X = B) by LU with partial pivoting, in place; *d_fail (device int) = 1-based zero pivot
int32_t small_lu_solve(double* A, int n, int lda, bool transpose, double* B, int nrhs, int ldb, int* d_fail);
// out (k x k) = rows idx[i]-1 (global, 1-based) of this rank's shard of U (m x k), zeros for rows of other ranks
int32_t gather_rows(const double* U, int64_t m, int64_t ld, int64_t row0, const long long* d_idx, int k, double* out,
                    int ldo);
int32_t small_zero_rows(double* A, int m, int n, int ld, const double* flag);
int32_t transpose(const double* in, int64_t rows, int64_t cols, int64_t ldin, double* out, int64_t ldout);
int32_t gather_cols(const double* in, int ldin, int rows, int ncols, const int* d_perm, const double* d_scale,
                    double* out, int ldout);
int32_t col_norms(const double* A, int lda, int rows, int n, std::vector<double>& out);   // host result
// One-sided Jacobi on the columns of A (rows x n, lda), in place: on return the columns are
// mutually orthogonal (A_in * V = A_out, V n x n accumulated if non-null).  sigma = column
// norms sorted descending, order[j] = column holding the j-th largest.
// well_scaled: A is a row / column scaling of a well-conditioned matrix (lets the blocked sweep use its faster
// Gram-form round; see jacobi.cu)
int32_t jacobi_svd(double* A, int rows, int n, int lda, double* V, int ldv, std::vector<double>& sigma,
                   std::vector<int>& order, int* sweeps_out, bool well_scaled = false);

// ---- POD driver (pod.cu) ---------------------------------------------------------------------
}  // namespace mor
