// Single-process multi-GPU mode: mor_b200_init_multi(ngpus) / MOR_B200_NGPUS.
//
// One Julia (or Python) process calls the host-pointer entry points exactly as on one GPU --
// reduce!(pod, SVD()) of /root/reference/src/DataReduction/POD.jl:156-166, deim_interpolation_indices of
// /root/reference/src/deim.jl:9-28 -- and the library uses the whole box: one worker thread per GPU (each bound to
// its device, with its own context, streams and NCCL rank: the in-process equivalent of ncclCommInitAll), the
// caller's column-major host matrix is split into contiguous ROW ranges (rank order = row order, so "lowest global
// index" keeps reproducing Julia's first-index tie-break), every worker runs the ordinary single-rank entry point on
// its rows (X + row0, same leading dimension), small replicated results come back from rank 0 and the tall result
// (rbasis) is written by every worker straight into its rows of the caller's output matrix.
// The DEIM candidate exchange uses the peers' mailboxes directly (cudaDeviceEnablePeerAccess, same address space).
//
// Entry points that are per-device by nature (the device-resident API, the single-rank sparse Galerkin product) and
// inputs that do not shard (a wide snapshot matrix, tiny problems) run on GPU 0 as a one-rank job ("solo").
#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

#include "internal.h"

namespace mor {

namespace {

struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int32_t()> job;
    bool has_job = false, done = false, quit = false;
    int32_t status = MOR_OK;
    std::string err;
};

struct Multi {
    int n = 0;
    std::vector<std::unique_ptr<Worker>> w;
    bool env_checked = false;
};
// Leaked on purpose: a process that exits without mor_b200_shutdown (a Julia session) must not run ~thread on the
// still-joinable worker threads during static destruction (std::terminate).
Multi& g_multi = *new Multi;
thread_local bool tl_worker = false;

void worker_main(int r, Worker* w) {
    bind_ctx(r);
    tl_worker = true;
    for (;;) {
        std::function<int32_t()> job;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return w->has_job || w->quit; });
            if (w->quit && !w->has_job) return;
            job = std::move(w->job);
            w->has_job = false;
        }
        const int32_t st = job();
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->status = st;
            w->err = st == MOR_OK ? std::string() : last_error_string();
            w->done = true;
        }
        w->cv.notify_all();
    }
}

void post(Worker* w, std::function<int32_t()> job) {
    {
        std::lock_guard<std::mutex> lk(w->mu);
        w->job = std::move(job);
        w->has_job = true;
        w->done = false;
    }
    w->cv.notify_all();
}

int32_t wait(Worker* w, std::string* err) {
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return w->done; });
    if (w->status != MOR_OK && err && err->empty()) *err = w->err;
    return w->status;
}

struct SoloGuard {   // GPU 0 acts as a one-rank job for the duration
    Ctx& c;
    int rank, nranks;
    explicit SoloGuard(Ctx& c_) : c(c_), rank(c_.rank), nranks(c_.nranks) {
        c.rank = 0;
        c.nranks = 1;
    }
    ~SoloGuard() {
        c.rank = rank;
        c.nranks = nranks;
    }
};

}  // namespace

bool is_worker() { return tl_worker; }
void multi_skip_env() { g_multi.env_checked = true; }
int multi_n() { return g_multi.n; }

int32_t multi_all(const std::function<int32_t(int)>& fn) {
    for (int r = 0; r < g_multi.n; ++r) post(g_multi.w[(size_t)r].get(), [&fn, r]() -> int32_t { return fn(r); });
    int32_t first = MOR_OK;
    std::string err;
    for (int r = 0; r < g_multi.n; ++r) {
        const int32_t st = wait(g_multi.w[(size_t)r].get(), &err);
        if (st != MOR_OK && first == MOR_OK) first = st;
    }
    if (first != MOR_OK) set_error(err);
    return first;
}

int32_t multi_solo(const std::function<int32_t()>& fn) {
    Worker* w = g_multi.w[0].get();
    post(w, [&fn]() -> int32_t {
        SoloGuard g(ctx());
        return fn();
    });
    std::string err;
    const int32_t st = wait(w, &err);
    if (st != MOR_OK) set_error(err);
    return st;
}

void multi_shard(int64_t m, int r, int64_t* row0, int64_t* rows) {
    const int64_t align = 4, n = g_multi.n;
    const int64_t blocks = (m + align - 1) / align, base = blocks / n, extra = blocks % n;
    const int64_t b0 = r * base + (r < extra ? r : extra), nb = base + (r < extra ? 1 : 0);
    const int64_t a = std::min(b0 * align, m), b = std::min((b0 + nb) * align, m);
    *row0 = a;
    *rows = b - a;
}

static int32_t start_multi(int ngpus) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("mor_b200_init_multi: no CUDA device available (this library has no CPU fallback)");
        return MOR_ERR_CUDA;
    }
    MOR_ARG(ngpus >= 1 && ngpus <= MBOX_MAXRANKS && ngpus <= count, "mor_b200_init_multi: ngpus must be in 1..min(8, visible devices)");
    g_multi.w.clear();
    for (int r = 0; r < ngpus; ++r) {
        g_multi.w.emplace_back(new Worker);
        Worker* w = g_multi.w.back().get();
        w->th = std::thread(worker_main, r, w);
    }
    g_multi.n = ngpus;
    unsigned char id[128] = {0};
    int32_t st = multi_all([](int r) -> int32_t { return init_device(ctx(), r); });
    if (st == MOR_OK && ngpus > 1) st = mor_b200_comm_unique_id(id);
    // ncclCommInitRank from one thread per GPU with a shared id: the in-process form of ncclCommInitAll
    if (st == MOR_OK) st = multi_all([&](int r) -> int32_t { return mor_b200_comm_init(r, g_multi.n, id); });
    // mailboxes: allocate everywhere first, then map the peers' (same address space: enable peer access, take the pointer)
    if (st == MOR_OK)
        st = multi_all([](int) -> int32_t {
            ctx().peers_tried = true;   // no CUDA IPC inside one process: the peers are mapped below
            return deim_mailbox();
        });
    if (st == MOR_OK && ngpus > 1)
        st = multi_all([](int r) -> int32_t {
            Ctx& c = ctx();
            bool ok = true;
            for (int p = 0; p < g_multi.n; ++p) {
                if (p == r) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, c.device, ctx_of(p).device) != cudaSuccess || !can) ok = false;
                if (ok) {
                    const cudaError_t e = cudaDeviceEnablePeerAccess(ctx_of(p).device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
                    cudaGetLastError();
                }
            }
            for (int p = 0; p < g_multi.n && ok; ++p) {
                c.peer_base[p] = ctx_of(p).box_base;
                c.peer_ll[p] = ctx_of(p).box_ll;
            }
            c.peers_tried = true;
            c.peers_ok = ok && getenv("MOR_B200_DEIM_EXCHANGE") == nullptr;
            c.peers_ipc = false;
            return MOR_OK;
        });
    if (st == MOR_OK && ngpus > 1) {   // all ranks or none
        bool all = true;
        for (int r = 0; r < ngpus; ++r) all = all && ctx_of(r).peers_ok;
        for (int r = 0; r < ngpus; ++r) ctx_of(r).peers_ok = all;
    }
    if (st != MOR_OK) {
        const std::string keep = last_error_string();
        multi_shutdown();
        set_error(keep);
    }
    return st;
}

int32_t multi_shutdown() {
    if (g_multi.n == 0) return MOR_OK;
    multi_all([](int) -> int32_t {
        mor_b200_comm_destroy();
        teardown_device(ctx());
        return MOR_OK;
    });
    for (auto& w : g_multi.w) {
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->quit = true;
        }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
    }
    g_multi.w.clear();
    g_multi.n = 0;
    return MOR_OK;
}

// MOR_B200_NGPUS=N (N >= 2) in the environment: the first library call of a process that has not been initialised
// otherwise starts the multi-GPU mode, so `reduce!(pod, SVD())` from a plain Julia session uses the whole box.
bool multi_on() {
    if (g_multi.n == 0 && !g_multi.env_checked && !tl_worker) {
        g_multi.env_checked = true;
        const char* e = getenv("MOR_B200_NGPUS");
        if (e && atoi(e) >= 2 && !ctx().ready) {
            int count = 0;
            if (cudaGetDeviceCount(&count) == cudaSuccess && count >= 2) start_multi(std::min(atoi(e), count));
            cudaGetLastError();
        }
    }
    return g_multi.n > 0;
}

}  // namespace mor

using namespace mor;

extern "C" {

int32_t mor_b200_init_multi(int32_t ngpus) {
    MOR_API_LOCK;
    if (is_worker()) {
        set_error("mor_b200_init_multi: called from a worker thread");
        return MOR_ERR_ARG;
    }
    g_multi.env_checked = true;
    if (g_multi.n == ngpus && ngpus > 0) return MOR_OK;
    if (g_multi.n > 0) multi_shutdown();
    if (ctx().ready) mor_b200_shutdown();
    return start_multi(ngpus);
}

int32_t mor_b200_multi_info(int32_t* ngpus) {
    if (ngpus) *ngpus = g_multi.n;
    return MOR_OK;
}

}  // extern "C"

// ---- sharded forms of the host-pointer entry points (called with the API lock held, from a non-worker thread) -------
namespace mor {

// problems this small, or wide ones, run on GPU 0 alone
static bool shards_well(int64_t m, int64_t n) { return m >= n && m >= (int64_t)4096 * g_multi.n && m * n >= (int64_t(1) << 22); }

int32_t multi_pod_factor(const double* X, const double* const* cols, int64_t m, int64_t n, int64_t ldx, int32_t method,
                         int64_t k, int64_t p, const double* Omega, uint64_t seed, mor_pod_t* handle, double* S_out,
                         int64_t* nS) {
    MOR_ARG((X || cols) && handle && S_out && nS && m >= 0 && n >= 1, "mor_b200_pod_factor: bad arguments");
    if (!shards_well(m, n)) {
        return multi_solo([&]() -> int32_t {
            return cols ? mor_b200_pod_factor_cols(cols, m, n, method, k, p, Omega, seed, handle, S_out, nS)
                        : mor_b200_pod_factor(X, m, n, ldx, method, k, p, Omega, seed, handle, S_out, nS);
        });
    }
    const int G = g_multi.n;
    std::unique_ptr<mor_pod_s> H(new mor_pod_s);
    H->parts.assign((size_t)G, nullptr);
    H->part_row0.assign((size_t)G, 0);
    H->part_rows.assign((size_t)G, 0);
    std::vector<std::vector<double>> S((size_t)G, std::vector<double>((size_t)std::min(m, n) + 1));
    std::vector<int64_t> ns((size_t)G, 0);
    std::vector<std::vector<const double*>> colp((size_t)G);
    for (int r = 0; r < G; ++r) {
        multi_shard(m, r, &H->part_row0[(size_t)r], &H->part_rows[(size_t)r]);
        if (cols) {
            colp[(size_t)r].resize((size_t)n);
            for (int64_t j = 0; j < n; ++j) {
                MOR_ARG(cols[j] != nullptr, "mor_b200_pod_factor_cols: null column pointer");
                colp[(size_t)r][(size_t)j] = cols[j] + H->part_row0[(size_t)r];
            }
        }
    }
    mor_pod_s* Hp = H.get();
    int32_t st = multi_all([&](int r) -> int32_t {
        const size_t i = (size_t)r;
        return cols ? mor_b200_pod_factor_cols(colp[i].data(), Hp->part_rows[i], n, method, k, p, Omega, seed, &Hp->parts[i],
                                               S[i].data(), &ns[i])
                    : mor_b200_pod_factor(X + Hp->part_row0[i], Hp->part_rows[i], n, ldx, method, k, p, Omega, seed,
                                          &Hp->parts[i], S[i].data(), &ns[i]);
    });
    if (st != MOR_OK) {
        const std::string keep = last_error_string();
        multi_all([&](int r) -> int32_t { return mor_b200_pod_free(Hp->parts[(size_t)r]); });
        set_error(keep);
        return st;
    }
    for (int64_t i = 0; i < ns[0]; ++i) S_out[i] = S[0][(size_t)i];
    *nS = ns[0];
    H->m = m;
    H->n = n;
    H->m_global = m;
    H->method = method;
    H->r = H->parts[0]->r;
    H->passes = H->parts[0]->passes;
    H->sweeps = H->parts[0]->sweeps;
    *handle = H.release();
    return MOR_OK;
}

int32_t multi_pod_basis(mor_pod_t h, int64_t k, double* U_out, int64_t ldu, double* V_out, int64_t ldv) {
    if (h->parts.empty())   // a factorisation that ran on GPU 0 alone
        return multi_solo([&]() -> int32_t { return mor_b200_pod_basis(h, k, U_out, ldu, V_out, ldv); });
    MOR_ARG(U_out != nullptr && ldu >= std::max<int64_t>(h->m, 1), "mor_b200_pod_basis: leading dimension too small");
    return multi_all([&](int r) -> int32_t {
        const size_t i = (size_t)r;
        return mor_b200_pod_basis(h->parts[i], k, U_out + h->part_row0[i], ldu, r == 0 ? V_out : nullptr, ldv);
    });
}

int32_t multi_pod_free(mor_pod_t h) {
    if (h->parts.empty()) return multi_solo([&]() -> int32_t { return mor_b200_pod_free(h); });
    multi_all([&](int r) -> int32_t { return mor_b200_pod_free(h->parts[(size_t)r]); });
    delete h;
    return MOR_OK;
}

int32_t multi_deim_indices(const double* B, int64_t m, int64_t k, int64_t ldb, int64_t* idx_out) {
    MOR_ARG(B != nullptr && idx_out != nullptr && m >= 0 && k >= 1, "mor_b200_deim_indices: bad arguments");
    if (m < (int64_t)8192 * g_multi.n)
        return multi_solo([&]() -> int32_t { return mor_b200_deim_indices(B, m, k, ldb, idx_out); });
    const int G = g_multi.n;
    std::vector<std::vector<int64_t>> idx((size_t)G, std::vector<int64_t>((size_t)k));
    int32_t st = multi_all([&](int r) -> int32_t {
        int64_t row0, rows;
        multi_shard(m, r, &row0, &rows);
        return mor_b200_deim_indices(B + row0, rows, k, ldb, idx[(size_t)r].data());
    });
    if (st == MOR_OK) memcpy(idx_out, idx[0].data(), sizeof(int64_t) * (size_t)k);
    return st;
}

int32_t multi_deim_projector(const double* U, int64_t m, int64_t kd, int64_t ldu, const double* V, int64_t kp, int64_t ldv,
                             const int64_t* idx, double* P_out, int64_t ldp) {
    MOR_ARG(U && V && idx && P_out && m >= 0 && kd >= 1 && kp >= 1 && ldp >= kp, "mor_b200_deim_projector: bad arguments");
    if (m < (int64_t)8192 * g_multi.n)
        return multi_solo([&]() -> int32_t { return mor_b200_deim_projector(U, m, kd, ldu, V, kp, ldv, idx, P_out, ldp); });
    const int G = g_multi.n;
    std::vector<std::vector<double>> P((size_t)G);
    for (int r = 1; r < G; ++r) P[(size_t)r].resize((size_t)kp * kd);
    return multi_all([&](int r) -> int32_t {
        int64_t row0, rows;
        multi_shard(m, r, &row0, &rows);
        return mor_b200_deim_projector(U + row0, rows, kd, ldu, V + row0, kp, ldv, idx, r == 0 ? P_out : P[(size_t)r].data(),
                                       r == 0 ? ldp : kp);
    });
}

// V' A V with the COLUMNS of A sharded: GPU r takes columns [c0, c1) of the CSC arrays and the rows [lo, hi) of V that
// those columns touch (for a stencil / MOL operator: its own rows plus a halo of the stencil width) straight from the
// caller's host matrix -- the halo needs no inter-GPU exchange because V lives on the host -- forms
// W_r = A[:, c0:c1]' V and the partial product W_r' V[c0:c1, :]; the k x k partials are all-reduced.
int32_t multi_galerkin_csc(int64_t m, const int64_t* colptr, const int64_t* rowval, const double* nzval, const double* V,
                           int64_t k, int64_t ldv, double* Ahat, int64_t lda) {
    MOR_ARG(colptr && V && Ahat && m >= 1 && k >= 1 && ldv >= m && lda >= k, "mor_b200_galerkin_csc: bad arguments");
    MOR_ARG(colptr[0] == 1 && colptr[m] >= 1, "mor_b200_galerkin_csc: colptr must be 1-based (Julia SparseMatrixCSC)");
    const int64_t nnz = colptr[m] - 1;
    MOR_ARG(nnz == 0 || (rowval && nzval), "mor_b200_galerkin_csc: null rowval / nzval");
    if (m < (int64_t)8192 * g_multi.n)
        return multi_solo([&]() -> int32_t { return mor_b200_galerkin_csc(m, colptr, rowval, nzval, V, k, ldv, Ahat, lda); });
    return multi_all([&](int r) -> int32_t {
        int64_t c0, nc;
        multi_shard(m, r, &c0, &nc);
        // rows of V this slice touches: its own rows (for the final contraction) and whatever its columns reference
        int64_t lo = c0, hi = c0 + nc;
        for (int64_t p = colptr[c0] - 1; p < colptr[c0 + nc] - 1; ++p) {
            const int64_t row = rowval[p] - 1;
            MOR_ARG(row >= 0 && row < m, "galerkin projection: row index outside 1..m in the CSC structure");
            lo = std::min(lo, row);
            hi = std::max(hi, row + 1);
        }
        Ctx& cx = ctx();
        DevBuf dA;
        MOR_TRY(dA.alloc(sizeof(double) * (size_t)k * k));
        MOR_TRY(galerkin_csc_slice(nc, colptr + c0, rowval, nzval, lo, hi - lo, c0 - lo, V, k, ldv, dA.as<double>()));
        MOR_TRY(comm_allreduce_sum(dA.as<double>(), (size_t)k * k));
        if (r == 0)
            MOR_CUDA(cudaMemcpy2DAsync(Ahat, (size_t)lda * 8, dA.p, (size_t)k * 8, (size_t)k * 8, (size_t)k, cudaMemcpyDeviceToHost,
                                       cx.stream));
        MOR_CUDA(cudaStreamSynchronize(cx.stream));
        return MOR_OK;
    });
}

int32_t multi_project(const double* V, int64_t m, int64_t k, int64_t ldv, const double* Xv, int64_t nvec, int64_t ldx,
                      double* out, int64_t ldo) {
    MOR_ARG(V && Xv && out && m >= 0 && k >= 1 && nvec >= 1 && ldo >= k, "mor_b200_project: bad arguments");
    if (m < (int64_t)8192 * g_multi.n)
        return multi_solo([&]() -> int32_t { return mor_b200_project(V, m, k, ldv, Xv, nvec, ldx, out, ldo); });
    const int G = g_multi.n;
    std::vector<std::vector<double>> tmp((size_t)G);
    for (int r = 1; r < G; ++r) tmp[(size_t)r].resize((size_t)k * nvec);
    return multi_all([&](int r) -> int32_t {
        int64_t row0, rows;
        multi_shard(m, r, &row0, &rows);
        return mor_b200_project(V + row0, rows, k, ldv, Xv + row0, nvec, ldx, r == 0 ? out : tmp[(size_t)r].data(),
                                r == 0 ? ldo : k);
    });
}

}  // namespace mor
