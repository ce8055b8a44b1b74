// Host <-> device transfer of column-major matrices for the host-pointer entry points.
//
// A Julia `Matrix{Float64}` (the snapshots of POD.jl:79-119, the basis of deim.jl:9) lives in PAGEABLE memory.
// Measured on the B200 boxes of this pool (tools/h2d_probe.cu, 16 GiB): cudaMemcpyAsync from pageable memory 8.2 GB/s
// (the driver stages through one internal buffer), cudaHostRegister of the caller's pages ahead of the copy
// 4-5 GB/s (page pinning is the bottleneck), host threads copying windows into a ring of pinned staging buffers
// from which the DMA engine reads: 36 GB/s with 8 threads; page-locked source (mor_b200_host_alloc): 54.7 GB/s.
// So: a page-locked pointer is handed to the DMA engine directly, anything else goes through the staging ring.
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "internal.h"

namespace mor {

namespace {

constexpr size_t STAGE_BYTES = size_t(64) << 20;

// A few persistent host threads that split one 2-D copy between them (the calling thread takes a share too).
struct CopyPool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<void(int)> job;
    int njobs = 0, next = 0, pending = 0;
    uint64_t gen = 0;
    bool quit = false;

    explicit CopyPool(int nthreads) {
        for (int t = 0; t < nthreads; ++t) th.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
        }
        cv_go.notify_all();
        for (auto& t : th) t.join();
    }
    void run() {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv_go.wait(lk, [&] { return quit || gen != seen; });
            if (quit) return;
            seen = gen;
            drain(lk);
        }
    }
    void drain(std::unique_lock<std::mutex>& lk) {
        while (next < njobs) {
            const int j = next++;
            lk.unlock();
            job(j);
            lk.lock();
            if (--pending == 0) cv_done.notify_all();
        }
    }
    void parallel_for(int n, const std::function<void(int)>& fn) {
        std::unique_lock<std::mutex> lk(mu);
        job = fn;
        njobs = n;
        next = 0;
        pending = n;
        ++gen;
        cv_go.notify_all();
        drain(lk);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};

int ingest_threads() {
    const char* e = getenv("MOR_B200_INGEST_THREADS");
    if (e && atoi(e) >= 1) return atoi(e);
    const int hw = (int)std::thread::hardware_concurrency();
    const int per = hw / std::max(1, multi_on() ? multi_n() : 1);
    return std::max(1, std::min(8, per));
}

int32_t stage_ready(Ctx& c) {
    if (c.stage_buf[0]) return MOR_OK;
    for (int b = 0; b < Ctx::NSTAGE; ++b) {
        MOR_CUDA(cudaHostAlloc(&c.stage_buf[b], STAGE_BYTES, cudaHostAllocPortable));
        MOR_CUDA(cudaEventCreateWithFlags(&c.stage_ev[b], cudaEventDisableTiming));
        c.stage_busy[b] = false;
    }
    c.copy_pool = new CopyPool(ingest_threads() - 1);
    return MOR_OK;
}

// dense (rows x cols, ld = rows) <-> strided column-major, split over the pool by columns / row pieces
void copy2d_host(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols, CopyPool* pool) {
    const int T = (int)pool->th.size() + 1;
    if (cols >= T) {
        pool->parallel_for(T, [=](int t) {
            const int64_t c0 = cols * t / T, c1 = cols * (t + 1) / T;
            for (int64_t j = c0; j < c1; ++j) memcpy(dst + j * ldd, src + j * lds, sizeof(double) * (size_t)rows);
        });
    } else {
        pool->parallel_for(T, [=](int t) {
            const int64_t r0 = rows * t / T, r1 = rows * (t + 1) / T;
            for (int64_t j = 0; j < cols; ++j) memcpy(dst + j * ldd + r0, src + j * lds + r0, sizeof(double) * (size_t)(r1 - r0));
        });
    }
}

}  // namespace

bool host_is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

void stage_release() {
    Ctx& c = ctx();
    for (int b = 0; b < Ctx::NSTAGE; ++b) {
        if (c.stage_buf[b]) {
            cudaEventDestroy(c.stage_ev[b]);
            cudaFreeHost(c.stage_buf[b]);
            c.stage_buf[b] = nullptr;
        }
    }
    delete static_cast<CopyPool*>(c.copy_pool);
    c.copy_pool = nullptr;
    cudaGetLastError();
}

// window shape for a rows x cols column-major region: whole columns while they fit, else pieces of one column
static void window_shape(int64_t rows, int64_t cols, int64_t* rw, int64_t* cw) {
    const int64_t cap = (int64_t)(STAGE_BYTES / sizeof(double));
    if (rows <= cap) {
        *rw = rows;
        *cw = std::max<int64_t>(1, std::min(cols, cap / std::max<int64_t>(rows, 1)));
    } else {
        *rw = cap;
        *cw = 1;
    }
}

int32_t h2d_2d(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return MOR_OK;
    Ctx& c = ctx();
    static const bool force_stage = getenv("MOR_B200_INGEST_FORCE_STAGING") != nullptr;   // test hook
    if ((host_is_pinned(src) && !force_stage) || (size_t)rows * cols * 8 < (size_t(1) << 20)) {
        MOR_CUDA(cudaMemcpy2DAsync(dst, (size_t)ldd * 8, src, (size_t)lds * 8, (size_t)rows * 8, (size_t)cols,
                                   cudaMemcpyHostToDevice, st));
        c.last_ingest_staged = false;
        return MOR_OK;
    }
    MOR_TRY(stage_ready(c));
    CopyPool* pool = static_cast<CopyPool*>(c.copy_pool);
    int64_t rw, cw;
    window_shape(rows, cols, &rw, &cw);
    for (int64_t c0 = 0; c0 < cols; c0 += cw)
        for (int64_t r0 = 0; r0 < rows; r0 += rw) {
            const int64_t nr = std::min(rw, rows - r0), nc = std::min(cw, cols - c0);
            const int b = c.stage_next;
            c.stage_next = (c.stage_next + 1) % Ctx::NSTAGE;
            if (c.stage_busy[b]) MOR_CUDA(cudaEventSynchronize(c.stage_ev[b]));
            double* sb = static_cast<double*>(c.stage_buf[b]);
            copy2d_host(sb, nr, src + r0 + c0 * lds, lds, nr, nc, pool);
            MOR_CUDA(cudaMemcpy2DAsync(dst + r0 + c0 * ldd, (size_t)ldd * 8, sb, (size_t)nr * 8, (size_t)nr * 8, (size_t)nc,
                                       cudaMemcpyHostToDevice, st));
            MOR_CUDA(cudaEventRecord(c.stage_ev[b], st));
            c.stage_busy[b] = true;
        }
    c.last_ingest_staged = true;
    return MOR_OK;
}

// sync: return after the data is in `dst`.  The staged path always does (it copies out of the ring on the host); a
// page-locked destination with sync == false only enqueues the DMA on `st`.
int32_t d2h_2d(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols, cudaStream_t st, bool sync) {
    if (rows <= 0 || cols <= 0) return MOR_OK;
    Ctx& c = ctx();
    static const bool force_stage = getenv("MOR_B200_INGEST_FORCE_STAGING") != nullptr;
    if ((host_is_pinned(dst) && !force_stage) || (size_t)rows * cols * 8 < (size_t(1) << 20)) {
        MOR_CUDA(cudaMemcpy2DAsync(dst, (size_t)ldd * 8, src, (size_t)lds * 8, (size_t)rows * 8, (size_t)cols,
                                   cudaMemcpyDeviceToHost, st));
        if (sync) MOR_CUDA(cudaStreamSynchronize(st));
        return MOR_OK;
    }
    MOR_TRY(stage_ready(c));
    CopyPool* pool = static_cast<CopyPool*>(c.copy_pool);
    for (int b = 0; b < Ctx::NSTAGE; ++b)
        if (c.stage_busy[b]) {
            MOR_CUDA(cudaEventSynchronize(c.stage_ev[b]));
            c.stage_busy[b] = false;
        }
    int64_t rw, cw;
    window_shape(rows, cols, &rw, &cw);
    struct Win { int64_t r0, c0, nr, nc; int b; };
    std::vector<Win> wins;
    for (int64_t c0 = 0; c0 < cols; c0 += cw)
        for (int64_t r0 = 0; r0 < rows; r0 += rw)
            wins.push_back({r0, c0, std::min(rw, rows - r0), std::min(cw, cols - c0), 0});
    auto issue = [&](size_t w) -> int32_t {
        Win& x = wins[w];
        x.b = (int)(w % Ctx::NSTAGE);
        MOR_CUDA(cudaMemcpy2DAsync(c.stage_buf[x.b], (size_t)x.nr * 8, src + x.r0 + x.c0 * lds, (size_t)lds * 8, (size_t)x.nr * 8,
                                   (size_t)x.nc, cudaMemcpyDeviceToHost, st));
        MOR_CUDA(cudaEventRecord(c.stage_ev[x.b], st));
        return MOR_OK;
    };
    const size_t ahead = Ctx::NSTAGE - 1;
    for (size_t w = 0; w < wins.size() && w < ahead; ++w) MOR_TRY(issue(w));
    for (size_t w = 0; w < wins.size(); ++w) {
        if (w + ahead < wins.size()) MOR_TRY(issue(w + ahead));
        const Win& x = wins[w];
        MOR_CUDA(cudaEventSynchronize(c.stage_ev[x.b]));
        copy2d_host(dst + x.r0 + x.c0 * ldd, ldd, static_cast<const double*>(c.stage_buf[x.b]), x.nr, x.nr, x.nc, pool);
    }
    return MOR_OK;
}

}  // namespace mor
