// Context, error plumbing, device matrices, lifecycle entry points of libmor_b200.so.
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <mutex>

#include "internal.h"

namespace mor {

static thread_local std::string g_last_error;
static Ctx g_ctxs[MBOX_MAXRANKS];          // [0]: the process-wide context; [r]: GPU r of the multi-GPU mode
static thread_local Ctx* tl_ctx = &g_ctxs[0];
static std::mutex g_mutex;
#define g_ctx (*tl_ctx)

void set_error(const std::string& msg) { g_last_error = msg; }

int32_t cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file,
             line, what);
    set_error(buf);
    // clear the sticky-free error state so later calls report their own failures
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? MOR_ERR_NOMEM : MOR_ERR_CUDA;
}

Ctx& ctx() { return *tl_ctx; }
Ctx& ctx_of(int r) { return g_ctxs[r]; }
void bind_ctx(int r) { tl_ctx = &g_ctxs[r]; }
const std::string& last_error_string() { return g_last_error; }

std::recursive_mutex& api_mutex() {
    static std::recursive_mutex m;
    return m;
}

int32_t require_ctx() {
    if (!g_ctx.ready) {
        // lazy default: device 0
        return mor_b200_init(0);
    }
    MOR_CUDA(cudaSetDevice(g_ctx.device));
    return MOR_OK;
}

// Small work buffers come from the stream-ordered pool (cudaMallocAsync): a plain
// cudaMalloc/cudaFree pair costs ~25 ms each next to a 100+ GB allocation (measured: 300 ms of
// a 980 ms POD step), the pool costs microseconds and never synchronises the device.
static constexpr size_t POOL_LIMIT = size_t(8) << 30;   // larger buffers (snapshot copies) use cudaMalloc
static constexpr uint64_t POOL_KEEP = uint64_t(8) << 30;  // freed work memory the pool may keep cached

// give cached pool memory back to the driver (used before retrying a failed cudaMalloc)
static void trim_pool() {
    cudaMemPool_t pool = nullptr;
    if (g_ctx.ready && cudaDeviceGetDefaultMemPool(&pool, g_ctx.device) == cudaSuccess) {
        cudaStreamSynchronize(g_ctx.stream);
        cudaMemPoolTrimTo(pool, 0);
    }
    if (g_ctx.ingest_p) {
        cudaFree(g_ctx.ingest_p);
        g_ctx.ingest_p = nullptr;
        g_ctx.ingest_bytes = 0;
    }
    cudaGetLastError();
}

int32_t ingest_alloc(DevBuf& buf, size_t bytes) {
    buf.release();
    if (bytes <= POOL_LIMIT) return buf.alloc(bytes);   // small: the pool is already fast
    if (g_ctx.ingest_p && g_ctx.ingest_bytes >= bytes) {
        buf.p = g_ctx.ingest_p;
        buf.bytes = g_ctx.ingest_bytes;
        g_ctx.ingest_p = nullptr;
        g_ctx.ingest_bytes = 0;
    } else {
        if (g_ctx.ingest_p) {
            cudaFree(g_ctx.ingest_p);
            g_ctx.ingest_p = nullptr;
            g_ctx.ingest_bytes = 0;
        }
        MOR_TRY(buf.alloc(bytes));
    }
    buf.cacheable = true;
    return MOR_OK;
}

int32_t DevBuf::alloc(size_t nbytes) {
    release();
    if (nbytes == 0) nbytes = 8;
    pooled = nbytes <= POOL_LIMIT && g_ctx.ready;
    cudaError_t e = pooled ? cudaMallocAsync(&p, nbytes, g_ctx.stream) : cudaMalloc(&p, nbytes);
    if (e == cudaErrorMemoryAllocation) {  // the pool may be sitting on the memory we need
        cudaGetLastError();
        trim_pool();
        e = pooled ? cudaMallocAsync(&p, nbytes, g_ctx.stream) : cudaMalloc(&p, nbytes);
    }
    if (e != cudaSuccess) {
        p = nullptr;
        return cuda_fail(e, pooled ? "cudaMallocAsync" : "cudaMalloc", __FILE__, __LINE__);
    }
    bytes = nbytes;
    return MOR_OK;
}
void DevBuf::release() {
    if (p) {
        if (pooled && g_ctx.ready) {
            cudaFreeAsync(p, g_ctx.stream);
        } else if (cacheable && g_ctx.ready && bytes > g_ctx.ingest_bytes) {
            if (g_ctx.ingest_p) cudaFree(g_ctx.ingest_p);   // park the larger buffer
            cudaStreamSynchronize(g_ctx.stream);
            g_ctx.ingest_p = p;
            g_ctx.ingest_bytes = bytes;
        } else {
            cudaFree(p);
        }
    }
    cacheable = false;
    p = nullptr;
    bytes = 0;
    pooled = false;
}

int32_t scratch(size_t bytes, void** out) {
    Ctx& c = g_ctx;
    if (bytes > c.scratch_bytes) {
        if (c.scratch_p) {
            cudaStreamSynchronize(c.stream);
            cudaFree(c.scratch_p);
            c.scratch_p = nullptr;
            c.scratch_bytes = 0;
        }
        size_t want = bytes + bytes / 4;
        cudaError_t e = cudaMalloc(&c.scratch_p, want);
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            trim_pool();
            e = cudaMalloc(&c.scratch_p, want);
        }
        if (e != cudaSuccess) {
            c.scratch_p = nullptr;
            return cuda_fail(e, "cudaMalloc(scratch)", __FILE__, __LINE__);
        }
        c.scratch_bytes = want;
    }
    *out = c.scratch_p;
    return MOR_OK;
}

static double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
DebugLaps::DebugLaps(const char* what) : on(getenv("MOR_B200_DEBUG") != nullptr), t0(0.0) {
    if (on) {
        cudaStreamSynchronize(g_ctx.stream);
        fprintf(stderr, "[mor_b200] %s\n", what);
        t0 = now_ms();
    }
}
void DebugLaps::lap(const char* label) {
    if (!on) return;
    cudaStreamSynchronize(g_ctx.stream);
    const double t = now_ms();
    fprintf(stderr, "[mor_b200]   %-28s %9.3f ms\n", label, t - t0);
    t0 = t;
}

Timer::Timer(int slot_) : slot(slot_) {}
Timer::~Timer() {
    if (running) {
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
    for (auto& s : spans) {
        cudaEventDestroy(s.first);
        cudaEventDestroy(s.second);
    }
}
void Timer::start() {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, ctx().stream);
    running = true;
}
void Timer::stop() {
    if (!running) return;
    cudaEventRecord(b, ctx().stream);
    spans.emplace_back(a, b);
    running = false;
}
void Timer::collect() {
    for (auto& s : spans) {
        cudaEventSynchronize(s.second);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.first, s.second) == cudaSuccess) ctx().timings[slot] += ms;
        cudaEventDestroy(s.first);
        cudaEventDestroy(s.second);
    }
    spans.clear();
}

int32_t init_device(Ctx& c, int device) {
    if (c.ready && c.device == device) return MOR_OK;
    MOR_ARG(device >= 0, "mor_b200_init: device must be >= 0");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("mor_b200_init: no CUDA device available (this library has no CPU fallback)");
        return MOR_ERR_CUDA;
    }
    MOR_ARG(device < count, "mor_b200_init: device ordinal out of range");
    MOR_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MOR_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("mor_b200_init: libmor_b200 is built for sm_100a (B200) only");
        return MOR_ERR_CUDA;
    }
    if (c.own_stream) cudaStreamDestroy(c.own_stream);
    MOR_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
    c.stream = c.own_stream;
    {   // keep up to POOL_KEEP of freed work buffers cached in the default pool
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = POOL_KEEP;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    c.device = device;
    c.num_sms = prop.multiProcessorCount;
    c.ready = true;
    return MOR_OK;
}

// the calling thread must be the one whose ctx() is `c` (DevBuf / mailbox helpers go through ctx())
void teardown_device(Ctx& c) {
    if (!c.ready) return;
    cudaSetDevice(c.device);
    cudaStreamSynchronize(c.stream);
    if (c.scratch_p) cudaFree(c.scratch_p);
    c.scratch_p = nullptr;
    c.scratch_bytes = 0;
    if (c.ingest_p) cudaFree(c.ingest_p);
    c.ingest_p = nullptr;
    c.ingest_bytes = 0;
    deim_mailbox_release();
    stage_release();
    if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
    c.copy_stream = nullptr;
    if (c.aux_stream) cudaStreamDestroy(c.aux_stream);
    c.aux_stream = nullptr;
    if (c.own_stream) cudaStreamDestroy(c.own_stream);
    c.own_stream = nullptr;
    c.stream = nullptr;
    c.ready = false;
}

}  // namespace mor

using namespace mor;

extern "C" {

int32_t mor_b200_version(void) { return MOR_B200_VERSION; }

const char* mor_b200_last_error(void) { return g_last_error.c_str(); }

int32_t mor_b200_init(int32_t device) {
    MOR_API_LOCK;
    if (multi_on() && !is_worker()) {
        set_error("mor_b200_init: the library is in single-process multi-GPU mode (mor_b200_init_multi); call mor_b200_shutdown first");
        return MOR_ERR_ARG;
    }
    multi_skip_env();            // an explicit single-device init wins over MOR_B200_NGPUS
    std::lock_guard<std::mutex> lock(g_mutex);
    return init_device(g_ctx, device);
}

int32_t mor_b200_shutdown(void) {
    MOR_API_LOCK;
    if (multi_on() && !is_worker()) return multi_shutdown();
    mor_b200_comm_destroy();
    std::lock_guard<std::mutex> lock(g_mutex);
    teardown_device(g_ctx);
    return MOR_OK;
}

int32_t mor_b200_set_stream(void* cuda_stream) {
    MOR_API_LOCK;
    if (multi_on() && !is_worker()) {
        set_error("mor_b200_set_stream: not available in single-process multi-GPU mode (every GPU runs on its own stream)");
        return MOR_ERR_ARG;
    }
    MOR_TRY(require_ctx());
    MOR_CUDA(cudaStreamSynchronize(g_ctx.stream));
    g_ctx.stream = cuda_stream ? (cudaStream_t)cuda_stream : g_ctx.own_stream;
    return MOR_OK;
}

int32_t mor_b200_sync(void) {
    MOR_API_LOCK;
    if (multi_on() && !is_worker()) return multi_all([](int) -> int32_t { return mor_b200_sync(); });
    MOR_TRY(require_ctx());
    MOR_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return MOR_OK;
}

int32_t mor_b200_trim(void) {
    MOR_API_LOCK;
    if (multi_on() && !is_worker()) return multi_all([](int) -> int32_t { return mor_b200_trim(); });
    MOR_TRY(require_ctx());
    trim_pool();
    return MOR_OK;
}

int32_t mor_b200_host_alloc(uint64_t bytes, void** out) {
    MOR_API_LOCK;
    MOR_ARG(out != nullptr, "mor_b200_host_alloc: null output pointer");
    MOR_TRY(require_ctx());
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes > 0 ? (size_t)bytes : 8, cudaHostAllocPortable);   // every GPU of the process may DMA from it
    if (e != cudaSuccess) {
        *out = nullptr;
        return cuda_fail(e, "cudaHostAlloc", __FILE__, __LINE__);
    }
    *out = p;
    return MOR_OK;
}

int32_t mor_b200_host_free(void* p) {
    MOR_API_LOCK;
    if (p) MOR_CUDA(cudaFreeHost(p));
    return MOR_OK;
}

int32_t mor_b200_last_timings(double* ms, int32_t n) {
    MOR_ARG(ms != nullptr && n >= 0, "mor_b200_last_timings: bad arguments");
    for (int i = 0; i < n; ++i) ms[i] = i < MOR_T_NSLOTS ? g_ctx.timings[i] : 0.0;
    return MOR_OK;
}

int64_t mor_b200_launch_count(int32_t reset) {
    MOR_API_LOCK;
    int64_t v = 0;
    const int n = multi_on() && !is_worker() ? multi_n() : 1;
    for (int r = 0; r < n; ++r) {
        Ctx& c = n > 1 ? ctx_of(r) : g_ctx;
        v += c.launches;
        if (reset) c.launches = 0;
    }
    return v;
}

// ---- device matrices ----------------------------------------------------------------------
int32_t mor_b200_mat_alloc(int64_t m, int64_t n, mor_mat_t* out) {
    MOR_API_LOCK;
    MOR_ROUTE_SOLO(mor_b200_mat_alloc(m, n, out));
    MOR_ARG(out != nullptr && m >= 0 && n >= 0, "mor_b200_mat_alloc: bad arguments");
    MOR_TRY(require_ctx());
    int64_t ld = (m + 1) & ~int64_t(1);  // even: 16-byte aligned columns (TMA / v2 loads)
    if (ld == 0) ld = 2;
    size_t bytes = (size_t)ld * (size_t)(n > 0 ? n : 1) * sizeof(double);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        trim_pool();
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(matrix)", __FILE__, __LINE__);
    if (ld != m) {  // keep the pad row finite
        MOR_CUDA(cudaMemsetAsync(p, 0, bytes, g_ctx.stream));
    }
    mor_mat_s* a = new mor_mat_s;
    a->p = (double*)p;
    a->m = m;
    a->n = n;
    a->ld = ld;
    a->owned = true;
    *out = a;
    return MOR_OK;
}

int32_t mor_b200_mat_wrap(void* device_ptr, int64_t m, int64_t n, int64_t ld, mor_mat_t* out) {
    MOR_API_LOCK;
    MOR_ROUTE_SOLO(mor_b200_mat_wrap(device_ptr, m, n, ld, out));
    MOR_ARG(out != nullptr && device_ptr != nullptr && m >= 0 && n >= 0 && ld >= m && ld >= 1,
            "mor_b200_mat_wrap: bad arguments");
    MOR_TRY(require_ctx());
    mor_mat_s* a = new mor_mat_s;
    a->p = (double*)device_ptr;
    a->m = m;
    a->n = n;
    a->ld = ld;
    a->owned = false;
    *out = a;
    return MOR_OK;
}

int32_t mor_b200_mat_free(mor_mat_t a) {
    MOR_API_LOCK;
    MOR_ROUTE_SOLO(mor_b200_mat_free(a));
    if (!a) return MOR_OK;
    if (a->owned && a->p) {
        if (g_ctx.ready) {
            cudaSetDevice(g_ctx.device);
            cudaStreamSynchronize(g_ctx.stream);
        }
        cudaFree(a->p);
    }
    delete a;
    return MOR_OK;
}

int32_t mor_b200_mat_info(mor_mat_t a, int64_t* m, int64_t* n, int64_t* ld, void** device_ptr) {
    MOR_ARG(a != nullptr, "mor_b200_mat_info: null matrix");
    if (m) *m = a->m;
    if (n) *n = a->n;
    if (ld) *ld = a->ld;
    if (device_ptr) *device_ptr = a->p;
    return MOR_OK;
}

int32_t mor_b200_mat_upload(mor_mat_t a, const double* host, int64_t ldh) {
    MOR_API_LOCK;
    MOR_ROUTE_SOLO(mor_b200_mat_upload(a, host, ldh));
    MOR_ARG(a != nullptr && host != nullptr && ldh >= a->m, "mor_b200_mat_upload: bad arguments");
    MOR_TRY(require_ctx());
    if (a->m == 0 || a->n == 0) return MOR_OK;
    MOR_TRY(h2d_2d(a->p, a->ld, host, ldh, a->m, a->n, g_ctx.stream));
    MOR_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return MOR_OK;
}

int32_t mor_b200_mat_download(mor_mat_t a, double* host, int64_t ldh) {
    MOR_API_LOCK;
    MOR_ROUTE_SOLO(mor_b200_mat_download(a, host, ldh));
    MOR_ARG(a != nullptr && host != nullptr && ldh >= a->m, "mor_b200_mat_download: bad arguments");
    MOR_TRY(require_ctx());
    if (a->m == 0 || a->n == 0) return MOR_OK;
    MOR_TRY(d2h_2d(host, ldh, a->p, a->ld, a->m, a->n, g_ctx.stream));
    return MOR_OK;
}

int32_t mor_b200_synth(mor_mat_t a, int32_t kind, uint64_t seed, int64_t global_row0, double param,
                       int64_t iparam) {
    MOR_API_LOCK;
    MOR_ROUTE_SOLO(mor_b200_synth(a, kind, seed, global_row0, param, iparam));
    MOR_ARG(a != nullptr && global_row0 >= 0, "mor_b200_synth: bad arguments");
    MOR_TRY(require_ctx());
    MOR_TRY(synth_fill(a->p, a->m, a->n, a->ld, kind, seed, global_row0, param, iparam));
    MOR_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return MOR_OK;
}

int32_t mor_b200_kat_check(mor_mat_t U, int64_t k, int64_t n, uint64_t seed, int64_t global_row0, double param,
                           double* max_err) {
    MOR_API_LOCK;
    MOR_ROUTE_SOLO(mor_b200_kat_check(U, k, n, seed, global_row0, param, max_err));
    MOR_ARG(U != nullptr && max_err != nullptr && k >= 1 && k <= U->n && n >= k && global_row0 >= 0,
            "mor_b200_kat_check: bad arguments");
    MOR_TRY(require_ctx());
    return kat_check_left(U->p, U->m, U->ld, (int)k, n, seed, global_row0, param, max_err);
}

}  // extern "C"
