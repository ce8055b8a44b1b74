// Host -> device ingest probe: how fast can a PAGEABLE host matrix (what a Julia `Matrix{Float64}` is) reach HBM?
//   nvcc -O3 -o h2d_probe h2d_probe.cu -lpthread && ./h2d_probe [GiB] [ngpus]
// Strategies, each timed end to end over the same buffer (every GPU takes a contiguous share):
//   pinned      cudaHostAlloc'ed source, cudaMemcpyAsync                       (upper bound)
//   pageable    plain malloc'ed source, cudaMemcpyAsync (the driver stages through its own bounce buffer)
//   register    cudaHostRegister per window (T host threads registering ahead), DMA from the caller's pages
//   bounce      T host threads memcpy windows into pinned ring buffers, DMA from there
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>

#include <atomic>
#include <thread>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e__ = (x);                                                           \
        if (e__ != cudaSuccess) {                                                      \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

int main(int argc, char** argv) {
    const size_t gib = argc > 1 ? atoll(argv[1]) : 8;
    int ng = argc > 2 ? atoi(argv[2]) : 1;
    int have = 0;
    CK(cudaGetDeviceCount(&have));
    if (ng > have) ng = have;
    const size_t total = gib << 30, share = total / ng;
    printf("h2d_probe: %zu GiB over %d GPU(s), %u host threads available\n", gib, ng, std::thread::hardware_concurrency());
    std::vector<char*> dev(ng);
    std::vector<cudaStream_t> st(ng);
    for (int g = 0; g < ng; ++g) {
        CK(cudaSetDevice(g));
        CK(cudaMalloc(&dev[g], share));
        CK(cudaStreamCreateWithFlags(&st[g], cudaStreamNonBlocking));
    }
    char* page = (char*)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (page == MAP_FAILED) { printf("mmap failed\n"); return 1; }
    {   // touch from several threads
        std::vector<std::thread> th;
        for (int t = 0; t < 16; ++t)
            th.emplace_back([=] { memset(page + total / 16 * t, t + 1, total / 16); });
        for (auto& x : th) x.join();
    }
    auto sync_all = [&] { for (int g = 0; g < ng; ++g) { CK(cudaSetDevice(g)); CK(cudaStreamSynchronize(st[g])); } };

    // ---- pinned ----
    {
        char* pin = nullptr;
        if (cudaHostAlloc(&pin, total, cudaHostAllocPortable) == cudaSuccess) {
            memcpy(pin, page, total);
            for (int rep = 0; rep < 2; ++rep) {
                const double t0 = now();
                for (int g = 0; g < ng; ++g) { CK(cudaSetDevice(g)); CK(cudaMemcpyAsync(dev[g], pin + share * g, share, cudaMemcpyHostToDevice, st[g])); }
                sync_all();
                printf("pinned      rep %d: %7.1f GB/s\n", rep, total / (now() - t0) * 1e-9);
            }
            cudaFreeHost(pin);
        } else { cudaGetLastError(); printf("pinned: cudaHostAlloc failed\n"); }
    }
    // ---- pageable, one host thread per GPU ----
    for (int rep = 0; rep < 2; ++rep) {
        const double t0 = now();
        std::vector<std::thread> th;
        for (int g = 0; g < ng; ++g)
            th.emplace_back([&, g] { CK(cudaSetDevice(g)); CK(cudaMemcpyAsync(dev[g], page + share * g, share, cudaMemcpyHostToDevice, st[g])); CK(cudaStreamSynchronize(st[g])); });
        for (auto& x : th) x.join();
        printf("pageable    rep %d: %7.1f GB/s\n", rep, total / (now() - t0) * 1e-9);
    }
    // ---- register windows ahead of the copy ----
    for (int T : {1, 2, 4}) {
        const size_t win = size_t(256) << 20;
        const double t0 = now();
        std::vector<std::thread> th;
        for (int g = 0; g < ng; ++g)
            th.emplace_back([&, g] {
                CK(cudaSetDevice(g));
                char* base = page + share * g;
                const size_t nw = (share + win - 1) / win;
                std::vector<std::atomic<int>> ready(nw);
                for (auto& r : ready) r = 0;
                std::atomic<size_t> next{0};
                std::vector<std::thread> reg;
                for (int t = 0; t < T; ++t)
                    reg.emplace_back([&] {
                        cudaSetDevice(g);
                        for (;;) {
                            const size_t w = next.fetch_add(1);
                            if (w >= nw) break;
                            const size_t off = w * win, len = off + win <= share ? win : share - off;
                            if (cudaHostRegister(base + off, len, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); ready[w] = -1; } else ready[w] = 1;
                        }
                    });
                for (size_t w = 0; w < nw; ++w) {
                    while (ready[w] == 0) sched_yield();
                    const size_t off = w * win, len = off + win <= share ? win : share - off;
                    CK(cudaMemcpyAsync(dev[g] + off, base + off, len, cudaMemcpyHostToDevice, st[g]));
                }
                for (auto& x : reg) x.join();
                CK(cudaStreamSynchronize(st[g]));
                for (size_t w = 0; w < nw; ++w) if (ready[w] == 1) cudaHostUnregister(base + w * win);
            });
        for (auto& x : th) x.join();
        printf("register T=%d (incl. unregister): %7.1f GB/s\n", T, total / (now() - t0) * 1e-9);
    }
    // ---- bounce buffers ----
    for (int T : {2, 4, 8}) {
        const size_t win = size_t(64) << 20;
        const int NB = 4;
        std::vector<std::vector<char*>> ring(ng, std::vector<char*>(NB));
        for (int g = 0; g < ng; ++g)
            for (int b = 0; b < NB; ++b) CK(cudaHostAlloc(&ring[g][b], win, cudaHostAllocPortable));
        const double t0 = now();
        std::vector<std::thread> th;
        for (int g = 0; g < ng; ++g)
            th.emplace_back([&, g] {
                CK(cudaSetDevice(g));
                char* base = page + share * g;
                const size_t nw = (share + win - 1) / win;
                std::vector<cudaEvent_t> done(NB);
                for (auto& e : done) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                for (size_t w = 0; w < nw; ++w) {
                    const int b = (int)(w % NB);
                    if (w >= (size_t)NB) CK(cudaEventSynchronize(done[b]));
                    const size_t off = w * win, len = off + win <= share ? win : share - off;
                    std::vector<std::thread> cp;
                    for (int t = 0; t < T; ++t)
                        cp.emplace_back([=] { const size_t a = len / T * t, e = t == T - 1 ? len : len / T * (t + 1); memcpy(ring[g][b] + a, base + off + a, e - a); });
                    for (auto& x : cp) x.join();
                    CK(cudaMemcpyAsync(dev[g] + off, ring[g][b], len, cudaMemcpyHostToDevice, st[g]));
                    CK(cudaEventRecord(done[b], st[g]));
                }
                CK(cudaStreamSynchronize(st[g]));
                for (auto& e : done) cudaEventDestroy(e);
            });
        for (auto& x : th) x.join();
        printf("bounce   T=%d per GPU: %7.1f GB/s\n", T, total / (now() - t0) * 1e-9);
        for (int g = 0; g < ng; ++g)
            for (int b = 0; b < NB; ++b) cudaFreeHost(ring[g][b]);
    }
    return 0;
}
