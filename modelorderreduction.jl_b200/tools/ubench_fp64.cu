// Micro-benchmarks that calibrate the FP64 rooflines on the box (no FP64 figure exists in
// MEASURED_PEAKS.json): DMMA.8x8x4 issue rate, DFMA rate, HBM read / copy bandwidth.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp64 ubench_fp64.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a0, double b0) {
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i][0] = threadIdx.x; acc[i][1] = i; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(acc[i][0], acc[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a0, double b0) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
    double a = a0, b = b0 + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_read(const double2* __restrict__ p, size_t n2, double* out) {
    double s = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * st < n2; i += 4 * st) {
        double2 a = p[i], b = p[i + st], c = p[i + 2 * st], d = p[i + 3 * st];
        s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
    }
    for (; i < n2; i += st) { double2 a = p[i]; s += a.x + a.y; }
    if (s == 1.2345e300) out[0] = s;
}
__global__ void __launch_bounds__(256) k_copy(const double2* __restrict__ p, double2* __restrict__ q, size_t n2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += st) q[i] = p[i];
}

template <class F> float time_ms(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sms %d clock %.0f MHz\n", prop.name, sms, prop.clockRate / 1000.0);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 4096;
    for (int warps_cfg = 0; warps_cfg < 3; ++warps_cfg) {
        int threads = warps_cfg == 0 ? 128 : (warps_cfg == 1 ? 256 : 512);
        int bps = warps_cfg == 2 ? 2 : 4;
        {
            float ms = time_ms([&] { k_dmma<8><<<sms * bps, threads>>>(out, iters, 1.0, 1e-3); }, 5);
            double flops = (double)sms * bps * (threads / 32) * iters * 8.0 * 512.0;
            printf("DMMA m8n8k4  threads/blk %d blk/sm %d : %.3f ms  %.2f TFLOP/s\n", threads, bps, ms, flops / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { k_dfma<16><<<sms * bps, threads>>>(out, iters, 0.999, 1e-3); }, 5);
            double flops = (double)sms * bps * threads * iters * 16.0 * 2.0;
            printf("DFMA         threads/blk %d blk/sm %d : %.3f ms  %.2f TFLOP/s\n", threads, bps, ms, flops / ms * 1e-9);
        }
    }
    size_t bytes = (size_t)8 << 30;
    double2 *p, *q; CK(cudaMalloc(&p, bytes)); CK(cudaMalloc(&q, bytes));
    CK(cudaMemset(p, 0, bytes)); CK(cudaMemset(q, 0, bytes));
    size_t n2 = bytes / 16;
    float ms = time_ms([&] { k_read<<<sms * 16, 256>>>(p, n2, out); }, 5);
    printf("HBM read  8 GiB: %.3f ms  %.1f GB/s\n", ms, bytes / ms * 1e-6);
    ms = time_ms([&] { k_copy<<<sms * 16, 256>>>(p, q, n2); }, 5);
    printf("HBM copy  8 GiB: %.3f ms  %.1f GB/s (read+write)\n", ms, 2.0 * bytes / ms * 1e-6);
    return 0;
}
