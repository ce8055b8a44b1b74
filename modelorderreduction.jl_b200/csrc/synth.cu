// Synthetic snapshot generators: counter-based Philox4x32-10, so the
// value of entry (global row, column) depends only on the seed -- the same matrix is
// produced for any number of row shards.  HBM write-bound.
#include <cmath>
#include <vector>

#include "internal.h"

namespace mor {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

// four N(0,1) variates from one Philox block (Box-Muller in fp32; the column scale applied
// afterwards is a full FP64 number, so entries carry 53-bit mantissas)
__device__ __forceinline__ void normal4(uint4 r, float z[4]) {
    const float inv32 = 2.3283064365386963e-10f;  // 2^-32
    float u0 = ((float)r.x + 0.5f) * inv32, u1 = ((float)r.y + 0.5f) * inv32;
    float u2 = ((float)r.z + 0.5f) * inv32, u3 = ((float)r.w + 0.5f) * inv32;
    u0 = fminf(fmaxf(u0, 1e-12f), 1.0f);
    u2 = fminf(fmaxf(u2, 1e-12f), 1.0f);
    float ra = sqrtf(-2.0f * __logf(u0)), rb = sqrtf(-2.0f * __logf(u2));
    float sa, ca, sb, cb;
    __sincosf(6.283185307179586f * u1, &sa, &ca);
    __sincosf(6.283185307179586f * u3, &sb, &cb);
    z[0] = ra * ca;
    z[1] = ra * sa;
    z[2] = rb * cb;
    z[3] = rb * sb;
}

// one thread = one (block of 4 global rows, column)
__global__ void __launch_bounds__(256)
synth_kernel(double* __restrict__ X, int64_t m, int64_t n, int64_t ld, uint64_t seed, int64_t row0,
             const double* __restrict__ colscale) {
    const int64_t gb0 = row0 >> 2;                         // first global 4-row block touched
    const int64_t nblk = ((row0 + m + 3) >> 2) - gb0;      // blocks overlapping [row0, row0+m)
    const int64_t total = nblk * n;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t col = t / nblk;
        const int64_t gb = gb0 + (t - col * nblk);
        uint4 r = philox4x32_10(make_uint4((uint32_t)gb, (uint32_t)(gb >> 32), (uint32_t)col, 0x6d6f72u), key);
        float z[4];
        normal4(r, z);
        const double s = colscale[col];
        double* dst = X + col * ld;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t lr = (gb << 2) + q - row0;  // local row
            if (lr >= 0 && lr < m) dst[lr] = (double)z[q] * s;
        }
    }
}

static int32_t launch_synth(double* X, int64_t m, int64_t n, int64_t ld, uint64_t seed, int64_t row0,
                            const std::vector<double>& scale) {
    if (m == 0 || n == 0) return MOR_OK;
    Ctx& c = ctx();
    DevBuf ds;
    MOR_TRY(ds.alloc(sizeof(double) * (size_t)n));
    MOR_CUDA(cudaMemcpyAsync(ds.p, scale.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c.stream));
    int64_t total = (((row0 + m + 3) >> 2) - (row0 >> 2)) * n;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)c.num_sms * 32;
    if (blocks > cap) blocks = cap;
    synth_kernel<<<(unsigned)blocks, 256, 0, c.stream>>>(X, m, n, ld, seed, row0, ds.as<double>());
    count_launch();
    MOR_CUDA(cudaGetLastError());
    MOR_CUDA(cudaStreamSynchronize(c.stream));  // ds is released on return
    return MOR_OK;
}

int32_t synth_fill(double* X, int64_t m, int64_t n, int64_t ld, int32_t kind, uint64_t seed,
                   int64_t row0, double param, int64_t iparam) {
    std::vector<double> scale((size_t)(n > 0 ? n : 1), 1.0);
    switch (kind) {
        case MOR_SYNTH_GRADED:
            for (int64_t j = 0; j < n; ++j)
                scale[j] = n > 1 ? std::pow(10.0, -param * (double)j / (double)(n - 1)) : 1.0;
            break;
        case MOR_SYNTH_GAUSS:
            for (int64_t j = 0; j < n; ++j) scale[j] = param;
            break;
        case MOR_SYNTH_LOWRANK:
            for (int64_t j = 0; j < n; ++j) {
                if (j < iparam) scale[j] = 1.0;
                else {
                    int64_t rest = n - iparam;
                    scale[j] = 1e-3 * (rest > 1 ? std::pow(10.0, -param * (double)(j - iparam) / (double)(rest - 1)) : 1.0);
                }
            }
            break;
        default:
            set_error("mor_b200_synth: unknown generator kind");
            return MOR_ERR_ARG;
    }
    return launch_synth(X, m, n, ld, seed, row0, scale);
}

int32_t fill_normal(double* X, int64_t m, int64_t n, int64_t ld, uint64_t seed, int64_t row0, double scale) {
    std::vector<double> s((size_t)(n > 0 ? n : 1), scale);
    return launch_synth(X, m, n, ld, seed, row0, s);
}

}  // namespace mor
