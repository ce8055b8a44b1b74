// gemm_nn: Y (m x q) = X (m x n) * W (n x q), X tall column-major FP64, W small -- forms the
// POD basis U_k = X * (V_k / sigma_k) (reference POD.jl:163 `u[:, 1:nmodes]`), the randomized
// sketch Y = X * Omega (POD.jl:27), and, with an upper-triangular W and Y aliasing X, the
// in-place CholeskyQR update X <- X * inv(R).  SURVEY.md kernels K3 and K5(a,c).
//
// Design (B200): one CTA per 128-row tile.  W is re-packed once into per-(column block,
// K chunk) slabs of 64 x 20 doubles (16 used), so a stage of the ring is sixteen 1 KB 1-D
// bulk copies of X column segments plus one 10 KB bulk copy of a W slab; the X segments land
// with a row stride of 132 doubles and the W slab has stride 20, both = 4 (mod 16) eight-byte
// banks: conflict-free DMMA.8x8x4 fragment loads.  Output column blocks are visited from
// right to left, so with a triangular W every block only reads columns at or left of itself
// and can be overwritten in place as soon as it is finished.
#include "internal.h"
#include "ptx.cuh"

namespace mor {
namespace {

constexpr int NN_RT = 128;     // rows per CTA
constexpr int NN_QT = 64;      // output columns per block
constexpr int NN_KC = 16;      // X columns per stage
constexpr int NN_RS = 132;     // shared row-tile stride (doubles)
constexpr int NN_WS = 20;      // shared W slab stride (doubles)
constexpr int NN_STAGES = 6;
constexpr int NN_XS = NN_KC * NN_RS;            // 2112 doubles
constexpr int NN_WSLAB = NN_QT * NN_WS;         // 1280 doubles
constexpr int NN_STAGE = NN_XS + NN_WSLAB;      // 3392 doubles = 27136 B
constexpr int NN_WARPS = 8;
constexpr int NN_THREADS = NN_WARPS * 32 + 32;
constexpr size_t NN_SMEM = (size_t)NN_STAGES * NN_STAGE * 8 + 2 * NN_STAGES * 8 + 128;

// Wp[(J * nchunks + c)][j][kk] = W[c*16 + kk][J*64 + j]  (zero outside W / for kk >= 16)
__global__ void __launch_bounds__(256)
pack_w_kernel(const double* __restrict__ W, int ldw, int n, int q, int nchunks, int upper,
              double* __restrict__ Wp) {
    const size_t slab = (size_t)blockIdx.x;  // J * nchunks + c
    const int J = (int)(slab / nchunks), c = (int)(slab % nchunks);
    for (int e = threadIdx.x; e < NN_WSLAB; e += blockDim.x) {
        const int j = e / NN_WS, kk = e % NN_WS;
        const int row = c * NN_KC + kk, col = J * NN_QT + j;
        double v = 0.0;
        if (kk < NN_KC && row < n && col < q && !(upper && row > col)) v = W[(size_t)col * ldw + row];
        Wp[slab * NN_WSLAB + e] = v;
    }
}

__global__ void __launch_bounds__(NN_THREADS, 1)
gemm_nn_kernel(const double* X, long long ldx, int n, const double* __restrict__ Wp, int nchunks, int q,
               long long m, double* Y, long long ldy, int upper) {  // X and Y may alias (upper)
    extern __shared__ unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)NN_STAGES * NN_STAGE);
    uint64_t* empty = full + NN_STAGES;

    const long long r0 = (long long)blockIdx.x * NN_RT;
    long long rows = m - r0;
    if (rows > NN_RT) rows = NN_RT;
    const uint32_t row_bytes = (uint32_t)(((rows + 1) & ~1LL) * 8);  // even row count: 16-byte multiple
    const int nJ = (q + NN_QT - 1) / NN_QT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NN_STAGES; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], NN_WARPS);
        }
        ptx::fence_barrier_init();
    }
    __syncthreads();

    if (warp == NN_WARPS) {
        // ---- producer warp: lanes 0..15 each copy one X column segment, lane 0 the W slab ----
        int stage = 0;
        uint32_t phase = 0;
        for (int J = nJ - 1; J >= 0; --J) {
            int kend = upper ? (J + 1) * NN_QT : n;
            if (kend > n) kend = n;
            const int cend = (kend + NN_KC - 1) / NN_KC;
            for (int c = 0; c < cend; ++c) {
                if (lane == 0) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    ptx::mbar_expect_tx(&full[stage], (uint32_t)(NN_KC * row_bytes + NN_WSLAB * 8));
                }
                __syncwarp();
                double* dst = ring + (size_t)stage * NN_STAGE;
                if (lane < NN_KC) {
                    int col = c * NN_KC + lane;
                    if (col >= n) col = n - 1;  // finite filler; its W rows are packed as zeros
                    ptx::bulk_load_1d(dst + lane * NN_RS, X + (size_t)col * ldx + r0, row_bytes, &full[stage]);
                }
                if (lane == 0)
                    ptx::bulk_load_1d(dst + NN_XS, Wp + ((size_t)J * nchunks + c) * NN_WSLAB, NN_WSLAB * 8,
                                      &full[stage]);
                if (++stage == NN_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        return;
    }

    // ---- consumer warps: 4 (rows) x 2 (columns), warp tile 32 x 32 ---------------------------
    const int wm0 = (warp & 3) * 32, wn0 = (warp >> 2) * 32;
    const int g = lane >> 2, tq = lane & 3;
    int stage = 0;
    uint32_t phase = 0;
    for (int J = nJ - 1; J >= 0; --J) {
        int kend = upper ? (J + 1) * NN_QT : n;
        if (kend > n) kend = n;
        const int cend = (kend + NN_KC - 1) / NN_KC;
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int c = 0; c < cend; ++c) {
            ptx::mbar_wait(&full[stage], phase);
            const double* Xs = ring + (size_t)stage * NN_STAGE;
            const double* ap = Xs + tq * NN_RS + wm0 + g;
            const double* bp = Xs + NN_XS + (wn0 + g) * NN_WS + tq;
#pragma unroll
            for (int s = 0; s < NN_KC / 4; ++s) {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = ap[4 * s * NN_RS + 8 * i];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = bp[8 * j * NN_WS + 4 * s];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) ptx::dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&empty[stage]);
            if (++stage == NN_STAGES) {
                stage = 0;
                phase ^= 1;
            }
        }
        // store this warp's 32 x 32 piece of block J
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long r = wm0 + 8 * i + g;
            if (r < rows) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int col = J * NN_QT + wn0 + 8 * j + 2 * tq;
                    if (col < q) Y[(size_t)col * ldy + r0 + r] = acc[i][j][0];
                    if (col + 1 < q) Y[(size_t)(col + 1) * ldy + r0 + r] = acc[i][j][1];
                }
            }
        }
    }
}

}  // namespace

int32_t gemm_nn(const double* X, int64_t ldx, int n, const double* W, int ldw, int q, int64_t m, double* Y,
                int64_t ldy, int upper) {
    MOR_ARG(X && W && Y && n >= 1 && q >= 1 && m >= 0 && ldw >= n, "gemm_nn: bad arguments");
    MOR_ARG(ldx % 2 == 0 && (uintptr_t)X % 16 == 0 && ldx >= ((m + 1) & ~int64_t(1)),
            "gemm_nn: X must be 16-byte aligned with an even leading dimension >= m rounded up to even");
    MOR_ARG(!(Y == X && !upper), "gemm_nn: in-place needs an upper-triangular W");
    MOR_ARG(!upper || q == n, "gemm_nn: triangular W must be square");
    if (m == 0) return MOR_OK;
    Ctx& cx = ctx();
    const int nchunks = (n + NN_KC - 1) / NN_KC, nJ = (q + NN_QT - 1) / NN_QT;
    double* Wp = nullptr;
    MOR_TRY(scratch(sizeof(double) * (size_t)nJ * nchunks * NN_WSLAB, reinterpret_cast<void**>(&Wp)));
    pack_w_kernel<<<(unsigned)(nJ * nchunks), 256, 0, cx.stream>>>(W, ldw, n, q, nchunks, upper, Wp);
    MOR_CUDA(cudaGetLastError());
    // per device (a process may drive several GPUs), so not cached in a static: ~1 us of host time
    MOR_CUDA(cudaFuncSetAttribute(gemm_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NN_SMEM));
    const int64_t tiles = (m + NN_RT - 1) / NN_RT;
    MOR_ARG(tiles < (int64_t(1) << 31), "gemm_nn: too many row tiles");
    gemm_nn_kernel<<<(unsigned)tiles, NN_THREADS, NN_SMEM, cx.stream>>>(X, ldx, n, Wp, nchunks, q, m, Y, ldy, upper);
    MOR_CUDA(cudaGetLastError());
    count_launch(2);
    return MOR_OK;
}

}  // namespace mor
