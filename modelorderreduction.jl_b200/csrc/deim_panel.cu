// Panel residual of the blocked DEIM index selection (deim.cu):
//     R (m x bw) = U[:, l0 : l0+bw] - U[:, 0 : l0] * C,      C = inv(P'U_l0) * P'U[:, l0 : l0+bw]  (l0 x bw)
// i.e. the residuals r = u_l - U_{l-1} c of reference src/deim.jl:21-23 for the next bw <= 16 basis
// columns with respect to the l0 interpolation points chosen so far, in ONE pass over the first l0
// columns (the reference, and the streaming deim_step_kernel, re-read them once per column).
//
// Design (B200): HBM-bound -- 8 m (l0 + bw) bytes read, 8 m bw written, 2 m l0 bw flop (4 flop/B at
// bw = 16, below the FP64 tensor balance of ~5.3).  Persistent CTAs (one per SM) walk 128-row tiles;
// a producer warp streams, through a 6-stage mbarrier ring that runs ahead across tile boundaries,
// sixteen 1 KB 1-D bulk copies per stage (16 basis columns x 128 rows, shared row stride 132 doubles)
// plus the matching 16 x 16 slab of C (stride 20), both = 4 (mod 16) eight-byte banks: conflict-free
// DMMA.8x8x4 fragment loads.  The last stage of every tile carries the bw columns U[:, l0 : l0+bw]
// themselves, so the epilogue (subtract, store) also reads from shared memory.  Eight consumer warps
// own 16 rows x 16 columns each (2 x 2 DMMA tiles).
#include "internal.h"
#include "ptx.cuh"

namespace mor {
namespace {

constexpr int PN_RT = 128;      // rows per tile
constexpr int PN_BW = 16;       // panel width = columns per stage
constexpr int PN_RS = 132;      // shared row-tile stride (doubles)
constexpr int PN_CS = 20;       // shared C slab stride (doubles)
constexpr int PN_STAGES = 6;
constexpr int PN_XS = PN_BW * PN_RS;          // 2112 doubles
constexpr int PN_SLAB = PN_BW * PN_CS;        // 320 doubles
constexpr int PN_STAGE = PN_XS + PN_SLAB;     // 2432 doubles = 19456 B
constexpr int PN_WARPS = 8;
constexpr int PN_THREADS = PN_WARPS * 32 + 32;
constexpr size_t PN_SMEM = (size_t)PN_STAGES * PN_STAGE * 8 + 2 * PN_STAGES * 8 + 128;

__global__ void __launch_bounds__(PN_THREADS, 1)
deim_panel_kernel(const double* __restrict__ U, long long ld, int l0, int bw, const double* __restrict__ Cp,
                  long long m, double* __restrict__ R, long long ldr) {
    extern __shared__ unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)PN_STAGES * PN_STAGE);
    uint64_t* empty = full + PN_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long tiles = (m + PN_RT - 1) / PN_RT;
    const int nch = l0 / PN_BW;   // l0 is a multiple of 16

    if (threadIdx.x == 0) {
        for (int s = 0; s < PN_STAGES; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], PN_WARPS);
        }
        ptx::fence_barrier_init();
    }
    __syncthreads();

    if (warp == PN_WARPS) {
        // ---- producer warp: lanes 0..15 copy one column segment each, lane 0 also the C slab ----
        int stage = 0;
        uint32_t phase = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const long long r0 = tile * PN_RT;
            long long rows = m - r0;
            if (rows > PN_RT) rows = PN_RT;
            const uint32_t row_bytes = (uint32_t)(((rows + 1) & ~1LL) * 8);   // 16-byte multiple
            for (int c = 0; c <= nch; ++c) {
                const bool last = c == nch;
                const int ncol = last ? bw : PN_BW;
                if (lane == 0) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    ptx::mbar_expect_tx(&full[stage], (uint32_t)(ncol * row_bytes + (last ? 0 : PN_SLAB * 8)));
                }
                __syncwarp();
                double* dst = ring + (size_t)stage * PN_STAGE;
                if (lane < ncol)
                    ptx::bulk_load_1d(dst + lane * PN_RS, U + (size_t)(c * PN_BW + lane) * ld + r0, row_bytes,
                                      &full[stage]);
                if (lane == 0 && !last)
                    ptx::bulk_load_1d(dst + PN_XS, Cp + (size_t)c * PN_SLAB, PN_SLAB * 8, &full[stage]);
                if (++stage == PN_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        return;
    }

    // ---- consumer warps: warp w owns rows 16 w .. 16 w + 15 of the tile, all 16 columns ----------
    const int wm0 = warp * 16;
    const int g = lane >> 2, tq = lane & 3;
    int stage = 0;
    uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long r0 = tile * PN_RT;
        long long rows = m - r0;
        if (rows > PN_RT) rows = PN_RT;
        double acc[2][2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int c = 0; c < nch; ++c) {
            ptx::mbar_wait(&full[stage], phase);
            const double* Xs = ring + (size_t)stage * PN_STAGE;
            const double* ap = Xs + tq * PN_RS + wm0 + g;
            const double* bp = Xs + PN_XS + g * PN_CS + tq;
#pragma unroll
            for (int s = 0; s < PN_BW / 4; ++s) {
                double a[2], b[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) a[i] = ap[4 * s * PN_RS + 8 * i];
#pragma unroll
                for (int j = 0; j < 2; ++j) b[j] = bp[8 * j * PN_CS + 4 * s];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) ptx::dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&empty[stage]);
            if (++stage == PN_STAGES) {
                stage = 0;
                phase ^= 1;
            }
        }
        // epilogue stage: the panel's own columns; R = U_blk - acc
        ptx::mbar_wait(&full[stage], phase);
        {
            const double* Xs = ring + (size_t)stage * PN_STAGE;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const long long r = wm0 + 8 * i + g;
                if (r < rows) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int col = 8 * j + 2 * tq;
                        if (col < bw) R[(size_t)col * ldr + r0 + r] = Xs[col * PN_RS + r] - acc[i][j][0];
                        if (col + 1 < bw) R[(size_t)(col + 1) * ldr + r0 + r] = Xs[(col + 1) * PN_RS + r] - acc[i][j][1];
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[stage]);
        if (++stage == PN_STAGES) {
            stage = 0;
            phase ^= 1;
        }
    }
}

}  // namespace

// Cp: nch = l0/16 slabs of 16 x 20 doubles, Cp[c][q][kk] = C[16 c + kk][q] (zero for kk >= 16 or q >= bw),
// written by deim_block_coef_kernel (deim.cu).
int32_t deim_panel(const double* U, int64_t ld, int l0, int bw, const double* Cp, int64_t m, double* R, int64_t ldr) {
    MOR_ARG(U && Cp && R && l0 >= PN_BW && l0 % PN_BW == 0 && bw >= 1 && bw <= PN_BW && m >= 0,
            "deim_panel: bad arguments");
    MOR_ARG(ld % 2 == 0 && (uintptr_t)U % 16 == 0 && ld >= ((m + 1) & ~int64_t(1)),
            "deim_panel: the basis must be 16-byte aligned with an even leading dimension >= m rounded up to even");
    if (m == 0) return MOR_OK;
    Ctx& cx = ctx();
    static bool attr_set = false;
    if (!attr_set) {
        MOR_CUDA(cudaFuncSetAttribute(deim_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PN_SMEM));
        attr_set = true;
    }
    const int64_t tiles = (m + PN_RT - 1) / PN_RT;
    const int grid = (int)(tiles < (int64_t)cx.num_sms ? tiles : (int64_t)cx.num_sms);
    deim_panel_kernel<<<grid, PN_THREADS, PN_SMEM, cx.stream>>>(U, ld, l0, bw, Cp, m, R, ldr);
    MOR_CUDA(cudaGetLastError());
    count_launch();
    return MOR_OK;
}

}  // namespace mor
