// Panel residual of the blocked DEIM index selection (deim.cu):
//     R (m x bw) = U[:, l0 : l0+bw] - U[:, 0 : l0] * C,      C = inv(P'U_l0) * P'U[:, l0 : l0+bw]  (l0 x bw)
// i.e. the residuals r = u_l - U_{l-1} c of reference src/deim.jl:21-23 for the next bw <= 16 basis
// columns with respect to the l0 interpolation points chosen so far, in ONE pass over the first l0
// columns (the reference, and the streaming deim_step_kernel, re-read them once per column).
//
// Design (B200): HBM-bound -- 8 m (l0 + bw) bytes read, 8 m bw written, 2 m l0 bw flop (4 flop/B at
// bw = 16, below the FP64 tensor balance of ~5.3).  Persistent CTAs (one per SM) walk row tiles; a
// producer warp streams, through an NS-stage mbarrier ring that runs ahead across tile boundaries, KC
// 1-D bulk copies per stage (KC basis columns x RT rows; default 256 rows x 16 columns = 2 KB contiguous
// runs, 32 KB per stage, 5 stages; shared row stride RT + 4 doubles) plus the matching 16 x 16 slab of C
// (stride 20), both = 4 (mod 16) eight-byte banks: conflict-free DMMA.8x8x4 fragment loads.  The last
// stages of every tile carry the bw columns U[:, l0 : l0+bw] themselves, so the epilogue (subtract, store)
// also reads from shared memory.  Eight consumer warps own RT / 8 rows x 16 columns each.
#include <cstdlib>

#include "internal.h"
#include "ptx.cuh"

namespace mor {
namespace {

constexpr int PN_BW = 16;       // panel width = columns per C slab
constexpr int PN_CS = 20;       // shared C slab stride (doubles)
constexpr int PN_SLAB = PN_BW * PN_CS;        // 320 doubles
constexpr int PN_WARPS = 8;
constexpr int PN_THREADS = PN_WARPS * 32 + 32;

// RT rows per tile x KC basis columns per stage (RT * KC = 2048 doubles = 16 KB of basis per stage):
// a stage is KC bulk copies of RT * 8 contiguous bytes each -- the longer the run, the fewer DRAM pages
// a stage touches.
template <int RT, int KC, int NS>
struct PanelCfg {
    static constexpr int RS = RT + 4;               // shared row-tile stride (doubles), = 4 (mod 16) banks
    static constexpr int XS = KC * RS;
    static constexpr int STAGE = XS + PN_SLAB;      // doubles
    static constexpr int MI = RT / (8 * PN_WARPS);  // 8-row DMMA tiles per warp
    static constexpr size_t SMEM = (size_t)NS * STAGE * 8 + 2 * NS * 8 + 128;
};

template <int RT, int KC, int NS>
__global__ void __launch_bounds__(PN_THREADS, 1)
deim_panel_kernel(const double* __restrict__ U, long long ld, int l0, int bw, const double* __restrict__ Cp,
                  long long m, double* __restrict__ R, long long ldr) {
    using Cfg = PanelCfg<RT, KC, NS>;
    constexpr int RS = Cfg::RS, XS = Cfg::XS, STAGE = Cfg::STAGE, MI = Cfg::MI, PN_STAGES = NS;
    extern __shared__ unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)PN_STAGES * STAGE);
    uint64_t* empty = full + PN_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long tiles = (m + RT - 1) / RT;
    const int nst = l0 / KC;                    // MMA stages per tile (l0 is a multiple of 16)
    const int nep = (bw + KC - 1) / KC;         // epilogue stages per tile (the panel's own columns)

    if (threadIdx.x == 0) {
        for (int s = 0; s < PN_STAGES; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], PN_WARPS);
        }
        ptx::fence_barrier_init();
    }
    __syncthreads();

    if (warp == PN_WARPS) {
        // ---- producer warp: lanes 0..KC-1 copy one column segment each, lane 0 also the C slab ----
        int stage = 0;
        uint32_t phase = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const long long r0 = tile * RT;
            long long rows = m - r0;
            if (rows > RT) rows = RT;
            const uint32_t row_bytes = (uint32_t)(((rows + 1) & ~1LL) * 8);   // 16-byte multiple
            for (int c = 0; c < nst + nep; ++c) {
                const bool ep = c >= nst;
                int ncol = KC;
                if (ep && (c - nst + 1) * KC > bw) ncol = bw - (c - nst) * KC;
                if (lane == 0) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    ptx::mbar_expect_tx(&full[stage], (uint32_t)(ncol * row_bytes + (ep ? 0 : PN_SLAB * 8)));
                }
                __syncwarp();
                double* dst = ring + (size_t)stage * STAGE;
                if (lane < ncol)
                    ptx::bulk_load_1d(dst + lane * RS, U + (size_t)(c * KC + lane) * ld + r0, row_bytes, &full[stage]);
                if (lane == 0 && !ep)
                    ptx::bulk_load_1d(dst + XS, Cp + (size_t)((c * KC) / PN_BW) * PN_SLAB, PN_SLAB * 8, &full[stage]);
                if (++stage == PN_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        return;
    }

    // ---- consumer warps: warp w owns rows 8 MI w .. 8 MI (w + 1) - 1 of the tile, all 16 columns -------
    const int wm0 = warp * 8 * MI;
    const int g = lane >> 2, tq = lane & 3;
    int stage = 0;
    uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long r0 = tile * RT;
        long long rows = m - r0;
        if (rows > RT) rows = RT;
        double acc[MI][2][2];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int c = 0; c < nst; ++c) {
            ptx::mbar_wait(&full[stage], phase);
            const double* Xs = ring + (size_t)stage * STAGE;
            const double* ap = Xs + tq * RS + wm0 + g;
            const double* bp = Xs + XS + g * PN_CS + ((c * KC) % PN_BW) + tq;
#pragma unroll
            for (int s = 0; s < KC / 4; ++s) {
                double a[MI], b[2];
#pragma unroll
                for (int i = 0; i < MI; ++i) a[i] = ap[4 * s * RS + 8 * i];
#pragma unroll
                for (int j = 0; j < 2; ++j) b[j] = bp[8 * j * PN_CS + 4 * s];
#pragma unroll
                for (int i = 0; i < MI; ++i)
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
        // epilogue stages: the panel's own columns, KC at a time; R = U_blk - acc
        for (int e = 0; e < nep; ++e) {
            ptx::mbar_wait(&full[stage], phase);
            const double* Xs = ring + (size_t)stage * STAGE;
            const int c0 = e * KC;
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const long long r = wm0 + 8 * i + g;
                if (r < rows) {
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int col = 8 * j + 2 * tq + h;
                            if (col >= c0 && col < c0 + KC && col < bw)
                                R[(size_t)col * ldr + r0 + r] = Xs[(col - c0) * RS + r] - acc[i][j][h];
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
}

template <int RT, int KC, int NS>
int32_t launch_panel(const double* U, int64_t ld, int l0, int bw, const double* Cp, int64_t m, double* R, int64_t ldr) {
    using Cfg = PanelCfg<RT, KC, NS>;
    Ctx& cx = ctx();
    // per device (a process may drive several GPUs), so not cached in a static: ~1 us of host time
    MOR_CUDA(cudaFuncSetAttribute(deim_panel_kernel<RT, KC, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)Cfg::SMEM));
    const int64_t tiles = (m + RT - 1) / RT;
    const int grid = (int)(tiles < (int64_t)cx.num_sms ? tiles : (int64_t)cx.num_sms);
    deim_panel_kernel<RT, KC, NS><<<grid, PN_THREADS, Cfg::SMEM, cx.stream>>>(U, ld, l0, bw, Cp, m, R, ldr);
    MOR_CUDA(cudaGetLastError());
    count_launch();
    return MOR_OK;
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
    // Tile shape (rows x columns per stage, stages), measured at 16M x 256 on B200 as the sum of the 15 panel
    // launches (322 GB of traffic): 128 x 16 x 6 (1 KB runs) 82.4 ms; 256 x 8 x 6 (2 KB runs) 54.5 ms; 256 x 8 x 8
    // 53.5; 256 x 8 x 10 54.4; 512 x 4 x 6 (4 KB runs, one k-step per stage) 73.2; 256 x 16 x 5 (2 KB runs, 32 KB
    // stages) 52.1 ms = 6.18 TB/s -- the default.  MOR_B200_PANEL_TILE selects the others (tests, experiments).
    const char* v = getenv("MOR_B200_PANEL_TILE");
    const int rt = v ? atoi(v) : 25616;
    if (rt == 512) return launch_panel<512, 4, 6>(U, ld, l0, bw, Cp, m, R, ldr);
    if (rt == 128) return launch_panel<128, 16, 6>(U, ld, l0, bw, Cp, m, R, ldr);
    if (rt == 256) return launch_panel<256, 8, 6>(U, ld, l0, bw, Cp, m, R, ldr);
    if (rt == 2568) return launch_panel<256, 8, 8>(U, ld, l0, bw, Cp, m, R, ldr);
    if (rt == 25610) return launch_panel<256, 8, 10>(U, ld, l0, bw, Cp, m, R, ldr);
    return launch_panel<256, 16, 5>(U, ld, l0, bw, Cp, m, R, ldr);
}

}  // namespace mor
