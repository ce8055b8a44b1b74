// gemm_tn: C (p x q) = A' * B with A (m x p) and B (m x q) tall column-major FP64 matrices
// sharing their m rows -- the contraction over the state dimension.  With A == B it is the
// symmetric Gram matrix X'X of the snapshot factorisation (only tiles on/above the diagonal
// are computed, the rest is mirrored): SURVEY.md kernels K1 and K5(b).  This is what
// replaces the dgeqrf/dgemm work inside LinearAlgebra.svd (reference POD.jl:13) and the
// `Q'A` product of RandomizedLinAlg.rsvd (reference POD.jl:27).
//
// Design (B200): every CTA owns one BM x BN tile of C and one contiguous group of row
// panels; it keeps the tile in registers for the whole launch.  One elected lane streams
// KB-row panels of the two operand column blocks into a STAGES-deep shared-memory ring with
// 2-D TMA tensor copies (rows past m are zero-filled by the TMA unit, so ragged m needs no
// code); eight warps read DMMA.8x8x4 fragments straight from that layout (no ninth producer
// warp: a 9-warp CTA is allocated like 12 warps and would cap the kernel at 168 registers).  KB = 20
// makes the column stride 20 doubles = 4 (mod 16) eight-byte banks, which is conflict free
// for the (column g, row t) fragment pattern without padding or swizzle.  CTAs of the same
// row group are adjacent in the grid, so a panel is fetched from HBM once and served to the
// other tiles from L2.  Partial tiles are summed over row groups in a fixed order by a
// second kernel (deterministic, no atomics) which also mirrors the lower triangle.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "internal.h"
#include "ptx.cuh"

namespace mor {

int32_t gemm_tn_streamk(const double* A, int64_t lda, int p, const double* B, int64_t ldb, int q, int64_t m, double* C,
                        int ldc, int accumulate, int mode);  // gemm_tn_sk.cu

namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

int32_t get_encode(EncodeTiledFn* fn) {
    static EncodeTiledFn cached = nullptr;
    if (!cached) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        MOR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || p == nullptr) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return MOR_ERR_CUDA;
        }
        cached = reinterpret_cast<EncodeTiledFn>(p);
    }
    *fn = cached;
    return MOR_OK;
}

}  // namespace

// tensor map over a column-major (rows x cols, ld) FP64 matrix; box = box_rows x box_cols
int32_t make_tensor_map(const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols,
                 CUtensorMap* out) {
    EncodeTiledFn enc = nullptr;
    MOR_TRY(get_encode(&enc));
    cuuint64_t gdim[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(double)};
    cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)box_cols};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) +
                  " (matrix must be 16-byte aligned with an even leading dimension)");
        return MOR_ERR_CUDA;
    }
    return MOR_OK;
}

namespace {

template <int BM, int BN, int WM, int WN, int KB, int STAGES, int MINB>
struct TnCfg {
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
    static constexpr int CONSUMER_WARPS = WARPS_M * WARPS_N;
    static constexpr int THREADS = CONSUMER_WARPS * 32;
    static constexpr int MI = WM / 8, NJ = WN / 8;
    static constexpr int STAGE_DOUBLES = (BM + BN) * KB;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_DOUBLES * 8 + 2 * STAGES * 8 + 128;
};

// t -> (ti, tj), enumeration of the tiles on/above the diagonal of a T x T tile grid, row by row
__device__ __forceinline__ void upper_tile(int t, int T, int& ti, int& tj) {
    int row = 0, len = T;
    while (t >= len) {
        t -= len;
        ++row;
        --len;
    }
    ti = row;
    tj = row + t;
}

template <int BM, int BN, int WM, int WN, int KB, int STAGES, int MINB>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, MINB)
gemm_tn_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int same,
               int tiles_q, int ntiles, int ngroups, long long npanels, double* __restrict__ part) {
    using Cfg = TnCfg<BM, BN, WM, WN, KB, STAGES, MINB>;
    extern __shared__ unsigned char smem_raw[];
    // 128-byte aligned ring (TMA destination alignment)
    double* ring = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)STAGES * Cfg::STAGE_DOUBLES);
    uint64_t* empty = full + STAGES;

    const int t = blockIdx.x % ntiles;
    const int grp = blockIdx.x / ntiles;
    int ti, tj;
    if (same) {
        upper_tile(t, tiles_q, ti, tj);
    } else {
        ti = t / tiles_q;
        tj = t % tiles_q;
    }
    const bool diag = same && ti == tj;
    const long long p0 = npanels * grp / ngroups, p1 = npanels * (grp + 1) / ngroups;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], Cfg::CONSUMER_WARPS);
        }
        ptx::fence_barrier_init();
    }
    __syncthreads();

    // ---- all eight warps run DMMA; lane 0 of warp 0 also drives the TMA ring, PREFETCH
    // panels ahead, refilling the stage every warp released two iterations ago (one full
    // iteration of slack, so the producer duty never waits in steady state) ---------------
    constexpr int PREFETCH = STAGES - 2;
    const bool producer = (threadIdx.x == 0);
    const uint32_t tx_bytes = (uint32_t)((diag ? BM : BM + BN) * KB * 8);
    const long long np = p1 - p0;
    auto issue = [&](long long j) {  // j-th panel of this CTA
        const int s = (int)(j % STAGES);
        const long long fill = j / STAGES;
        if (fill > 0) ptx::mbar_wait(&empty[s], (uint32_t)((fill - 1) & 1));
        double* dst = ring + (size_t)s * Cfg::STAGE_DOUBLES;
        ptx::mbar_expect_tx(&full[s], tx_bytes);
        ptx::tma_load_2d(dst, &mapA, (int)((p0 + j) * KB), ti * BM, &full[s]);
        if (!diag) ptx::tma_load_2d(dst + BM * KB, &mapB, (int)((p0 + j) * KB), tj * BN, &full[s]);
    };
    if (producer) {
        ptx::tma_prefetch_desc(&mapA);
        if (!diag) ptx::tma_prefetch_desc(&mapB);
        for (long long j = 0; j < PREFETCH && j < np; ++j) issue(j);
    }

    const int wm0 = (warp % Cfg::WARPS_M) * WM, wn0 = (warp / Cfg::WARPS_M) * WN;
    const int g = lane >> 2, tq = lane & 3;
    double acc[Cfg::MI][Cfg::NJ][2];
#pragma unroll
    for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    int stage = 0;
    uint32_t phase = 0;
    for (long long it = 0; it < np; ++it) {
        if (producer && it + PREFETCH < np) issue(it + PREFETCH);
        __syncwarp();
        ptx::mbar_wait(&full[stage], phase);
        const double* As = ring + (size_t)stage * Cfg::STAGE_DOUBLES;
        const double* Bs = diag ? As : As + BM * KB;
        const double* ap = As + (wm0 + g) * KB + tq;
        const double* bp = Bs + (wn0 + g) * KB + tq;
#pragma unroll
        for (int s = 0; s < KB / 4; ++s) {
            double a[Cfg::MI], b[Cfg::NJ];
#pragma unroll
            for (int i = 0; i < Cfg::MI; ++i) a[i] = ap[i * 8 * KB + 4 * s];
#pragma unroll
            for (int j = 0; j < Cfg::NJ; ++j) b[j] = bp[j * 8 * KB + 4 * s];
#pragma unroll
            for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
                for (int j = 0; j < Cfg::NJ; ++j) ptx::dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[stage]);
        if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
        }
    }

    // ---- epilogue: partial tile (column-major BM x BN) to the workspace ------------------
    double* out = part + (size_t)blockIdx.x * (BM * BN);
#pragma unroll
    for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::NJ; ++j) {
            const int r = wm0 + 8 * i + g, c = wn0 + 8 * j + 2 * tq;
            out[(size_t)c * BM + r] = acc[i][j][0];
            out[(size_t)(c + 1) * BM + r] = acc[i][j][1];
        }
}

// C tile = sum over row groups (fixed order); mirror below the diagonal when symmetric
__global__ void __launch_bounds__(256)
gemm_tn_reduce_kernel(const double* __restrict__ part, int ngroups, int ntiles, int BM, int BN, int same,
                      int tiles_q, int p, int q, double* __restrict__ C, int ldc, int accumulate) {
    const int t = blockIdx.y;
    int ti, tj;
    if (same) {
        upper_tile(t, tiles_q, ti, tj);
    } else {
        ti = t / tiles_q;
        tj = t % tiles_q;
    }
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= BM * BN) return;
    const int r = e % BM, c = e / BM;
    const int gi = ti * BM + r, gj = tj * BN + c;
    if (gi >= p || gj >= q) return;
    if (same && ti == tj && r > c) return;  // written by the mirrored element: exact symmetry
    double s = 0.0;
    const double* src = part + (size_t)t * (BM * BN) + e;
    for (int gidx = 0; gidx < ngroups; ++gidx) s += src[(size_t)gidx * ntiles * (BM * BN)];
    if (accumulate) s += C[(size_t)gj * ldc + gi];
    C[(size_t)gj * ldc + gi] = s;
    if (same && gi != gj) C[(size_t)gi * ldc + gj] = s;
}

template <int BM, int BN, int WM, int WN, int KB, int STAGES, int MINB>
int32_t launch_tn(const double* A, int64_t lda, int p, const double* B, int64_t ldb, int q, int64_t m, double* C,
                  int ldc, int accumulate) {
    using Cfg = TnCfg<BM, BN, WM, WN, KB, STAGES, MINB>;
    Ctx& cx = ctx();
    const bool same = (A == B && lda == ldb && p == q);
    static_assert(BM == BN, "square tiles");
    const int tiles_p = (p + BM - 1) / BM, tiles_q = (q + BN - 1) / BN;
    const int ntiles = same ? tiles_q * (tiles_q + 1) / 2 : tiles_p * tiles_q;
    const long long npanels = (m + KB - 1) / KB;
    const int slots = cx.num_sms * MINB;
    long long ngroups = slots / ntiles;
    if (ngroups < 1) ngroups = 1;
    if (ngroups > npanels) ngroups = npanels > 0 ? npanels : 1;

    CUtensorMap mapA, mapB;
    MOR_TRY(make_tensor_map(A, m, p, lda, KB, BM, &mapA));
    MOR_TRY(make_tensor_map(B, m, q, ldb, KB, BN, &mapB));

    double* part = nullptr;
    MOR_TRY(scratch(sizeof(double) * (size_t)ngroups * ntiles * BM * BN, reinterpret_cast<void**>(&part)));

    auto kern = gemm_tn_kernel<BM, BN, WM, WN, KB, STAGES, MINB>;
    // per device (a process may drive several GPUs), so not cached in a static: ~1 us of host time
    MOR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    kern<<<(unsigned)(ngroups * ntiles), Cfg::THREADS, Cfg::SMEM, cx.stream>>>(mapA, mapB, same ? 1 : 0, tiles_q,
                                                                                 ntiles, (int)ngroups, npanels, part);
    MOR_CUDA(cudaGetLastError());
    dim3 rgrid((BM * BN + 255) / 256, ntiles);
    gemm_tn_reduce_kernel<<<rgrid, 256, 0, cx.stream>>>(part, (int)ngroups, ntiles, BM, BN, same ? 1 : 0, tiles_q, p,
                                                        q, C, ldc, accumulate);
    MOR_CUDA(cudaGetLastError());
    count_launch(2);
    return MOR_OK;
}

}  // namespace

int32_t gemm_tn(const double* A, int64_t lda, int p, const double* B, int64_t ldb, int q, int64_t m, double* C,
                int ldc, int accumulate) {
    MOR_ARG(A && B && C && p >= 1 && q >= 1 && m >= 0 && ldc >= p, "gemm_tn: bad arguments");
    MOR_ARG(lda % 2 == 0 && ldb % 2 == 0 && (uintptr_t)A % 16 == 0 && (uintptr_t)B % 16 == 0,
            "gemm_tn: operands must be 16-byte aligned with even leading dimensions");
    MOR_ARG(m < (int64_t(1) << 31), "gemm_tn: more than 2^31 rows per GPU shard");
    if (m == 0) {
        if (!accumulate) MOR_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 8, 0, (size_t)p * 8, (size_t)q, ctx().stream));
        return MOR_OK;
    }
    if (p <= 64 || q <= 64) return launch_tn<64, 64, 16, 32, 20, 5, 2>(A, lda, p, B, ldb, q, m, C, ldc, accumulate);
    // Outputs wider than 64 columns run the balanced segment kernel of gemm_tn_sk.cu, schedule 3:
    // the lock-step (tile, row group) assignment of launch_tn below, but every CTA capped at the fair
    // share of the work, diagonal blocks computed as upper triangles (17 instead of 32 DMMAs per warp
    // and k-step), and the cut-off tails handed to the diagonal-block CTAs and to the SMs the plain
    // grid leaves idle.  Measured on B200, Gram of 16M x 1024: 515 ms vs 563 ms for launch_tn.
    // MOR_B200_GEMM_TN_STREAMK (A/B timing only): "classic" = launch_tn 563 ms; "2" = the classic schedule
    // through the segment kernel 570 ms (the machinery costs 1%).
    static const char* sk = getenv("MOR_B200_GEMM_TN_STREAMK");
    if (!sk || strcmp(sk, "classic") != 0) return gemm_tn_streamk(A, lda, p, B, ldb, q, m, C, ldc, accumulate, (sk && sk[0] == '2') ? 2 : 3);
    return launch_tn<128, 128, 32, 64, 20, 5, 1>(A, lda, p, B, ldb, q, m, C, ldc, accumulate);
}

}  // namespace mor
