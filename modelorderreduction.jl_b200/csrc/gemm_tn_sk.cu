// gemm_tn for outputs wider than 64 columns: the balanced "segment" kernel (default schedule: mode 3).
//
// The plain decomposition of gemm_tn.cu -- one CTA per (128 x 128 tile, row group), 36 upper
// tiles x 4 groups = 144 CTAs at n = 1024 -- runs the DMMA pipe at ~100% on the SMs it uses (ncu:
// sm__pipe_tensor_subpipe_dmma_cycles_active 97.3% = 144/148), but leaves 4 SMs idle and computes
// the useless lower halves of the 8 diagonal tiles: 0.84 of the DMMA roofline on the kernel's own
// minimal flops.  Here a CTA executes a LIST of segments (tile, panel range); for each segment it
// zeroes its accumulators, streams the panels through the same TMA ring as gemm_tn.cu (one elected
// producer lane, PREFETCH panels ahead, running across segment boundaries) and writes a partial tile;
// a second kernel adds the partial tiles of every output tile in a fixed order (deterministic) and
// mirrors the lower triangle.  Schedules (host side, gemm_tn_streamk):
//   mode 3 (default)  lock-step first, balance second: the classic (tile, group) assignment, all CTAs
//           of a group sweeping the same rows together (one HBM fetch per panel, shared through L2),
//           but each CTA capped at the fair share total/#SMs; diagonal blocks are computed as upper
//           triangles only (run_segment_tri: 17 instead of 32 DMMAs per warp and k-step, perfectly
//           balanced over the 8 warps), and the cut-off tails of the off-diagonal tiles (13% of the
//           rows) go to the diagonal-block CTAs and to the otherwise idle SMs.  515 ms vs 563 ms.
//   mode 2  the classic schedule run through this kernel (no cap; A/B timing of the machinery: 570 ms).
//   Tile-major stream-K (705 ms) and chunk-major 64 x 64 blocks (721 ms) were measured in round 1 and removed;
//   DESIGN.md section 4 keeps the numbers.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "internal.h"
#include "ptx.cuh"

namespace mor {

int32_t make_tensor_map(const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols,
                        CUtensorMap* out);  // gemm_tn.cu

namespace {

constexpr int KB = 20, STAGES = 5, PREFETCH = STAGES - 2;
constexpr int STAGE_DOUBLES = 256 * KB;  // room for two 128-column operand blocks
constexpr size_t SK_SMEM = (size_t)STAGES * STAGE_DOUBLES * 8 + 2 * STAGES * 8 + 128;

struct Seg {
    int a_col, b_col;      // first column of the A / B operand block
    int small;             // 0: 128 x 128 tile, 1: 64 x 64 tile, 2: diagonal 128 x 128 block, upper triangle only
    int diag;              // A block == B block (only one is loaded)
    long long p_begin, p_end;
    long long part_off;    // offset (doubles) of this segment's partial tile in the workspace
};

struct TileInfo {
    int a_col, b_col, small, diag;
    long long part_base;   // first partial tile of this output tile in the workspace (doubles)
    int nparts;            // number of partial tiles, stored back to back
};

struct Maps {
    CUtensorMap a_big, b_big, a_small, b_small;
};

struct Producer {
    const Maps* maps;
    const Seg* segs;
    double* ring;
    uint64_t *full, *empty;
    int seg;            // cursor: segment and panel of the next load
    long long panel;

    // load the j-th panel of this CTA's piece into stage j % STAGES
    __device__ __forceinline__ void issue(long long j) {
        const Seg sg = segs[seg];
        const int s = (int)(j % STAGES);
        const long long fill = j / STAGES;
        if (fill > 0) ptx::mbar_wait(&empty[s], (uint32_t)((fill - 1) & 1));
        double* dst = ring + (size_t)s * STAGE_DOUBLES;
        const int bw = sg.small == 1 ? 64 : 128;
        ptx::mbar_expect_tx(&full[s], (uint32_t)((sg.diag ? bw : 2 * bw) * KB * 8));
        const int row = (int)(panel * KB);
        ptx::tma_load_2d(dst, sg.small == 1 ? &maps->a_small : &maps->a_big, row, sg.a_col, &full[s]);
        if (!sg.diag) ptx::tma_load_2d(dst + bw * KB, sg.small == 1 ? &maps->b_small : &maps->b_big, row, sg.b_col, &full[s]);
        if (++panel == sg.p_end) {
            ++seg;
            panel = segs[seg].p_begin;  // segs[] ends with a sentinel entry
        }
    }
};

template <int WM, int WN, int BMT>
__device__ __forceinline__ void run_segment(const Seg sg, double* ring, uint64_t* full, uint64_t* empty, int& stage,
                                            uint32_t& phase, long long& consumed, long long np_total, bool producer,
                                            Producer& prod, double* __restrict__ part) {
    constexpr int MI = WM / 8, NJ = WN / 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm0 = (warp & 3) * WM, wn0 = (warp >> 2) * WN;
    const int g = lane >> 2, tq = lane & 3;
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (long long p = sg.p_begin; p < sg.p_end; ++p) {
        if (producer && consumed + PREFETCH < np_total) prod.issue(consumed + PREFETCH);
        __syncwarp();
        ptx::mbar_wait(&full[stage], phase);
        const double* As = ring + (size_t)stage * STAGE_DOUBLES;
        const double* Bs = sg.diag ? As : As + BMT * KB;
        const double* ap = As + (wm0 + g) * KB + tq;
        const double* bp = Bs + (wn0 + g) * KB + tq;
#pragma unroll
        for (int s = 0; s < KB / 4; ++s) {
            double a[MI], b[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = ap[i * 8 * KB + 4 * s];
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = bp[j * 8 * KB + 4 * s];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) ptx::dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[stage]);
        if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
        }
        ++consumed;
    }
    double* out = part + sg.part_off;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int r = wm0 + 8 * i + g, c = wn0 + 8 * j + 2 * tq;
            out[(size_t)c * BMT + r] = acc[i][j][0];
            out[(size_t)(c + 1) * BMT + r] = acc[i][j][1];
        }
}

// Diagonal 128 x 128 block of a Gram matrix, upper triangle only, perfectly balanced over the 8 warps:
// in units of 8 x 8 DMMA blocks the triangle has 136 blocks; warp W takes block-rows W and 15 - W
// (16 - W and W + 1 blocks: 17 each instead of 32 for a full tile).  A == B here, so the a-fragments
// are two of the b-fragments: 16 - W shared-memory loads feed 17 DMMAs.  The eight warps run eight
// specialised copies of the loop; one run-time-predicated copy (W a variable, 24 predicated DMMA slots)
// was measured 30% slower.  Measured pace: that of 21 DMMAs per k-step (cost_of below).
template <int W>
__device__ __forceinline__ void run_segment_tri(const Seg sg, double* ring, uint64_t* full, uint64_t* empty, int& stage,
                                                uint32_t& phase, long long& consumed, long long np_total, bool producer,
                                                Producer& prod, double* __restrict__ part) {
    constexpr int N0 = 16 - W, N1 = W + 1, R1 = 15 - W;   // row W: block-cols W..15; row 15-W: block-cols 15-W..15
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    double acc0[N0][2], acc1[N1][2];
#pragma unroll
    for (int j = 0; j < N0; ++j) acc0[j][0] = acc0[j][1] = 0.0;
#pragma unroll
    for (int j = 0; j < N1; ++j) acc1[j][0] = acc1[j][1] = 0.0;
    for (long long p = sg.p_begin; p < sg.p_end; ++p) {
        if (producer && consumed + PREFETCH < np_total) prod.issue(consumed + PREFETCH);
        __syncwarp();
        ptx::mbar_wait(&full[stage], phase);
        const double* bp = ring + (size_t)stage * STAGE_DOUBLES + (8 * W + g) * KB + tq;
#pragma unroll
        for (int s = 0; s < KB / 4; ++s) {
            double b[N0];   // fragments of block-columns W..15
#pragma unroll
            for (int j = 0; j < N0; ++j) b[j] = bp[j * 8 * KB + 4 * s];
#pragma unroll
            for (int j = 0; j < N0; ++j) ptx::dmma(acc0[j][0], acc0[j][1], b[0], b[j]);
#pragma unroll
            for (int j = 0; j < N1; ++j) ptx::dmma(acc1[j][0], acc1[j][1], b[R1 - W], b[R1 - W + j]);
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[stage]);
        if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
        }
        ++consumed;
    }
    double* out = part + sg.part_off;
#pragma unroll
    for (int j = 0; j < N0; ++j) {
        const int r = 8 * W + g, c = 8 * (W + j) + 2 * tq;
        out[(size_t)c * 128 + r] = acc0[j][0];
        out[(size_t)(c + 1) * 128 + r] = acc0[j][1];
    }
#pragma unroll
    for (int j = 0; j < N1; ++j) {
        const int r = 8 * R1 + g, c = 8 * (R1 + j) + 2 * tq;
        out[(size_t)c * 128 + r] = acc1[j][0];
        out[(size_t)(c + 1) * 128 + r] = acc1[j][1];
    }
}

__global__ void __launch_bounds__(256, 1)
gemm_tn_sk_kernel(const __grid_constant__ Maps maps, const Seg* __restrict__ segs, const int* __restrict__ cta_seg,
                  double* __restrict__ part, unsigned long long* __restrict__ tlog) {
    extern __shared__ unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)STAGES * STAGE_DOUBLES);
    uint64_t* empty = full + STAGES;
    const int s0 = cta_seg[blockIdx.x], s1 = cta_seg[blockIdx.x + 1];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 8);
        }
        ptx::fence_barrier_init();
    }
    __syncthreads();
    long long np_total = 0;
    for (int s = s0; s < s1; ++s) np_total += segs[s].p_end - segs[s].p_begin;
    const bool producer = (threadIdx.x == 0);
    Producer prod{&maps, segs, ring, full, empty, s0, 0};
    if (producer && s1 > s0) {
        prod.panel = segs[s0].p_begin;
        ptx::tma_prefetch_desc(&maps.a_big);
        ptx::tma_prefetch_desc(&maps.b_big);
        ptx::tma_prefetch_desc(&maps.a_small);
        ptx::tma_prefetch_desc(&maps.b_small);
        for (long long j = 0; j < PREFETCH && j < np_total; ++j) prod.issue(j);
    }
    int stage = 0;
    uint32_t phase = 0;
    long long consumed = 0;
    unsigned long long t_start = 0, t_first = 0;
    if (tlog && threadIdx.x == 32) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
    for (int s = s0; s < s1; ++s) {
        const Seg sg = segs[s];
        if (tlog && threadIdx.x == 32 && s == s0 + 1) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_first));
        if (sg.small == 1) {
            run_segment<16, 32, 64>(sg, ring, full, empty, stage, phase, consumed, np_total, producer, prod, part);
        } else if (sg.small == 2) {
#define MOR_TRI(Wn) case Wn: run_segment_tri<Wn>(sg, ring, full, empty, stage, phase, consumed, np_total, producer, prod, part); break;
            // logical row pair W for physical warp w: the two warps that share an SM sub-partition
            // (w and w + 4) get pairs W and 7 - W, i.e. 16 - W and 9 + W fragment loads: 25 on every SMSP
            const int wphys = threadIdx.x >> 5;
            switch (wphys < 4 ? wphys : 11 - wphys) { MOR_TRI(0) MOR_TRI(1) MOR_TRI(2) MOR_TRI(3) MOR_TRI(4) MOR_TRI(5) MOR_TRI(6) MOR_TRI(7) }
#undef MOR_TRI
        } else {
            run_segment<32, 64, 128>(sg, ring, full, empty, stage, phase, consumed, np_total, producer, prod, part);
        }
    }
    if (tlog && threadIdx.x == 32) {   // debug: per-CTA timeline (start, end of the first segment, end), ns
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
        tlog[3 * blockIdx.x] = t_start;
        tlog[3 * blockIdx.x + 1] = t_first ? t_first : t_end;
        tlog[3 * blockIdx.x + 2] = t_end;
    }
}

__global__ void __launch_bounds__(256)
gemm_tn_sk_reduce_kernel(const double* __restrict__ part, const TileInfo* __restrict__ tiles, int same, int p, int q, double* __restrict__ C, int ldc,
                         int accumulate) {
    const TileInfo t = tiles[blockIdx.y];
    const int bm = t.small == 1 ? 64 : 128;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= bm * bm) return;
    const int r = e % bm, c = e / bm;
    const int gi = t.a_col + r, gj = t.b_col + c;
    if (gi >= p || gj >= q) return;
    if (same && t.diag && r > c) return;  // written by the mirrored element: exact symmetry
    double s = 0.0;
    const double* src = part + t.part_base + e;
    for (int i = 0; i < t.nparts; ++i) s += src[(size_t)i * bm * bm];
    if (accumulate) s += C[(size_t)gj * ldc + gi];
    C[(size_t)gj * ldc + gi] = s;
    if (same && gi != gj) C[(size_t)gi * ldc + gj] = s;
}

}  // namespace

// C (p x q) = A' * B, or the symmetric Gram matrix when A == B.  Requires p, q > 64.  mode: 3 (default) or 2.
int32_t gemm_tn_streamk(const double* A, int64_t lda, int p, const double* B, int64_t ldb, int q, int64_t m, double* C,
                        int ldc, int accumulate, int mode) {
    Ctx& cx = ctx();
    const bool same = (A == B && lda == ldb && p == q);
    const long long P = (m + KB - 1) / KB;
    std::vector<TileInfo> tiles;
    std::vector<Seg> segs;
    std::vector<int> cta_seg(1, 0);
    long long part_doubles = 0;
    int nct = cx.num_sms;
    {
        // Lock-step first, balance second.  Primary assignment = the classic one: G row groups x one CTA
        // per (128 x 128 tile | diagonal block), every CTA starting at the first panel of its group so the
        // group sweeps its rows together and shares them through L2.  mode 3 then caps every CTA at the
        // fair share Tq = total / #SMs: a primary CTA stops Tq worth of panels into its group, the cut-off
        // tails go to a pool that is handed to the CTAs with spare capacity -- the diagonal-block CTAs
        // (upper triangle only, 17/32 of the DMMAs of a full tile, see run_segment_tri) and the SMs the
        // classic grid leaves idle.
        // mode 2 = no cap (the classic schedule run through this kernel, for A/B timing).
        std::vector<int> slot_tiles;                 // primary slots per group -> first tile index
        std::vector<int> slot_ntiles;                // 1 (big tile) or up to 3 (split diagonal block)
        const int tp = (p + 127) / 128, tq_ = (q + 127) / 128;
        for (int ti = 0; ti < tp; ++ti)
            for (int tj = same ? ti : 0; tj < tq_; ++tj) {
                slot_tiles.push_back((int)tiles.size());
                if (same && ti == tj && mode == 3 && ti * 128 + 128 <= p) {
                    tiles.push_back({ti * 128, ti * 128, 2, 1, 0, 0});   // triangular diagonal block
                    slot_ntiles.push_back(1);
                } else {
                    tiles.push_back({ti * 128, tj * 128, 0, (same && ti == tj) ? 1 : 0, 0, 0});
                    slot_ntiles.push_back(1);
                }
            }
        const int nslots = (int)slot_tiles.size();
        int G = nct / nslots;
        if (G < 1) G = 1;
        if (G > P) G = (int)std::max<long long>(1, P);
        const int nprim = G * nslots;
        if (nprim > nct) nct = nprim;                // more tiles than SMs: several waves
        // cost per panel in DMMAs per warp per k-step; the triangular diagonal block issues 17 but was
        // measured (per-CTA timeline, MOR_B200_SK_TLOG) at the pace of 21: 16 - W loads feed only 17 DMMAs
        static const int tri_cost = getenv("MOR_B200_SK_TRI_COST") ? atoi(getenv("MOR_B200_SK_TRI_COST")) : 21;
        auto cost_of = [](const TileInfo& t) { return t.small == 1 ? 8 : (t.small == 2 ? tri_cost : 32); };
        long long total = 0;
        for (const TileInfo& t : tiles) total += cost_of(t) * P;
        const long long Tq = (mode == 3) ? (total + nct - 1) / nct : (long long)32 * P + 1;
        struct Piece { int tile; long long a, b; };
        std::vector<std::vector<Piece>> work((size_t)nct);
        std::vector<long long> left((size_t)nct, Tq);
        std::vector<Piece> pool;
        // Two-level accumulation: no register accumulator runs over more than L panels (82K rows at
        // L = 4096) before it is flushed into its own partial tile; the partial tiles are added by the
        // reduce kernel.  This bounds the sqrt(#terms) * eps growth of a 16M-term dot product (the
        // known-answer run at 16M x 1024 went from 3.6e-11 to the 1e-12 level in sigma) for the price
        // of a 128 KB store per 10 ms of DMMA work; L grows with the problem so that the partial-tile
        // workspace stays below ~1 GB.
        const long long L = std::max<long long>(4096, (P * (long long)tiles.size() + 8191) / 8192);
        auto push_piece = [&](int cta_id, int tile, long long a, long long b) {
            for (long long x = a; x < b; x += L) work[(size_t)cta_id].push_back({tile, x, std::min(b, x + L)});
        };
        for (int g = 0; g < G; ++g) {
            const long long ga = P * g / G, gb = P * (g + 1) / G;
            for (int sl = 0; sl < nslots; ++sl) {
                const int c = g * nslots + sl;
                for (int u = 0; u < slot_ntiles[sl]; ++u) {
                    const int t = slot_tiles[sl] + u;
                    const int cost = cost_of(tiles[t]);
                    long long take = std::min(gb - ga, left[c] / cost);
                    if (take > 0) {
                        push_piece(c, t, ga, ga + take);
                        left[c] -= take * cost;
                    }
                    if (ga + take < gb) pool.push_back({t, ga + take, gb});
                }
            }
        }
        // hand the pool to the CTAs with spare capacity (round-robin over them, largest spare first is
        // not needed: pieces are split freely); the last candidate takes whatever rounding leaves over
        int c = 0;
        for (Piece pc : pool) {
            const int cost = cost_of(tiles[pc.tile]);
            int guard = 0;
            while (pc.a < pc.b) {
                if (left[c] < (long long)cost * 4 && guard < nct) {   // too little room for a useful slice
                    c = (c + 1) % nct;
                    ++guard;
                    continue;
                }
                const long long take = guard >= nct ? pc.b - pc.a : std::min(pc.b - pc.a, left[c] / cost);
                push_piece(c, pc.tile, pc.a, pc.a + take);
                left[c] -= take * cost;
                pc.a += take;
                guard = 0;
            }
        }
        // partial-tile slots: contiguous per tile
        std::vector<int> count(tiles.size(), 0);
        for (const auto& wl : work)
            for (const Piece& pc : wl) ++count[(size_t)pc.tile];
        for (size_t t = 0; t < tiles.size(); ++t) {
            tiles[t].part_base = part_doubles;
            tiles[t].nparts = count[t];
            part_doubles += (long long)count[t] * (tiles[t].small == 1 ? 4096 : 16384);
            count[t] = 0;
        }
        cta_seg.clear();
        for (int cc = 0; cc < nct; ++cc) {
            cta_seg.push_back((int)segs.size());
            for (const Piece& pc : work[(size_t)cc]) {
                const TileInfo& t = tiles[(size_t)pc.tile];
                segs.push_back({t.a_col, t.b_col, t.small, t.diag, pc.a, pc.b,
                                t.part_base + (long long)count[(size_t)pc.tile]++ * (t.small == 1 ? 4096 : 16384)});
            }
        }
        cta_seg.push_back((int)segs.size());
    }
    segs.push_back({0, 0, 0, 0, 0, 0, 0});  // sentinel for the producer cursor

    Maps maps;
    MOR_TRY(make_tensor_map(A, m, p, lda, KB, 128, &maps.a_big));
    MOR_TRY(make_tensor_map(B, m, q, ldb, KB, 128, &maps.b_big));
    MOR_TRY(make_tensor_map(A, m, p, lda, KB, 64, &maps.a_small));
    MOR_TRY(make_tensor_map(B, m, q, ldb, KB, 64, &maps.b_small));

    // workspace: partial tiles | segs | cta_seg | tiles
    const size_t b_part = sizeof(double) * (size_t)part_doubles;
    const size_t b_segs = (sizeof(Seg) * segs.size() + 255) & ~size_t(255);
    const size_t b_cta = (sizeof(int) * cta_seg.size() + 255) & ~size_t(255);
    const size_t b_tiles = (sizeof(TileInfo) * tiles.size() + 255) & ~size_t(255);
    unsigned char* ws = nullptr;
    MOR_TRY(scratch(b_part + b_segs + b_cta + b_tiles + 256, reinterpret_cast<void**>(&ws)));
    double* part = reinterpret_cast<double*>(ws);
    Seg* d_segs = reinterpret_cast<Seg*>(ws + ((b_part + 255) & ~size_t(255)));
    int* d_cta = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(d_segs) + b_segs);
    TileInfo* d_tiles = reinterpret_cast<TileInfo*>(reinterpret_cast<unsigned char*>(d_cta) + b_cta);
    MOR_CUDA(cudaMemcpyAsync(d_segs, segs.data(), sizeof(Seg) * segs.size(), cudaMemcpyHostToDevice, cx.stream));
    MOR_CUDA(cudaMemcpyAsync(d_cta, cta_seg.data(), sizeof(int) * cta_seg.size(), cudaMemcpyHostToDevice, cx.stream));
    MOR_CUDA(cudaMemcpyAsync(d_tiles, tiles.data(), sizeof(TileInfo) * tiles.size(), cudaMemcpyHostToDevice, cx.stream));

    // per device (a process may drive several GPUs), so not cached in a static: ~1 us of host time
    MOR_CUDA(cudaFuncSetAttribute(gemm_tn_sk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM));
    static const bool want_tlog = getenv("MOR_B200_SK_TLOG") != nullptr;
    DevBuf tlog;
    if (want_tlog) MOR_TRY(tlog.alloc(sizeof(unsigned long long) * 3 * (size_t)nct));
    gemm_tn_sk_kernel<<<nct, 256, SK_SMEM, cx.stream>>>(maps, d_segs, d_cta, part, tlog.as<unsigned long long>());
    MOR_CUDA(cudaGetLastError());
    if (want_tlog) {
        std::vector<unsigned long long> h(3 * (size_t)nct);
        MOR_CUDA(cudaMemcpyAsync(h.data(), tlog.p, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost, cx.stream));
        MOR_CUDA(cudaStreamSynchronize(cx.stream));
        unsigned long long t0 = ~0ull;
        for (int c = 0; c < nct; ++c) t0 = std::min(t0, h[3 * (size_t)c]);
        fprintf(stderr, "[mor_b200] gemm_tn_sk mode %d: %d CTAs, %zu segments; per CTA: nseg start first_seg_end end (ms)\n", mode, nct,
                segs.size() - 1);
        for (int c = 0; c < nct; ++c)
            fprintf(stderr, "[mor_b200]   cta %3d nseg %2d  %8.3f %8.3f %8.3f\n", c, cta_seg[(size_t)c + 1] - cta_seg[(size_t)c],
                    (h[3 * (size_t)c] - t0) * 1e-6, (h[3 * (size_t)c + 1] - t0) * 1e-6, (h[3 * (size_t)c + 2] - t0) * 1e-6);
    }
    gemm_tn_sk_reduce_kernel<<<dim3(64, (unsigned)tiles.size()), 256, 0, cx.stream>>>(part, d_tiles, same ? 1 : 0,
                                                                                   p, q, C, ldc, accumulate);
    MOR_CUDA(cudaGetLastError());
    count_launch(2);
    return MOR_OK;
}

}  // namespace mor
