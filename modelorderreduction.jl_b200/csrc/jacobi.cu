// One-sided (Hestenes) Jacobi SVD of the small replicated core: the n x n triangular factor
// of the snapshot matrix or the (n x l) transposed projection of the randomized SVD.  This is
// the piece of LinearAlgebra.svd / svd!(B) (reference POD.jl:13,27) that is left after the
// tall-skinny reduction (SURVEY.md kernel K6).
//
// A sweep is n-1 rounds of a round-robin tournament; the n/2 column pairs of a round are
// disjoint and rotated concurrently, one CTA per pair: three dot products (a, b, c), the
// rotation that zeroes c, applied to the two columns of A and of the accumulated V.  The
// rotation threshold is |c| <= tol * sqrt(a b) with tol = rows * eps, which gives the singular
// values to high relative accuracy (the neglected cosines enter sigma only at second order).
//
// A whole sweep is ONE cooperative launch: the CTAs stay resident, keep the four columns of
// a pair in registers between the dot products and the rotation (one L2 round trip instead
// of two), and meet at a grid barrier between rounds instead of paying a kernel launch per
// round (measured: 12.5 us per round with per-round launches).  Loads/stores bypass L1
// (ld.cg / st.cg): other SMs rewrite these columns every round.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <numeric>
#include <vector>

#include <cooperative_groups.h>

#include "internal.h"

namespace mor {
namespace {

constexpr int JT = 128;  // threads per pair

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < JT / 32; ++i) s += red[i];
    __syncthreads();
    return s;
}

namespace cg = cooperative_groups;
constexpr int ST = 256;  // threads per CTA of the sweep kernel

__device__ __forceinline__ void pair_of(int b, int round, int N1, int& p, int& q) {
    if (b == 0) {
        p = N1;
        q = round;
    } else {
        p = (round + b) % N1;
        q = (round - b + N1) % N1;
    }
    if (p > q) {
        const int t = p;
        p = q;
        q = t;
    }
}

// R = elements per thread per column held in registers (rows, vrows <= R * ST); R = 0 streams
template <int R>
__global__ void __launch_bounds__(ST, R == 4 ? 4 : 1)   // R = 4: 512 resident CTAs = all pairs of n = 1024
jacobi_sweep_kernel(double* __restrict__ A, int lda, int rows, double* __restrict__ V, int ldv, int vrows, int n,
                    int n_even, double tol, SmallInfo* __restrict__ info) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double red[3][ST / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N1 = n_even - 1, npairs = n_even / 2;
    int my_rot = 0;
    double my_maxcos = 0.0;
    for (int round = 0; round < N1; ++round) {
        for (int b = blockIdx.x; b < npairs; b += gridDim.x) {
            int p, q;
            pair_of(b, round, N1, p, q);
            if (q >= n) continue;  // the bye of an odd tournament (uniform per CTA)
            double* ap = A + (size_t)p * lda;
            double* aq = A + (size_t)q * lda;
            double* vp = V ? V + (size_t)p * ldv : nullptr;
            double* vq = V ? V + (size_t)q * ldv : nullptr;
            constexpr int RR = R > 0 ? R : 1;
            double xa[RR], ya[RR], xv[RR], yv[RR];
            double sa = 0.0, sb = 0.0, sc = 0.0;
            if (R > 0) {
#pragma unroll
                for (int i = 0; i < RR; ++i) {
                    const int r = tid + i * ST;
                    xa[i] = r < rows ? __ldcg(ap + r) : 0.0;
                    ya[i] = r < rows ? __ldcg(aq + r) : 0.0;
                }
                if (V) {
#pragma unroll
                    for (int i = 0; i < RR; ++i) {
                        const int r = tid + i * ST;
                        xv[i] = r < vrows ? __ldcg(vp + r) : 0.0;
                        yv[i] = r < vrows ? __ldcg(vq + r) : 0.0;
                    }
                }
#pragma unroll
                for (int i = 0; i < RR; ++i) {
                    sa = fma(xa[i], xa[i], sa);
                    sb = fma(ya[i], ya[i], sb);
                    sc = fma(xa[i], ya[i], sc);
                }
            } else {
                for (int r = tid; r < rows; r += ST) {
                    const double x = __ldcg(ap + r), y = __ldcg(aq + r);
                    sa = fma(x, x, sa);
                    sb = fma(y, y, sb);
                    sc = fma(x, y, sc);
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                sa += __shfl_xor_sync(0xffffffffu, sa, off);
                sb += __shfl_xor_sync(0xffffffffu, sb, off);
                sc += __shfl_xor_sync(0xffffffffu, sc, off);
            }
            if (lane == 0) {
                red[0][warp] = sa;
                red[1][warp] = sb;
                red[2][warp] = sc;
            }
            __syncthreads();
            double a = 0.0, bb = 0.0, c = 0.0;
#pragma unroll
            for (int w = 0; w < ST / 32; ++w) {
                a += red[0][w];
                bb += red[1][w];
                c += red[2][w];
            }
            __syncthreads();  // red is reused by the next pair of this CTA
            const double denom = sqrt(a) * sqrt(bb);
            if (!(fabs(c) > tol * denom)) continue;  // also leaves NaN / zero columns alone
            if (tid == 0) {
                ++my_rot;
                my_maxcos = fmax(my_maxcos, fabs(c) / denom);
            }
            const double zeta = (bb - a) / (2.0 * c);
            const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
            if (R > 0) {
#pragma unroll
                for (int i = 0; i < RR; ++i) {
                    const int r = tid + i * ST;
                    if (r < rows) {
                        __stcg(ap + r, cs * xa[i] - sn * ya[i]);
                        __stcg(aq + r, sn * xa[i] + cs * ya[i]);
                    }
                }
                if (V) {
#pragma unroll
                    for (int i = 0; i < RR; ++i) {
                        const int r = tid + i * ST;
                        if (r < vrows) {
                            __stcg(vp + r, cs * xv[i] - sn * yv[i]);
                            __stcg(vq + r, sn * xv[i] + cs * yv[i]);
                        }
                    }
                }
            } else {
                for (int r = tid; r < rows; r += ST) {
                    const double x = __ldcg(ap + r), y = __ldcg(aq + r);
                    __stcg(ap + r, cs * x - sn * y);
                    __stcg(aq + r, sn * x + cs * y);
                }
                if (V) {
                    for (int r = tid; r < vrows; r += ST) {
                        const double x = __ldcg(vp + r), y = __ldcg(vq + r);
                        __stcg(vp + r, cs * x - sn * y);
                        __stcg(vq + r, sn * x + cs * y);
                    }
                }
            }
        }
        grid.sync();
    }
    if (tid == 0 && my_rot > 0) {
        atomicAdd(&info->rotations, my_rot);
        atomicMax(reinterpret_cast<unsigned long long*>(&info->max_cos), (unsigned long long)__double_as_longlong(my_maxcos));
    }
}

template <int R>
int32_t launch_sweep(double* A, int lda, int rows, double* V, int ldv, int vrows, int n, int n_even, double tol,
                     SmallInfo* d_info) {
    Ctx& cx = ctx();
    auto kern = jacobi_sweep_kernel<R>;
    int per_sm = 0;
    MOR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ST, 0));
    if (per_sm < 1) {
        set_error("jacobi_sweep_kernel cannot be made resident");
        return MOR_ERR_CUDA;
    }
    int grid = n_even / 2;
    if (grid > per_sm * cx.num_sms) grid = per_sm * cx.num_sms;
    void* args[] = {&A, &lda, &rows, &V, &ldv, &vrows, &n, &n_even, &tol, &d_info};
    MOR_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(ST), args, 0, cx.stream));
    count_launch();
    return MOR_OK;
}

// ---- blocked sweep (no V, rows <= 4 * ST): a CTA owns a PAIR OF 4-COLUMN BLOCKS --------------
// The pair-per-CTA sweep above moves every column through L2 once per rotation (16 MB per round
// at n = 1024, ~4.6 us per round, 1023 rounds).  Here the tournament runs over blocks of 4 columns:
// a CTA keeps the 8 columns of its block pair in registers, performs all 16 cross rotations (4 steps
// of 4 disjoint pairs, one batched 12-value block reduction per step) and only then writes them
// back -- 4x fewer rounds and 4x less L2 traffic per rotation.  The 6 + 6 pairs inside the two
// blocks are rotated in the first round of the sweep (3 extra steps).
constexpr int BC = 4;  // columns per block

struct StepPairs {
    int u[4], v[4];
};
// steps 0..3: cross pairs (a, 4 + (a + s) % 4); steps 4..6: pairs inside both blocks
__device__ constexpr StepPairs kSteps[7] = {
    {{0, 1, 2, 3}, {4, 5, 6, 7}}, {{0, 1, 2, 3}, {5, 6, 7, 4}}, {{0, 1, 2, 3}, {6, 7, 4, 5}}, {{0, 1, 2, 3}, {7, 4, 5, 6}},
    {{0, 2, 4, 6}, {1, 3, 5, 7}}, {{0, 1, 4, 5}, {2, 3, 6, 7}}, {{0, 1, 4, 5}, {3, 2, 7, 6}}};

template <int S>
__device__ __forceinline__ void block_step(double (&c)[8][4], double (*red)[12][ST / 32], int buf, double tol,
                                           int& my_rot, double& my_maxcos) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double sum[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        constexpr StepPairs sp = kSteps[S];
        const int u = sp.u[k], v = sp.v[k];
        double sa = 0.0, sb = 0.0, sc = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            sa = fma(c[u][i], c[u][i], sa);
            sb = fma(c[v][i], c[v][i], sb);
            sc = fma(c[u][i], c[v][i], sc);
        }
        sum[3 * k] = sa;
        sum[3 * k + 1] = sb;
        sum[3 * k + 2] = sc;
    }
#pragma unroll
    for (int t = 0; t < 12; ++t) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum[t] += __shfl_xor_sync(0xffffffffu, sum[t], off);
    }
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < 12; ++t) red[buf][t][warp] = sum[t];
    }
    __syncthreads();
    // lane k < 4 of every warp works out the rotation of pair k (the FP64 sqrt / divide chain is the
    // expensive part of a step: done once per pair in parallel lanes, not 4x sequentially by all
    // threads), then the (cs, sn) pairs are broadcast with shuffles -- no second block barrier.
    double cs = 1.0, sn = 0.0;
    {
        const int k = lane & 3;
        double a = 0.0, bb = 0.0, cc = 0.0;
#pragma unroll
        for (int w = 0; w < ST / 32; ++w) {
            a += red[buf][3 * k][w];
            bb += red[buf][3 * k + 1][w];
            cc += red[buf][3 * k + 2][w];
        }
        const double denom = sqrt(a) * sqrt(bb);
        if (fabs(cc) > tol * denom) {
            const double zeta = (bb - a) / (2.0 * cc);
            const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            cs = rsqrt(1.0 + t * t);
            sn = cs * t;
            if (threadIdx.x < 4) {
                ++my_rot;
                my_maxcos = fmax(my_maxcos, fabs(cc) / denom);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        constexpr StepPairs sp = kSteps[S];
        const int u = sp.u[k], v = sp.v[k];
        const double ck = __shfl_sync(0xffffffffu, cs, k), sk = __shfl_sync(0xffffffffu, sn, k);
        if (sk != 0.0) {  // uniform across the CTA: every warp derives it from the same sums
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double x = c[u][i], y = c[v][i];
                c[u][i] = ck * x - sk * y;
                c[v][i] = sk * x + ck * y;
            }
        }
    }
}

__global__ void __launch_bounds__(ST, 2)
jacobi_block_sweep_kernel(double* __restrict__ A, int lda, int rows, int n, int nblk_even, double tol,
                          SmallInfo* __restrict__ info) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double red[2][12][ST / 32];
    const int tid = threadIdx.x;
    const int N1 = nblk_even - 1, npairs = nblk_even / 2;
    int my_rot = 0;
    double my_maxcos = 0.0;
    for (int round = 0; round < N1; ++round) {
        for (int b = blockIdx.x; b < npairs; b += gridDim.x) {
            int I, J;
            pair_of(b, round, N1, I, J);
            double c[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = (j < BC ? I * BC + j : J * BC + (j - BC));
                const double* src = A + (size_t)col * lda;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = tid + i * ST;
                    c[j][i] = (col < n && r < rows) ? __ldcg(src + r) : 0.0;
                }
            }
            if (round == 0) {  // pairs inside the two blocks: once per sweep
                block_step<4>(c, red, 0, tol, my_rot, my_maxcos);
                block_step<5>(c, red, 1, tol, my_rot, my_maxcos);
                block_step<6>(c, red, 0, tol, my_rot, my_maxcos);
                __syncthreads();
            }
            block_step<0>(c, red, 1, tol, my_rot, my_maxcos);
            block_step<1>(c, red, 0, tol, my_rot, my_maxcos);
            block_step<2>(c, red, 1, tol, my_rot, my_maxcos);
            block_step<3>(c, red, 0, tol, my_rot, my_maxcos);
            __syncthreads();  // red[] is reused by the next block pair of this CTA
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = (j < BC ? I * BC + j : J * BC + (j - BC));
                double* dst = A + (size_t)col * lda;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = tid + i * ST;
                    if (col < n && r < rows) __stcg(dst + r, c[j][i]);
                }
            }
        }
        grid.sync();
    }
    if (tid < 4 && my_rot > 0) {
        atomicAdd(&info->rotations, my_rot);
        atomicMax(reinterpret_cast<unsigned long long*>(&info->max_cos), (unsigned long long)__double_as_longlong(my_maxcos));
    }
}

// ---- blocked sweep, Gram form ------------------------------------------------------------------------------------
// Same tournament, same rotations, one block-wide reduction per round instead of four.  The four steps of
// jacobi_block_sweep_kernel each need the three dot products of four column pairs: reduce -> barrier -> rotate, four
// times in a row.  Here the whole 8 x 8 Gram matrix G = C'C of the block pair (36 dot products) is reduced ONCE; every
// warp then runs the same rotation sequence on its own copy of G in shared memory -- a rotation of columns (u, v) acts
// on G as G <- R'GR, so the later steps of the round see exactly the dot products they would recompute -- accumulates
// the 8 x 8 orthogonal J = R1 R2 ..., and each thread applies J to the 8 x 4 values it holds (C <- C J).  The round's
// dependent chain shrinks from 4 x (reduce + barrier + sqrt/div chain) to 1 reduce + barrier + 4 short chains.
// Rounding: the entries of G are updated algebraically inside a round (errors of a few eps relative to |a_u||a_v|,
// far below the rotation threshold rows * eps) and recomputed from the columns in the next round.
__global__ void __launch_bounds__(ST, 1)
jacobi_block_sweep_gram_kernel(double* __restrict__ A, int lda, int rows, int n, int nblk_even, double tol,
                               SmallInfo* __restrict__ info, int* __restrict__ ver, int epoch_base) {
    // No grid barrier between the rounds: a block pair only depends on the two CTAs that held its blocks in the
    // previous round.  ver[X] counts the rounds block X has been through (monotone over the whole SVD); the owner of
    // a round publishes it (release) after its stores, the next owner waits for it (acquire).  The launch stays
    // cooperative so that all CTAs are co-resident and the waits cannot deadlock.
    __shared__ double part[ST / 32][36];
    __shared__ double Gw[ST / 32][64], Jw[ST / 32][64], Tw[ST / 32][64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N1 = nblk_even - 1, npairs = nblk_even / 2;
    int my_rot = 0;
    double my_maxcos = 0.0;
    double* G = Gw[warp];
    double* J = Jw[warp];
    double* T = Tw[warp];
    for (int round = 0; round < N1; ++round) {
        for (int b = blockIdx.x; b < npairs; b += gridDim.x) {
            int I, Jb;
            pair_of(b, round, N1, I, Jb);
            const int need = epoch_base + round;
            if (tid < 2) {   // both blocks have been through the previous round
                const int* vp = ver + (tid == 0 ? I : Jb);
                const long long t0 = clock64();
                int seen;
                do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(vp) : "memory");
                    if (seen < need && clock64() - t0 > (1LL << 30)) {   // ~0.5 s: report, never hang
                        atomicExch(&info->sync_fail, 1);
                        break;
                    }
                } while (seen < need);
            }
            __syncthreads();
            double c[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = (j < BC ? I * BC + j : Jb * BC + (j - BC));
                const double* src = A + (size_t)col * lda;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = tid + i * ST;
                    c[j][i] = (col < n && r < rows) ? __ldcg(src + r) : 0.0;
                }
            }
            // the 36 dot products of the 8 columns, reduced over the CTA in one go
            {
                double g[36];
                int e = 0;
#pragma unroll
                for (int u = 0; u < 8; ++u)
#pragma unroll
                    for (int v = u; v < 8; ++v) {
                        double sacc = 0.0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) sacc = fma(c[u][i], c[v][i], sacc);
                        g[e++] = sacc;
                    }
#pragma unroll
                for (int t = 0; t < 36; ++t) {
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) g[t] += __shfl_xor_sync(0xffffffffu, g[t], off);
                }
                if (lane == 0) {
#pragma unroll
                    for (int t = 0; t < 36; ++t) part[warp][t] = g[t];
                }
            }
            __syncthreads();
            // every warp builds its own copy of G (and J = I) and runs the same rotation sequence on it
            for (int t = lane; t < 36; t += 32) {
                double sacc = 0.0;
#pragma unroll
                for (int w = 0; w < ST / 32; ++w) sacc += part[w][t];
                // unpack the upper-triangle index t -> (u, v)
                int u = 0, rem = t;
                while (rem >= 8 - u) {
                    rem -= 8 - u;
                    ++u;
                }
                const int v = u + rem;
                G[u * 8 + v] = sacc;
                G[v * 8 + u] = sacc;
            }
            J[lane] = (lane >> 3) == (lane & 7) ? 1.0 : 0.0;
            J[lane + 32] = ((lane + 32) >> 3) == (lane & 7) ? 1.0 : 0.0;
            __syncwarp();
            bool any_rot = false;
            const int first = round == 0 ? 0 : 3;   // sequence 4, 5, 6 (pairs inside the two blocks, once per sweep), then 0, 1, 2, 3
            for (int q = first; q < 7; ++q) {
                const int S = q < 3 ? q + 4 : q - 3;
                const int k = lane & 3, i8 = lane >> 2;
                const int u = kSteps[S].u[k], v = kSteps[S].v[k];
                double cs = 1.0, sn = 0.0;
                {
                    const double a = G[u * 8 + u], bb = G[v * 8 + v], cc = G[u * 8 + v];
                    const double denom = sqrt(a) * sqrt(bb);
                    if (fabs(cc) > tol * denom) {
                        const double zeta = (bb - a) / (2.0 * cc);
                        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        cs = rsqrt(1.0 + t * t);
                        sn = cs * t;
                        if (tid < 4) {
                            ++my_rot;
                            my_maxcos = fmax(my_maxcos, fabs(cc) / denom);
                        }
                    }
                }
                // (all lanes with the same k computed the same (cs, sn): no broadcast needed)
                if (__ballot_sync(0xffffffffu, sn != 0.0) == 0u) continue;
                any_rot = true;
                // T = G R (columns u, v of row i8), J <- J R
                {
                    const double gu = G[i8 * 8 + u], gv = G[i8 * 8 + v];
                    T[i8 * 8 + u] = cs * gu - sn * gv;
                    T[i8 * 8 + v] = sn * gu + cs * gv;
                    const double ju = J[i8 * 8 + u], jv = J[i8 * 8 + v];
                    J[i8 * 8 + u] = cs * ju - sn * jv;
                    J[i8 * 8 + v] = sn * ju + cs * jv;
                }
                __syncwarp();
                // G = R' T (rows u, v of column i8)
                {
                    const double tu = T[u * 8 + i8], tv = T[v * 8 + i8];
                    G[u * 8 + i8] = cs * tu - sn * tv;
                    G[v * 8 + i8] = sn * tu + cs * tv;
                }
                __syncwarp();
            }
            if (any_rot) {   // C <- C J on the 8 x 4 values of this thread
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double x[8];
#pragma unroll
                    for (int h = 0; h < 8; ++h) x[h] = c[h][i];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        double sacc = 0.0;
#pragma unroll
                        for (int h = 0; h < 8; ++h) sacc = fma(x[h], J[h * 8 + j], sacc);
                        c[j][i] = sacc;
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int col = (j < BC ? I * BC + j : Jb * BC + (j - BC));
                    double* dst = A + (size_t)col * lda;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = tid + i * ST;
                        if (col < n && r < rows) __stcg(dst + r, c[j][i]);
                    }
                }
            }
            __threadfence();  // this thread's column stores, before the versions are published
            __syncthreads();  // (part[] is also reused by the next block pair of this CTA)
            if (tid < 2) {
                int* vp = ver + (tid == 0 ? I : Jb);
                asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(vp), "r"(need + 1) : "memory");
            }
        }
    }
    if (tid < 4 && my_rot > 0) {
        atomicAdd(&info->rotations, my_rot);
        atomicMax(reinterpret_cast<unsigned long long*>(&info->max_cos), (unsigned long long)__double_as_longlong(my_maxcos));
    }
}

int32_t launch_block_sweep(double* A, int lda, int rows, int n, double tol, SmallInfo* d_info, int* d_ver, int sweep,
                           bool gram_form) {
    Ctx& cx = ctx();
    const int nblk = (n + BC - 1) / BC;
    int nblk_even = (nblk + 1) & ~1;
    void* kern = gram_form ? (void*)jacobi_block_sweep_gram_kernel : (void*)jacobi_block_sweep_kernel;
    int per_sm = 0;
    if (gram_form) MOR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_block_sweep_gram_kernel, ST, 0));
    else MOR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_block_sweep_kernel, ST, 0));
    if (per_sm < 1) {
        set_error("jacobi_block_sweep_kernel cannot be made resident");
        return MOR_ERR_CUDA;
    }
    int grid = nblk_even / 2;
    if (grid > per_sm * cx.num_sms) grid = per_sm * cx.num_sms;
    int epoch_base = sweep * (nblk_even - 1);   // every block goes through nblk_even - 1 rounds per sweep
    void* args[] = {&A, &lda, &rows, &n, &nblk_even, &tol, &d_info, &d_ver, &epoch_base};
    MOR_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(ST), args, 0, cx.stream));
    count_launch();
    return MOR_OK;
}

__global__ void __launch_bounds__(JT)
colnorm_kernel(const double* __restrict__ A, int lda, int rows, double* __restrict__ out) {
    __shared__ double red[JT / 32];
    const double* a = A + (size_t)blockIdx.x * lda;
    // scaled two-pass norm: immune to overflow/underflow of the squares
    double mx = 0.0;
    for (int r = threadIdx.x; r < rows; r += JT) mx = fmax(mx, fabs(a[r]));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = 0.0;
#pragma unroll
    for (int i = 0; i < JT / 32; ++i) mx = fmax(mx, red[i]);
    __syncthreads();
    if (!(mx > 0.0) || !isfinite(mx)) {
        if (threadIdx.x == 0) out[blockIdx.x] = mx;
        return;
    }
    double s = 0.0;
    for (int r = threadIdx.x; r < rows; r += JT) {
        const double x = a[r] / mx;
        s = fma(x, x, s);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[blockIdx.x] = mx * sqrt(s);
}

// out[:, j] = in[:, perm[j]] * (scale ? scale[j] : 1)
__global__ void gather_cols_kernel(const double* __restrict__ in, int ldin, int rows, const int* __restrict__ perm,
                                   const double* __restrict__ scale, double* __restrict__ out, int ldout) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= rows) return;
    const double s = scale ? scale[j] : 1.0;
    out[(size_t)j * ldout + i] = in[(size_t)perm[j] * ldin + i] * s;
}

}  // namespace

int32_t gather_cols(const double* in, int ldin, int rows, int ncols, const int* d_perm, const double* d_scale,
                    double* out, int ldout) {
    if (rows <= 0 || ncols <= 0) return MOR_OK;
    gather_cols_kernel<<<dim3((rows + 127) / 128, ncols), 128, 0, ctx().stream>>>(in, ldin, rows, d_perm, d_scale,
                                                                                 out, ldout);
    count_launch();
    MOR_CUDA(cudaGetLastError());
    return MOR_OK;
}

int32_t col_norms(const double* A, int lda, int rows, int n, std::vector<double>& out) {
    DevBuf norms;
    MOR_TRY(norms.alloc(sizeof(double) * (size_t)n));
    colnorm_kernel<<<n, JT, 0, ctx().stream>>>(A, lda, rows, norms.as<double>());
    count_launch();
    MOR_CUDA(cudaGetLastError());
    out.resize((size_t)n);
    MOR_CUDA(cudaMemcpyAsync(out.data(), norms.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx().stream));
    MOR_CUDA(cudaStreamSynchronize(ctx().stream));
    return MOR_OK;
}

int32_t jacobi_svd(double* A, int rows, int n, int lda, double* V, int ldv, std::vector<double>& sigma,
                   std::vector<int>& order, int* sweeps_out, bool well_scaled) {
    Ctx& cx = ctx();
    MOR_ARG(A && rows >= 1 && n >= 1 && lda >= rows, "jacobi_svd: bad arguments");
    DevBuf info_buf, norms;
    MOR_TRY(info_buf.alloc(sizeof(SmallInfo)));
    MOR_TRY(norms.alloc(sizeof(double) * (size_t)n));
    SmallInfo* d_info = info_buf.as<SmallInfo>();
    DevBuf ver_buf;   // block versions of the barrier-free blocked sweep
    MOR_TRY(ver_buf.alloc(sizeof(int) * (size_t)(n / BC + 4)));
    MOR_CUDA(cudaMemsetAsync(ver_buf.p, 0, ver_buf.bytes, cx.stream));
    if (V) MOR_TRY(small_identity(V, n, ldv));
    const int n_even = (n + 1) & ~1;
    // rotate while |cos(a_p, a_q)| > tol.  dgesvj uses sqrt(rows)*eps; the computed dot product of a
    // tiny and a large column carries rounding noise of about that size, which keeps triggering
    // rotations for tens of sweeps on graded matrices (30 sweeps at 256 x 256, kappa 1e8), so the
    // threshold sits a factor sqrt(rows) above the noise.  Effect on sigma is second order (tol^2).
    const double tol = (double)(rows > 64 ? rows : 64) * 2.220446049250313e-16;
    int sweeps = 0;
    double prev_mu = 0.0;
    const int max_sweeps = 40;
    bool converged = (n == 1);
    while (!converged && sweeps < max_sweeps) {
        MOR_TRY(small_info_reset(d_info));
        const int big = rows > n ? rows : n;
        // The Gram form of the round (one block-wide reduction instead of four) updates the pair block's Gram matrix
        // algebraically inside a round.  That is accurate when the factor is a column / row scaling of a
        // well-conditioned matrix (the one-pass path, certified by the scaled Gram test) and NOT when it is a strongly
        // graded multi-pass factor: on the 256 x 256 core of a Burgers POD it keeps rotating on rounding noise and never
        // meets the threshold.  So: only where the caller vouches for the scaling, and only for the first sweeps -- if
        // it has not converged by then, the four-step form (dot products recomputed from the columns before every
        // rotation) takes over from wherever the orthogonal transformations so far have left the columns.
        static const bool gram_env = !(getenv("MOR_B200_JACOBI_GRAM") && getenv("MOR_B200_JACOBI_GRAM")[0] == '0');
        const bool gram_form = well_scaled && gram_env && sweeps < 6;
        if (V == nullptr && rows <= 4 * ST && n >= 16)
            MOR_TRY(launch_block_sweep(A, lda, rows, n, tol, d_info, ver_buf.as<int>(), sweeps, gram_form));
        else if (big <= 4 * ST) MOR_TRY(launch_sweep<4>(A, lda, rows, V, ldv, n, n, n_even, tol, d_info));
        else if (big <= 16 * ST) MOR_TRY(launch_sweep<16>(A, lda, rows, V, ldv, n, n, n_even, tol, d_info));
        else MOR_TRY(launch_sweep<0>(A, lda, rows, V, ldv, n, n, n_even, tol, d_info));
        SmallInfo h;
        MOR_TRY(small_info_read(d_info, &h));
        if (h.sync_fail) {
            set_error("Jacobi SVD: a block-version wait of the barrier-free sweep timed out");
            return MOR_ERR_CUDA;
        }
        ++sweeps;
        if (getenv("MOR_B200_DEBUG"))
            fprintf(stderr, "[mor_b200]     jacobi sweep %d: %d rotations, max cosine %.3e\n", sweeps, h.rotations, h.max_cos);
        // Quadratic convergence: if the largest cosine met in this sweep was mu, what is left after
        // it is C * mu^2.  C is estimated from the previous sweep (mu / mu_prev^2, at least 1) and a
        // 10x margin is applied; once that bound is below the rotation threshold the sweep that
        // would merely verify "no rotations" is skipped.
        const double mu = h.max_cos;
        bool predicted = false;
        if (sweeps >= 2 && mu < 1e-6 && prev_mu > 0.0 && prev_mu < 1e-2) {
            const double cq = std::max(1.0, mu / (prev_mu * prev_mu));
            predicted = 10.0 * cq * mu * mu < tol;
        }
        prev_mu = mu;
        converged = (h.rotations == 0) || predicted;
    }
    if (sweeps_out) *sweeps_out = sweeps;
    if (!converged) {
        set_error("Jacobi SVD of the " + std::to_string(rows) + " x " + std::to_string(n) +
                  " core did not converge in " + std::to_string(max_sweeps) + " sweeps");
        return MOR_ERR_NOCONV;
    }
    colnorm_kernel<<<n, JT, 0, cx.stream>>>(A, lda, rows, norms.as<double>());
    count_launch();
    MOR_CUDA(cudaGetLastError());
    std::vector<double> s((size_t)n);
    MOR_CUDA(cudaMemcpyAsync(s.data(), norms.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, cx.stream));
    MOR_CUDA(cudaStreamSynchronize(cx.stream));
    order.resize((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        const bool nx = std::isnan(s[x]), ny = std::isnan(s[y]);
        if (nx || ny) return nx && !ny;
        return s[x] > s[y];
    });
    sigma.resize((size_t)n);
    for (int j = 0; j < n; ++j) sigma[j] = s[order[j]];
    return MOR_OK;
}

}  // namespace mor
