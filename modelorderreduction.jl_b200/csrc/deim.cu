// DEIM interpolation-index selection -- replaces the body of deim_interpolation_indices,
// /root/reference/src/deim.jl:9-28.
//
// Per greedy step l (reference lines in brackets):
//   deim_step_kernel    r_i = |u_l[i] - sum_j U[i,j] c_j| fused with the abs-argmax   [l.22-24]
//                       -- one streaming pass over columns 1..l of the basis, HBM-bound.
//   deim_local_kernel   reduce the per-CTA candidates, gather the winner's basis row [l.16-20]
//   (ncclAllGather)     (value, global index, row) of every rank                  (multi-GPU)
//   deim_update_kernel  pick the global winner (value, then LOWEST global index = Julia's
//                       first-index tie-break, NaN maximal), append its row to P'U and
//                       solve (P'U) c = P'u_{l+1} for the next step                    [l.21]
//
// The small solve keeps the LU factors of P'U incrementally: P'U grows by one row and one
// column per step and the greedy choice IS partial pivoting (the pivot is the residual of
// largest magnitude), so the bordered Doolittle update reproduces dgetrf's factorisation
// without row exchanges.  Z = inv(U-factor) is kept explicitly so every phase is a parallel
// dot product instead of a 255-long dependent substitution chain.
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "internal.h"

namespace mor {

struct Cand {
    double val;
    long long idx;
};

#define MOR_IDX_NONE LLONG_MAX

// true when candidate a beats candidate b (Julia argmax semantics + lowest-index tie-break)
__device__ __forceinline__ bool cand_better(double va, long long ia, double vb, long long ib) {
    if (ia == MOR_IDX_NONE) return false;
    if (ib == MOR_IDX_NONE) return true;
    const bool na = va != va, nb = vb != vb;
    if (na | nb) return (na & nb) ? (ia < ib) : na;
    return (va > vb) || (va == vb && ia < ib);
}

__device__ __forceinline__ void cand_warp_reduce(double& v, long long& i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, off);
        long long oi = __shfl_xor_sync(0xffffffffu, i, off);
        if (cand_better(ov, oi, v, i)) {
            v = ov;
            i = oi;
        }
    }
}

// block-wide reduce; result valid in thread 0
template <int NWARPS>
__device__ __forceinline__ void cand_block_reduce(double& v, long long& i, double* sv, long long* si) {
    cand_warp_reduce(v, i);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sv[warp] = v;
        si[warp] = i;
    }
    __syncthreads();
    if (warp == 0) {
        v = lane < NWARPS ? sv[lane] : -1.0;
        i = lane < NWARPS ? si[lane] : MOR_IDX_NONE;
        cand_warp_reduce(v, i);
    }
}

__device__ __forceinline__ double2 ld_stream2(const double* p) {
    double2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream1(const double* p) {
    double v;
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// running best of one thread; rows arrive in ascending order so ties keep the earlier row
__device__ __forceinline__ void cand_take(double& bv, long long& bi, double v, long long idx) {
    if (!(bv != bv) && ((v != v) || v > bv)) {
        bv = v;
        bi = idx;
    }
}

constexpr int DEIM_THREADS = 256;
constexpr int DEIM_BLOCK = 16;   // panel width of the blocked path (= the 16-column stages of deim_panel_kernel)
constexpr int DEIM_SUB = 4;      // sub-panel width inside a panel (deim_step4_kernel)

// Fused residual + abs-argmax over this rank's rows.  l = number of previous columns.
// VEC: 16-byte loads (needs 16B-aligned base and even ld).  Each CTA walks chunks of
// 2*R*256 consecutive rows; for every column the CTA reads one contiguous 4*R KB run.
template <int R, bool VEC>
__global__ void __launch_bounds__(DEIM_THREADS)
deim_step_kernel(const double* __restrict__ U, int64_t m, int64_t ld, int l,
                 const double* __restrict__ cg, int64_t row_offset, Cand* __restrict__ out) {
    __shared__ double sc[1024];
    __shared__ double sv[DEIM_THREADS / 32];
    __shared__ long long si[DEIM_THREADS / 32];
    const int tid = threadIdx.x;
    for (int j = tid; j < l; j += DEIM_THREADS) sc[j] = cg[j];
    __syncthreads();

    const int64_t CH = 2 * R * DEIM_THREADS;
    const int64_t nchunks = (m + CH - 1) / CH;
    const double* __restrict__ ul = U + (int64_t)l * ld;
    double bv = -1.0;
    long long bi = MOR_IDX_NONE;

    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int64_t c0 = chunk * CH;
        if (c0 + CH <= m) {
            if (VEC) {
                double2 acc[R];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = make_double2(0.0, 0.0);
                const double* p = U + c0 + 2 * tid;
                int j = 0;
                // two columns per trip: 2*R independent 16-byte loads in flight per thread
                for (; j + 2 <= l; j += 2) {
                    double2 v0[R], v1[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) v0[r] = ld_stream2(p + r * (2 * DEIM_THREADS));
#pragma unroll
                    for (int r = 0; r < R; ++r) v1[r] = ld_stream2(p + ld + r * (2 * DEIM_THREADS));
                    const double ca = sc[j], cb = sc[j + 1];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        acc[r].x = fma(v0[r].x, ca, acc[r].x);
                        acc[r].y = fma(v0[r].y, ca, acc[r].y);
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        acc[r].x = fma(v1[r].x, cb, acc[r].x);
                        acc[r].y = fma(v1[r].y, cb, acc[r].y);
                    }
                    p += 2 * ld;
                }
                if (j < l) {
                    const double cj = sc[j];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const double2 v = ld_stream2(p + r * (2 * DEIM_THREADS));
                        acc[r].x = fma(v.x, cj, acc[r].x);
                        acc[r].y = fma(v.y, cj, acc[r].y);
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int64_t row = c0 + 2 * tid + r * (2 * DEIM_THREADS);
                    const double2 u = ld_stream2(ul + row);
                    cand_take(bv, bi, fabs(u.x - acc[r].x), row_offset + row);
                    cand_take(bv, bi, fabs(u.y - acc[r].y), row_offset + row + 1);
                }
            } else {
                double acc[2 * R];
#pragma unroll
                for (int r = 0; r < 2 * R; ++r) acc[r] = 0.0;
                const double* p = U + c0 + tid;
#pragma unroll 2
                for (int j = 0; j < l; ++j) {
                    const double cj = sc[j];
#pragma unroll
                    for (int r = 0; r < 2 * R; ++r)
                        acc[r] = fma(ld_stream1(p + r * DEIM_THREADS), cj, acc[r]);
                    p += ld;
                }
#pragma unroll
                for (int r = 0; r < 2 * R; ++r) {
                    const int64_t row = c0 + tid + r * DEIM_THREADS;
                    cand_take(bv, bi, fabs(ld_stream1(ul + row) - acc[r]), row_offset + row);
                }
            }
        } else {  // ragged last chunk
            for (int64_t row = c0 + tid; row < m; row += DEIM_THREADS) {
                double acc = 0.0;
                const double* p = U + row;
                for (int j = 0; j < l; ++j) acc = fma(p[(int64_t)j * ld], sc[j], acc);
                cand_take(bv, bi, fabs(ul[row] - acc), row_offset + row);
            }
        }
    }
    cand_block_reduce<DEIM_THREADS / 32>(bv, bi, sv, si);
    if (tid == 0) {
        out[blockIdx.x].val = bv;
        out[blockIdx.x].idx = bi;
    }
}

// Second blocking level: residuals of the next sw <= 4 columns of a panel P1 with respect to the s0 <= 12
// interpolation points already chosen inside that panel,
//     R2[:, q] = P1[:, s0 + q] - sum_{t < s0} P1[:, t] * D[t][q],
// written to the sub-panel R2 and fused with the abs-argmax of its first column (= the next greedy step).
// One streaming pass over s0 + sw panel columns instead of one pass per greedy step.  D: 16 x 4 column-major.
template <int R>
__global__ void __launch_bounds__(DEIM_THREADS)
deim_step4_kernel(const double* __restrict__ P1, int64_t m, int64_t ld1, int s0, int sw, const double* __restrict__ Dg,
                  double* __restrict__ R2, int64_t ld2, int64_t row_offset, Cand* __restrict__ out) {
    __shared__ double sd[DEIM_BLOCK * DEIM_SUB];
    __shared__ double sv[DEIM_THREADS / 32];
    __shared__ long long si[DEIM_THREADS / 32];
    const int tid = threadIdx.x;
    if (tid < DEIM_BLOCK * DEIM_SUB) sd[tid] = Dg[tid];
    __syncthreads();

    const int64_t CH = 2 * R * DEIM_THREADS;
    const int64_t nchunks = (m + CH - 1) / CH;
    double bv = -1.0;
    long long bi = MOR_IDX_NONE;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int64_t c0 = chunk * CH;
        if (c0 + CH <= m) {
            double2 acc[R][DEIM_SUB];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int q = 0; q < DEIM_SUB; ++q) acc[r][q] = make_double2(0.0, 0.0);
            const double* p = P1 + c0 + 2 * tid;
#pragma unroll 2
            for (int t = 0; t < s0; ++t) {
                double2 v[R];
#pragma unroll
                for (int r = 0; r < R; ++r) v[r] = ld_stream2(p + r * (2 * DEIM_THREADS));
#pragma unroll
                for (int q = 0; q < DEIM_SUB; ++q) {
                    const double d = sd[q * DEIM_BLOCK + t];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        acc[r][q].x = fma(v[r].x, d, acc[r][q].x);
                        acc[r][q].y = fma(v[r].y, d, acc[r][q].y);
                    }
                }
                p += ld1;
            }
#pragma unroll
            for (int q = 0; q < DEIM_SUB; ++q) {
                if (q < sw) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int64_t row = c0 + 2 * tid + r * (2 * DEIM_THREADS);
                        const double2 u = ld_stream2(p + r * (2 * DEIM_THREADS));
                        const double2 res = make_double2(u.x - acc[r][q].x, u.y - acc[r][q].y);
                        *reinterpret_cast<double2*>(R2 + (int64_t)q * ld2 + row) = res;
                        if (q == 0) {
                            cand_take(bv, bi, fabs(res.x), row_offset + row);
                            cand_take(bv, bi, fabs(res.y), row_offset + row + 1);
                        }
                    }
                }
                p += ld1;
            }
        } else {  // ragged last chunk
            for (int64_t row = c0 + tid; row < m; row += DEIM_THREADS) {
                double acc[DEIM_SUB] = {0.0, 0.0, 0.0, 0.0};
                const double* p = P1 + row;
                for (int t = 0; t < s0; ++t) {
                    const double v = p[(int64_t)t * ld1];
#pragma unroll
                    for (int q = 0; q < DEIM_SUB; ++q) acc[q] = fma(v, sd[q * DEIM_BLOCK + t], acc[q]);
                }
#pragma unroll
                for (int q = 0; q < DEIM_SUB; ++q) {
                    if (q < sw) {
                        const double res = p[(int64_t)(s0 + q) * ld1] - acc[q];
                        R2[(int64_t)q * ld2 + row] = res;
                        if (q == 0) cand_take(bv, bi, fabs(res), row_offset + row);
                    }
                }
            }
        }
    }
    cand_block_reduce<DEIM_THREADS / 32>(bv, bi, sv, si);
    if (tid == 0) {
        out[blockIdx.x].val = bv;
        out[blockIdx.x].idx = bi;
    }
}

// D (16 x 4 column-major, zero padded) = Zb[0:s0, 0:s0] * Ufb[0:s0, s0 : s0+sw] from the panel's own factors
// (row-major with stride kb): column q solves (P'P1_s0) d = P'P1[:, s0+q].
__global__ void __launch_bounds__(DEIM_BLOCK * DEIM_SUB)
deim_sub_coef_kernel(const double* __restrict__ Zb, const double* __restrict__ Ufb, int kb, int s0, int sw,
                     double* __restrict__ D) {
    const int t = threadIdx.x % DEIM_BLOCK, q = threadIdx.x / DEIM_BLOCK;
    double s = 0.0;
    if (t < s0 && q < sw)
        for (int r = t; r < s0; ++r) s = fma(Zb[t * kb + r], Ufb[r * kb + s0 + q], s);
    D[q * DEIM_BLOCK + t] = s;
}

// Reduce the per-CTA candidates of this rank and gather the winner's basis row:
// send = [value, bits(global index), U[idx, 0..k)]
// (blocked path: followed by the winner's row of the current residual panel, Pn[idx, 0..bw), padded to 16)
__global__ void __launch_bounds__(1024)
deim_local_kernel(const Cand* __restrict__ cands, int ncand, const double* __restrict__ U, int64_t ld,
                  int64_t row_offset, int k, double* __restrict__ send, const double* __restrict__ Pn,
                  int64_t ldp, int bw, const double* __restrict__ P2, int64_t ld2, int sw, int kcols) {
    __shared__ double sv[32];
    __shared__ long long si[32];
    __shared__ long long widx;
    double bv = -1.0;
    long long bi = MOR_IDX_NONE;
    for (int t = threadIdx.x; t < ncand; t += 1024) {
        const double v = cands[t].val;
        const long long i = cands[t].idx;
        if (cand_better(v, i, bv, bi)) {
            bv = v;
            bi = i;
        }
    }
    cand_block_reduce<32>(bv, bi, sv, si);
    if (threadIdx.x == 0) {
        widx = bi;
        send[0] = bv;
        send[1] = __longlong_as_double(bi);
    }
    __syncthreads();
    const long long g = widx;
    for (int j = threadIdx.x; j < k; j += 1024)
        send[2 + j] = (g == MOR_IDX_NONE || j >= kcols) ? 0.0 : U[(g - row_offset) + (int64_t)j * ld];   // later columns: deferred
    if (Pn != nullptr && threadIdx.x < DEIM_BLOCK)
        send[2 + k + threadIdx.x] =
            (g == MOR_IDX_NONE || (int)threadIdx.x >= bw) ? 0.0 : Pn[(g - row_offset) + (int64_t)threadIdx.x * ldp];
    if (P2 != nullptr && threadIdx.x >= 32 && threadIdx.x < 32 + DEIM_SUB) {
        const int q = threadIdx.x - 32;
        send[2 + k + DEIM_BLOCK + q] = (g == MOR_IDX_NONE || q >= sw) ? 0.0 : P2[(g - row_offset) + (int64_t)q * ld2];
    }
}

__device__ __forceinline__ double warp_sum(double s) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    return s;
}

// Replicated on every rank: choose the global winner, append its row to W = P'U (row i),
// extend the LU factors (L row i, U-factor row i, Z column i) and compute the coefficient
// vector c (length i+1) of the NEXT step.  All k x k arrays are row-major with stride k; Zt is
// the transpose of Z, kept so that every phase reads with the output index contiguous.
// Every phase is a (triangular) matrix-vector product: thread (x, y) accumulates the terms
// t = y (mod ways) of output x, the `ways` partial sums are combined through shared memory --
// a sequential depth of k / ways instead of k on this latency-critical single-CTA kernel.
struct UpdState {
    double *Wm, *Uf, *Lm, *Z, *Zt, *c;   // k x k row-major (stride k) factors and the next coefficient vector
    long long* idx_out;                  // selected indices (1-based) or nullptr
    int i, k;                            // step (row being appended) and dimension of this state
    int rowoff;                          // offset of this state's row inside a gathered record
    int last;                            // 1: the final greedy step (a zero pivot no longer matters)
    int jmax;                            // U-factor row i is formed for columns [i, jmax) (k unless later columns are deferred)
};
struct UpdArgs {
    UpdState st[2];                      // blockIdx.x selects: 0 = P'U of the basis, 1 = P'R of the current panel
};

__global__ void __launch_bounds__(1024)
deim_update_kernel(const double* __restrict__ gathered, int nranks, int stride, UpdArgs args, int* __restrict__ status) {
    const UpdState& S = args.st[blockIdx.x];
    const int i = S.i, k = S.k;
    double* __restrict__ Wm = S.Wm;
    double* __restrict__ Uf = S.Uf;
    double* __restrict__ Lm = S.Lm;
    double* __restrict__ Z = S.Z;
    double* __restrict__ Zt = S.Zt;
    double* __restrict__ c = S.c;
    long long* __restrict__ idx_out = S.idx_out;
    __shared__ double w[1024], ell[1024], ucol[1024], y[1024], part[1024];
    __shared__ int win_s;
    const int tid = threadIdx.x;
    const int KP = (k + 31) & ~31;       // outputs per phase, padded to whole warps
    const int ways = 1024 / KP;          // reduction split (k <= 1024 => ways >= 1)
    const int x = tid % KP, yy = tid / KP;
    const bool active = yy < ways;
    if (tid == 0) {
        int win = 0;
        double bv = -1.0;
        long long bi = MOR_IDX_NONE;
        for (int r = 0; r < nranks; ++r) {
            const double v = gathered[(size_t)r * stride];
            const long long ix = __double_as_longlong(gathered[(size_t)r * stride + 1]);
            if (cand_better(v, ix, bv, bi)) {
                bv = v;
                bi = ix;
                win = r;
            }
        }
        win_s = win;
        if (idx_out != nullptr) idx_out[i] = bi + 1;  // 1-based, like Julia
    }
    __syncthreads();
    const double* src = gathered + (size_t)win_s * stride + S.rowoff;
    for (int j = tid; j < k; j += 1024) {
        const double v = src[j];
        w[j] = v;
        Wm[(size_t)i * k + j] = v;
    }
    __syncthreads();
    // combine the `ways` partial sums of output x (valid in threads with yy == 0)
    auto combine = [&](double s) -> double {
        part[tid] = s;
        __syncthreads();
        double tot = 0.0;
        if (yy == 0)
            for (int q = 0; q < ways; ++q) tot += part[q * KP + x];
        __syncthreads();
        return tot;
    };
    // L row: ell[j] = sum_{t <= j} w[t] * Z[t][j],  j < i   (Z = inv(Uf[0:i,0:i]), upper triangular)
    {
        double s = 0.0;
        if (active && x < i) {
#pragma unroll 8
            for (int t = yy; t <= x; t += ways) s = fma(w[t], __ldcg(&Z[(size_t)t * k + x]), s);
        }
        s = combine(s);
        if (yy == 0 && x < i) {
            ell[x] = s;
            Lm[(size_t)i * k + x] = s;
        }
    }
    __syncthreads();
    // U-factor row i:  Uf[i][j] = w[j] - sum_{t<i} ell[t] * Uf[t][j],  j >= i
    {
        const int j = i + x;
        double s = 0.0;
        if (active && j < S.jmax) {
#pragma unroll 8
            for (int t = yy; t < i; t += ways) s = fma(ell[t], __ldcg(&Uf[(size_t)t * k + j]), s);
        }
        s = combine(s);
        if (yy == 0 && j < S.jmax) Uf[(size_t)i * k + j] = w[j] - s;
    }
    __syncthreads();
    for (int t = tid; t < i; t += 1024) ucol[t] = Uf[(size_t)t * k + i];
    __syncthreads();
    const double pivot = Uf[(size_t)i * k + i];
    if (tid == 0 && pivot == 0.0 && !S.last && idx_out != nullptr && *status == 0) *status = i + 1;
    // Z column i:  Z[r][i] = -(sum_{t=r}^{i-1} Z[r][t] * Uf[t][i]) / pivot,  r < i;  Z[i][i] = 1/pivot
    {
        double s = 0.0;
        if (active && x < i) {
#pragma unroll 8
            for (int t = x + yy; t < i; t += ways) s = fma(__ldcg(&Zt[(size_t)t * k + x]), ucol[t], s);
        }
        s = combine(s);
        if (yy == 0 && x <= i) {
            const double z = (x == i) ? 1.0 / pivot : -s / pivot;
            Z[(size_t)x * k + i] = z;
            Zt[(size_t)i * k + x] = z;
        }
    }
    if (i + 1 >= S.jmax) return;
    __syncthreads();
    for (int t = tid; t <= i; t += 1024) y[t] = Uf[(size_t)t * k + i + 1];
    __syncthreads();
    // c = Z[0:i+1, 0:i+1] * y   (y = inv(L) * P'u_{l+1} is column i+1 of the U-factor)
    {
        double s = 0.0;
        if (active && x <= i) {
#pragma unroll 8
            for (int t = x + yy; t <= i; t += ways) s = fma(Zt[(size_t)t * k + x], y[t], s);   // row i was written above
        }
        s = combine(s);
        if (yy == 0 && x <= i) c[x] = s;
    }
}

// Deferred columns (pipelined host ingest: the basis arrives 16 columns at a time, so the rows of the points
// chosen so far cannot be gathered beyond the columns that have landed).  At the panel boundary l1:
//   deim_gather_block_kernel   W12[t][q] = U[idx[t], l1 + q], t < l1 (zero for rows of other ranks; all-reduced)
//   deim_fwd_subst_kernel      Uf[0:l1, l1 + q] = inv(L[0:l1, 0:l1]) * W12[:, q]  (unit lower L, one warp per column q)
// which is what the per-step bordered update would have left in those columns of the U-factor.
__global__ void __launch_bounds__(128)
deim_gather_block_kernel(const double* __restrict__ U, int64_t m, int64_t ld, int64_t row_offset,
                         const long long* __restrict__ idx, int l1, int bw, double* __restrict__ W12) {
    const int t = blockIdx.x * 8 + threadIdx.x / DEIM_BLOCK, q = threadIdx.x % DEIM_BLOCK;
    if (t >= l1) return;
    const long long loc = idx[t] - 1 - row_offset;
    W12[t * DEIM_BLOCK + q] = (loc >= 0 && loc < m && q < bw) ? U[loc + (int64_t)(l1 + q) * ld] : 0.0;
}

__global__ void __launch_bounds__(32 * DEIM_BLOCK)
deim_fwd_subst_kernel(const double* __restrict__ Lm, const double* __restrict__ W12, int k, int l1, int bw,
                      double* Uf) {
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (q >= bw) return;
    double* y = Uf + l1 + q;                      // y[t * k] = Uf[t][l1 + q]
    for (int t = 0; t < l1; ++t) {
        double s = 0.0;
        for (int r = lane; r < t; r += 32) s = fma(__ldcg(&Lm[(size_t)t * k + r]), __ldcg(&y[(size_t)r * k]), s);
        s = warp_sum(s);
        if (lane == 0) y[(size_t)t * k] = W12[t * DEIM_BLOCK + q] - s;
        __syncwarp();
    }
}

// Coefficients of the next panel, packed for deim_panel_kernel: C = Z[0:l1, 0:l1] * Uf[0:l1, l1 : l1+bw]
// (Uf[:, j] = inv(L) P'u_j is already part of the U-factor, Z = inv(Uf[0:l1, 0:l1])), i.e. column q of C
// solves (P'U_l1) c = P'u_{l1+q}, deim.jl:21.  One CTA per 16-row slab: Cp[c][q][kk] = C[16 c + kk][q].
__global__ void __launch_bounds__(DEIM_BLOCK * 20)
deim_block_coef_kernel(const double* __restrict__ Z, const double* __restrict__ Uf, int k, int l1, int bw,
                       double* __restrict__ Cp) {
    const int e = threadIdx.x, q = e / 20, kk = e % 20;
    const int t = blockIdx.x * DEIM_BLOCK + kk;
    double s0 = 0.0, s1 = 0.0;
    if (kk < DEIM_BLOCK && q < bw && t < l1) {
        const double* z = Z + (size_t)t * k;
        const double* u = Uf + l1 + q;
        int r = t;
        for (; r + 2 <= l1; r += 2) {
            s0 = fma(z[r], u[(size_t)r * k], s0);
            s1 = fma(z[r + 1], u[(size_t)(r + 1) * k], s1);
        }
        if (r < l1) s0 = fma(z[r], u[(size_t)r * k], s0);
    }
    Cp[(size_t)blockIdx.x * (DEIM_BLOCK * 20) + e] = s0 + s1;
}

// MOR_B200_DEIM_BLOCK=0 selects the streaming path (one pass over columns 1..l per greedy step, the
// reference's own access pattern); default is the blocked path whenever the basis layout allows it.
static bool deim_blocked_enabled() {
    const char* e = getenv("MOR_B200_DEIM_BLOCK");
    return !(e && e[0] == '0');
}

int32_t deim_indices_device(const double* B, int64_t m, int64_t k, int64_t ld, int64_t* idx_host,
                            const cudaEvent_t* col_ready, int group_cols) {
    Ctx& cx = ctx();
    MOR_ARG(B != nullptr && idx_host != nullptr, "deim_interpolation_indices: null pointer");
    MOR_ARG(k >= 1 && k <= 1024, "deim_interpolation_indices: number of basis columns must be in 1..1024");
    MOR_ARG(m >= 0 && ld >= (m > 0 ? m : 1), "deim_interpolation_indices: bad m / ld");
    for (double& t : cx.timings) t = 0.0;

    std::vector<int64_t> all_m;
    MOR_TRY(comm_allgather_i64(m, all_m));
    int64_t row_offset = 0, m_global = 0;
    for (int r = 0; r < cx.nranks; ++r) {
        if (r < cx.rank) row_offset += all_m[r];
        m_global += all_m[r];
    }
    MOR_ARG(m_global >= 1, "deim_interpolation_indices: empty basis");

    constexpr int R = 4;
    const int64_t CH = 2 * R * DEIM_THREADS;
    int64_t nchunks = (m + CH - 1) / CH;
    int grid = (int)(nchunks < (int64_t)cx.num_sms * 8 ? nchunks : (int64_t)cx.num_sms * 8);
    if (grid < 1) grid = 1;
    const bool vec = ((uintptr_t)B % 16 == 0) && (ld % 2 == 0);

    // Blocked path (k > 16): the residuals of 16 columns at a time with respect to all earlier points come
    // from ONE pass over the earlier columns (deim_panel_kernel); the greedy steps inside a block run the
    // same streaming kernel on that 16-column residual panel -- about 8 m k (k / 32 + 18) bytes of traffic
    // instead of 4 m k (k + 1).  It needs a 16-byte aligned basis with an even leading dimension (bulk
    // copies) and 128 m bytes of workspace; every rank must take the same path (the gathered records differ).
    const int64_t ldp = (m + 1) & ~int64_t(1);
    bool blocked = deim_blocked_enabled() && k > DEIM_BLOCK && (m == 0 || (vec && ld >= ldp));
    DevBuf panel;
    if (blocked && m > 0 && panel.alloc(sizeof(double) * (size_t)ldp * (DEIM_BLOCK + DEIM_SUB)) != MOR_OK) {
        cudaGetLastError();   // no room for the panel: the streaming path needs no workspace
        blocked = false;
    }
    if (cx.nranks > 1) {   // all ranks or none
        std::vector<int64_t> votes;
        MOR_TRY(comm_allgather_i64(blocked ? 1 : 0, votes));
        for (int64_t v : votes) blocked = blocked && v != 0;
        if (!blocked) panel.release();
    }

    const size_t kk = (size_t)k * k;
    // gathered record: value, index, basis row, panel row (16), sub-panel row (4)
    const int rec = (int)k + 2 + (blocked ? DEIM_BLOCK + DEIM_SUB : 0);
    DevBuf state, cands, xfer;
    // state: Wm, Uf, Lm, Z, Zt (k*k each), c (k), idx (k int64), status; block state: 5 * 16*16 + 16; packed C
    // block state 5 * 16*16 + 16, sub-block state 5 * 4*4 + 4, sub-panel coefficients D 16 x 4
    const size_t bdoubles = 5 * DEIM_BLOCK * DEIM_BLOCK + DEIM_BLOCK + 5 * DEIM_SUB * DEIM_SUB + DEIM_SUB +
                            DEIM_BLOCK * DEIM_SUB;
    const size_t cp_doubles = (size_t)((k + DEIM_BLOCK - 1) / DEIM_BLOCK) * DEIM_BLOCK * 20 + (size_t)k * DEIM_BLOCK;
    MOR_TRY(state.alloc(sizeof(double) * (5 * kk + k + bdoubles + cp_doubles) + sizeof(long long) * k + 16));
    MOR_TRY(cands.alloc(sizeof(Cand) * (size_t)grid));
    // one record slot per greedy step: the update of the k x k factors (blocked path: on the side stream) may
    // still be reading step i's record while step i+1's is being written
    const size_t slot = (size_t)rec * (size_t)(cx.nranks + 1);
    MOR_TRY(xfer.alloc(sizeof(double) * slot * (size_t)(blocked ? k : 1)));
    MOR_CUDA(cudaMemsetAsync(state.p, 0, state.bytes, cx.stream));
    double* Wm = state.as<double>();
    double* Uf = Wm + kk;
    double* Lm = Uf + kk;
    double* Z = Lm + kk;
    double* Zt = Z + kk;
    double* c = Zt + kk;
    double* bst = c + k;                  // block state
    double* Cp = bst + bdoubles;          // packed panel coefficients
    long long* idx_d = reinterpret_cast<long long*>(Cp + cp_doubles);
    int* status_d = reinterpret_cast<int*>(idx_d + k);
    const size_t bb = (size_t)DEIM_BLOCK * DEIM_BLOCK, ss = (size_t)DEIM_SUB * DEIM_SUB;
    double* sst = bst + 5 * bb + DEIM_BLOCK;        // sub-block state
    double* Dsub = sst + 5 * ss + DEIM_SUB;         // sub-panel coefficients
    double* R2 = blocked && m > 0 ? panel.as<double>() + (size_t)ldp * DEIM_BLOCK : nullptr;
    double* W12 = Cp + (cp_doubles - (size_t)k * DEIM_BLOCK);   // deferred columns of the chosen rows (k x 16)
    // pipelined ingest: later columns of the basis are not there yet, so the k x k update is confined to the
    // columns of the current panel and the rest is caught up at each panel boundary
    const bool defer = blocked && col_ready != nullptr;
    if (!blocked && col_ready != nullptr)
        for (int g = 0; g * group_cols < (int)k; ++g) MOR_CUDA(cudaStreamWaitEvent(cx.stream, col_ready[g], 0));
    // Blocked path: inside a block only the 16 x 16 factors of the panel decide the next coefficients, so the
    // (much longer) bordered update of the k x k factors runs on a side stream, off the per-step latency chain;
    // the main stream joins it at the next block boundary, where the panel coefficients need Z and Uf.
    cudaEvent_t ev_rec = nullptr, ev_side = nullptr;
    if (blocked) {
        if (!cx.copy_stream) MOR_CUDA(cudaStreamCreateWithFlags(&cx.copy_stream, cudaStreamNonBlocking));
        MOR_CUDA(cudaEventCreateWithFlags(&ev_rec, cudaEventDisableTiming));
        MOR_CUDA(cudaEventCreateWithFlags(&ev_side, cudaEventDisableTiming));
        MOR_CUDA(cudaEventRecord(ev_rec, cx.stream));          // the memset above
        MOR_CUDA(cudaStreamWaitEvent(cx.copy_stream, ev_rec, 0));
    }
    cudaStream_t side = blocked ? cx.copy_stream : cx.stream;

    Timer t_total(MOR_T_TOTAL), t_step(MOR_T_DEIM_STEP), t_tail(MOR_T_DEIM_TAIL), t_panel(MOR_T_DEIM_PANEL);
    t_total.start();
    for (int i = 0; i < (int)k; ++i) {
        const int l0 = blocked ? (i / DEIM_BLOCK) * DEIM_BLOCK : 0;   // first column of the current block
        const int j = i - l0;                                          // step inside the block
        const int bw = blocked ? (int)(k - l0 < DEIM_BLOCK ? k - l0 : DEIM_BLOCK) : (int)k;
        // the matrix the streaming kernel works on: the basis itself, or the residual panel of this block
        const double* Pn = (blocked && l0 > 0) ? panel.as<double>() : B + (size_t)l0 * ld;
        const int64_t ldn = (blocked && l0 > 0) ? ldp : ld;
        // second level: sub-panels of 4 columns inside the panel
        const int s0 = blocked ? (j / DEIM_SUB) * DEIM_SUB : 0;       // first panel column of the current sub-panel
        const int jj = j - s0;                                        // step inside the sub-panel
        const int sw = blocked ? (bw - s0 < DEIM_SUB ? bw - s0 : DEIM_SUB) : 0;
        const double* P2 = (blocked && s0 > 0) ? R2 : Pn + (size_t)s0 * ldn;
        const int64_t ld2 = (blocked && s0 > 0) ? ldp : ldn;
        double* send = xfer.as<double>() + (blocked ? slot * (size_t)i : 0);
        double* gathered = cx.nranks > 1 ? send + rec : send;
        // pipelined host ingest: columns [i, i + group_cols) are first touched by this step (the panel of a
        // block reads columns < l0 + 16 and DEIM_BLOCK == group_cols; the streaming step i reads columns <= i)
        if (defer && col_ready != nullptr && i % group_cols == 0)
            MOR_CUDA(cudaStreamWaitEvent(cx.stream, col_ready[i / group_cols], 0));
        if (blocked && j == 0 && l0 > 0) {
            MOR_CUDA(cudaEventRecord(ev_side, side));       // all earlier updates of the k x k factors
            MOR_CUDA(cudaStreamWaitEvent(cx.stream, ev_side, 0));
            t_panel.start();
            if (defer) {
                deim_gather_block_kernel<<<(l0 + 7) / 8, 128, 0, cx.stream>>>(B, m, ld, row_offset, idx_d, l0, bw, W12);
                if (cx.nranks > 1) MOR_TRY(comm_allreduce_sum(W12, (size_t)l0 * DEIM_BLOCK));
                deim_fwd_subst_kernel<<<1, 32 * DEIM_BLOCK, 0, cx.stream>>>(Lm, W12, (int)k, l0, bw, Uf);
                count_launch(2);
            }
            deim_block_coef_kernel<<<l0 / DEIM_BLOCK, DEIM_BLOCK * 20, 0, cx.stream>>>(Z, Uf, (int)k, l0, bw, Cp);
            count_launch();
            if (m > 0) MOR_TRY(deim_panel(B, ld, l0, bw, Cp, m, panel.as<double>(), ldp));
            t_panel.stop();
        }
        t_step.start();
        if (blocked && s0 > 0 && jj == 0) {
            // residuals of the next 4 panel columns in one pass over the s0 earlier ones, fused with this step's argmax
            deim_sub_coef_kernel<<<1, DEIM_BLOCK * DEIM_SUB, 0, cx.stream>>>(bst + 3 * bb, bst + bb, bw, s0, sw, Dsub);
            deim_step4_kernel<R><<<grid, DEIM_THREADS, 0, cx.stream>>>(Pn, m, ldn, s0, sw, Dsub, R2, ldp, row_offset,
                                                                       cands.as<Cand>());
            count_launch();
        } else if (blocked) {
            deim_step_kernel<R, true><<<grid, DEIM_THREADS, 0, cx.stream>>>(P2, m, ld2, jj, sst + 5 * ss, row_offset,
                                                                            cands.as<Cand>());
        } else if (vec) {
            deim_step_kernel<R, true><<<grid, DEIM_THREADS, 0, cx.stream>>>(B, m, ld, i, c, row_offset, cands.as<Cand>());
        } else {
            deim_step_kernel<R, false><<<grid, DEIM_THREADS, 0, cx.stream>>>(B, m, ld, i, c, row_offset, cands.as<Cand>());
        }
        t_step.stop();
        t_tail.start();
        deim_local_kernel<<<1, 1024, 0, cx.stream>>>(cands.as<Cand>(), grid, B, ld, row_offset, (int)k, send,
                                                     blocked ? Pn : nullptr, ldn, bw, blocked ? P2 : nullptr, ld2, sw,
                                                     defer ? l0 + bw : (int)k);
        if (cx.nranks > 1) MOR_TRY(comm_allgather(send, gathered, sizeof(double) * (size_t)rec));
        UpdArgs ua;
        ua.st[0] = UpdState{Wm, Uf, Lm, Z, Zt, c, idx_d, i, (int)k, 2, i == (int)k - 1 ? 1 : 0, defer ? l0 + bw : (int)k};
        ua.st[1] = ua.st[0];
        if (blocked) {
            MOR_CUDA(cudaEventRecord(ev_rec, cx.stream));   // record i is complete
            MOR_CUDA(cudaStreamWaitEvent(side, ev_rec, 0));
            deim_update_kernel<<<1, 1024, 0, side>>>(gathered, cx.nranks, rec, ua, status_d);
            UpdArgs ub;   // the panel's 16 x 16 factors and the sub-panel's 4 x 4 factors: two CTAs of one launch
            ub.st[0] = UpdState{bst, bst + bb, bst + 2 * bb, bst + 3 * bb, bst + 4 * bb, bst + 5 * bb, nullptr, j, bw,
                                2 + (int)k, 1, bw};
            ub.st[1] = UpdState{sst, sst + ss, sst + 2 * ss, sst + 3 * ss, sst + 4 * ss, sst + 5 * ss, nullptr, jj, sw,
                                2 + (int)k + DEIM_BLOCK, 1, sw};
            deim_update_kernel<<<2, 1024, 0, cx.stream>>>(gathered, cx.nranks, rec, ub, status_d);
            count_launch(4);
        } else {
            deim_update_kernel<<<1, 1024, 0, cx.stream>>>(gathered, cx.nranks, rec, ua, status_d);
            count_launch(3);
        }
        t_tail.stop();
    }
    if (blocked) {
        MOR_CUDA(cudaEventRecord(ev_side, side));
        MOR_CUDA(cudaStreamWaitEvent(cx.stream, ev_side, 0));
    }
    t_total.stop();
    MOR_CUDA(cudaGetLastError());
    std::vector<long long> idx_h((size_t)k);
    int status_h = 0;
    MOR_CUDA(cudaMemcpyAsync(idx_h.data(), idx_d, sizeof(long long) * (size_t)k, cudaMemcpyDeviceToHost, cx.stream));
    MOR_CUDA(cudaMemcpyAsync(&status_h, status_d, sizeof(int), cudaMemcpyDeviceToHost, cx.stream));
    MOR_CUDA(cudaStreamSynchronize(cx.stream));
    t_total.collect();
    t_step.collect();
    t_tail.collect();
    t_panel.collect();
    if (ev_rec) cudaEventDestroy(ev_rec);
    if (ev_side) cudaEventDestroy(ev_side);
    cx.timings[MOR_T_DEIM_BLOCKED] = blocked ? 1.0 : 0.0;
    for (int64_t i = 0; i < k; ++i) idx_host[i] = (int64_t)idx_h[(size_t)i];
    if (status_h != 0) {
        set_error("deim_interpolation_indices: P'U is exactly singular at step " + std::to_string(status_h + 1) +
                  " (SingularException in the reference)");
        return MOR_ERR_SINGULAR;
    }
    return MOR_OK;
}

// DEIM projector of the reduced nonlinear term, reference src/deim.jl:150:
//     projector = ((@view U[indices, :])' \ (U' * V))'        (pod_dim x deim_dim)
// U'V is a tall-skinny contraction over the row shards (gemm_tn + all-reduce), U[indices, :] is
// gathered from whichever rank owns each selected row, the kd x kd solve is LU with partial
// pivoting (Julia's `\` for a general square matrix), replicated on every rank.
int32_t deim_projector_device(const double* U, int64_t m, int kd, int64_t ldu, const double* V, int kp, int64_t ldv,
                              const int64_t* idx_host, double* P_dev, int64_t ldp) {
    Ctx& cx = ctx();
    MOR_ARG(U && V && idx_host && P_dev, "deim_projector: null pointer");
    MOR_ARG(kd >= 1 && kd <= 1024 && kp >= 1 && kp <= 4096 && ldp >= kp, "deim_projector: bad dimensions");
    std::vector<int64_t> all_m;
    MOR_TRY(comm_allgather_i64(m, all_m));
    int64_t row_offset = 0, m_global = 0;
    for (int r = 0; r < cx.nranks; ++r) {
        if (r < cx.rank) row_offset += all_m[r];
        m_global += all_m[r];
    }
    for (int i = 0; i < kd; ++i)
        MOR_ARG(idx_host[i] >= 1 && idx_host[i] <= m_global, "deim_projector: interpolation index out of range");
    DevBuf UtV, PtU, d_idx, d_fail;
    MOR_TRY(UtV.alloc(sizeof(double) * (size_t)kd * kp));
    MOR_TRY(PtU.alloc(sizeof(double) * (size_t)kd * kd));
    MOR_TRY(d_idx.alloc(sizeof(long long) * (size_t)kd));
    MOR_TRY(d_fail.alloc(sizeof(int)));
    std::vector<long long> idx_ll(idx_host, idx_host + kd);
    MOR_CUDA(cudaMemcpyAsync(d_idx.p, idx_ll.data(), sizeof(long long) * (size_t)kd, cudaMemcpyHostToDevice, cx.stream));
    MOR_TRY(gemm_tn(U, ldu, kd, V, ldv, kp, m, UtV.as<double>(), kd));
    MOR_TRY(gather_rows(U, m, ldu, row_offset, d_idx.as<long long>(), kd, PtU.as<double>(), kd));
    if (cx.nranks > 1) {
        MOR_TRY(comm_allreduce_sum(UtV.as<double>(), (size_t)kd * kp));
        MOR_TRY(comm_allreduce_sum(PtU.as<double>(), (size_t)kd * kd));
    }
    MOR_TRY(small_lu_solve(PtU.as<double>(), kd, kd, /*transpose=*/true, UtV.as<double>(), kp, kd, d_fail.as<int>()));
    MOR_TRY(transpose(UtV.as<double>(), kd, kp, kd, P_dev, ldp));
    int fail = 0;
    MOR_CUDA(cudaMemcpyAsync(&fail, d_fail.p, sizeof(int), cudaMemcpyDeviceToHost, cx.stream));
    MOR_CUDA(cudaStreamSynchronize(cx.stream));
    if (fail != 0) {
        set_error("deim_projector: U[indices, :] is exactly singular (SingularException in the reference)");
        return MOR_ERR_SINGULAR;
    }
    return MOR_OK;
}

}  // namespace mor

using namespace mor;

extern "C" {

int32_t mor_b200_deim_projector_dev(mor_mat_t U, mor_mat_t V, const int64_t* idx, mor_mat_t P_out) {
    MOR_API_LOCK;
    MOR_ARG(U && V && idx && P_out, "mor_b200_deim_projector_dev: null argument");
    MOR_ARG(U->m == V->m && P_out->m == V->n && P_out->n == U->n, "mor_b200_deim_projector_dev: P_out must be pod_dim x deim_dim");
    MOR_TRY(require_ctx());
    return deim_projector_device(U->p, U->m, (int)U->n, U->ld, V->p, (int)V->n, V->ld, idx, P_out->p, P_out->ld);
}

int32_t mor_b200_deim_projector(const double* U, int64_t m, int64_t kd, int64_t ldu, const double* V, int64_t kp,
                                int64_t ldv, const int64_t* idx, double* P_out, int64_t ldp) {
    MOR_API_LOCK;
    MOR_ARG(U && V && idx && P_out, "mor_b200_deim_projector: null pointer");
    MOR_ARG(m >= 0 && kd >= 1 && kp >= 1 && ldu >= (m > 0 ? m : 1) && ldv >= (m > 0 ? m : 1) && ldp >= kp,
            "mor_b200_deim_projector: bad dimensions");
    MOR_TRY(require_ctx());
    mor_mat_t dU = nullptr, dV = nullptr, dP = nullptr;
    int32_t st = mor_b200_mat_alloc(m, kd, &dU);
    if (st == MOR_OK) st = mor_b200_mat_alloc(m, kp, &dV);
    if (st == MOR_OK) st = mor_b200_mat_alloc(kp, kd, &dP);
    if (st == MOR_OK) st = mor_b200_mat_upload(dU, U, ldu);
    if (st == MOR_OK) st = mor_b200_mat_upload(dV, V, ldv);
    if (st == MOR_OK) st = mor_b200_deim_projector_dev(dU, dV, idx, dP);
    if (st == MOR_OK) st = mor_b200_mat_download(dP, P_out, ldp);
    mor_b200_mat_free(dU);
    mor_b200_mat_free(dV);
    mor_b200_mat_free(dP);
    return st;
}

int32_t mor_b200_deim_indices_dev(mor_mat_t B, int64_t* idx_out) {
    MOR_API_LOCK;
    MOR_ARG(B != nullptr, "mor_b200_deim_indices_dev: null matrix");
    MOR_TRY(require_ctx());
    return deim_indices_device(B->p, B->m, B->n, B->ld, idx_out);
}

int32_t mor_b200_deim_indices(const double* B, int64_t m, int64_t k, int64_t ldb, int64_t* idx_out) {
    MOR_API_LOCK;
    MOR_ARG(B != nullptr && idx_out != nullptr, "mor_b200_deim_indices: null pointer");
    MOR_ARG(m >= 0 && k >= 1 && ldb >= (m > 0 ? m : 1), "mor_b200_deim_indices: bad dimensions");
    MOR_TRY(require_ctx());
    Ctx& cx = ctx();
    mor_mat_t d = nullptr;
    MOR_TRY(mor_b200_mat_alloc(m, k, &d));
    // Pipelined ingest: the basis travels in groups of 16 columns on an upload stream; the greedy loop touches
    // columns left to right (a panel needs the columns up to its own), so it starts on group 0 while the rest is
    // still on the PCIe link -- the 34 GB of a 16M x 256 basis hide all but the last panel of the computation.
    constexpr int GROUP = DEIM_BLOCK;
    const int ngroups = (int)((k + GROUP - 1) / GROUP);
    std::vector<cudaEvent_t> ev((size_t)ngroups + 2, nullptr);   // [ngroups] = start of the upload, [ngroups+1] = pad memset done
    int32_t st = MOR_OK;
    auto cu = [&](cudaError_t e, const char* what) {
        if (st == MOR_OK && e != cudaSuccess) st = cuda_fail(e, what, __FILE__, __LINE__);
    };
    if (!cx.aux_stream) cu(cudaStreamCreateWithFlags(&cx.aux_stream, cudaStreamNonBlocking), "cudaStreamCreate(aux)");
    for (auto& e : ev) cu(cudaEventCreate(&e), "cudaEventCreate");
    if (st == MOR_OK) {
        cu(cudaEventRecord(ev[(size_t)ngroups + 1], cx.stream), "cudaEventRecord");   // mat_alloc zeroes the pad row on cx.stream
        cu(cudaStreamWaitEvent(cx.aux_stream, ev[(size_t)ngroups + 1], 0), "cudaStreamWaitEvent");
        cu(cudaEventRecord(ev[(size_t)ngroups], cx.aux_stream), "cudaEventRecord");
        for (int g = 0; g < ngroups && st == MOR_OK; ++g) {
            const int64_t c0 = (int64_t)g * GROUP, nc = k - c0 < GROUP ? k - c0 : GROUP;
            if (m > 0)
                cu(cudaMemcpy2DAsync(d->p + (size_t)c0 * d->ld, (size_t)d->ld * 8, B + (size_t)c0 * ldb, (size_t)ldb * 8,
                                     (size_t)m * 8, (size_t)nc, cudaMemcpyHostToDevice, cx.aux_stream),
                   "cudaMemcpy2DAsync(basis columns)");
            cu(cudaEventRecord(ev[(size_t)g], cx.aux_stream), "cudaEventRecord");
        }
    }
    if (st == MOR_OK) st = deim_indices_device(d->p, d->m, d->n, d->ld, idx_out, ev.data(), GROUP);
    if (cx.aux_stream) cudaStreamSynchronize(cx.aux_stream);     // nothing may still read the caller's buffer
    if (st == MOR_OK) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev[(size_t)ngroups], ev[(size_t)ngroups - 1]) == cudaSuccess) cx.timings[MOR_T_H2D] = ms;
    }
    for (auto& e : ev)
        if (e) cudaEventDestroy(e);
    mor_b200_mat_free(d);
    return st;
}

}  // extern "C"
