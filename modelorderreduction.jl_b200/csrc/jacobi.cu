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
                   std::vector<int>& order, int* sweeps_out) {
    Ctx& cx = ctx();
    MOR_ARG(A && rows >= 1 && n >= 1 && lda >= rows, "jacobi_svd: bad arguments");
    DevBuf info_buf, norms;
    MOR_TRY(info_buf.alloc(sizeof(SmallInfo)));
    MOR_TRY(norms.alloc(sizeof(double) * (size_t)n));
    SmallInfo* d_info = info_buf.as<SmallInfo>();
    if (V) MOR_TRY(small_identity(V, n, ldv));
    const int n_even = (n + 1) & ~1;
    // rotate while |cos(a_p, a_q)| > tol.  dgesvj uses sqrt(rows)*eps; the computed dot product of a
    // tiny and a large column carries rounding noise of about that size, which keeps triggering
    // rotations for tens of sweeps on graded matrices (30 sweeps at 256 x 256, kappa 1e8), so the
    // threshold sits a factor sqrt(rows) above the noise.  Effect on sigma is second order (tol^2).
    const double tol = (double)(rows > 64 ? rows : 64) * 2.220446049250313e-16;
    int sweeps = 0;
    const int max_sweeps = 40;
    bool converged = (n == 1);
    while (!converged && sweeps < max_sweeps) {
        MOR_TRY(small_info_reset(d_info));
        const int big = rows > n ? rows : n;
        if (big <= 4 * ST) MOR_TRY(launch_sweep<4>(A, lda, rows, V, ldv, n, n, n_even, tol, d_info));
        else if (big <= 16 * ST) MOR_TRY(launch_sweep<16>(A, lda, rows, V, ldv, n, n, n_even, tol, d_info));
        else MOR_TRY(launch_sweep<0>(A, lda, rows, V, ldv, n, n, n_even, tol, d_info));
        SmallInfo h;
        MOR_TRY(small_info_read(d_info, &h));
        ++sweeps;
        // quadratic convergence: if the largest cosine met in this sweep was mu, what is left is
        // O(mu^2) -- below the rotation threshold once mu < 3e-9, so the verifying sweep is skipped
        converged = (h.rotations == 0) || (h.max_cos < 3e-9);
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
