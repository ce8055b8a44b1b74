// One-sided (Hestenes) Jacobi SVD of the small replicated core: the n x n triangular factor
// of the snapshot matrix or the (n x l) transposed projection of the randomized SVD.  This is
// the piece of LinearAlgebra.svd / svd!(B) (reference POD.jl:13,27) that is left after the
// tall-skinny reduction (SURVEY.md kernel K6).
//
// A sweep is n-1 rounds of a round-robin tournament; the n/2 column pairs of a round are
// disjoint and rotated concurrently, one CTA per pair: three dot products (a, b, c), the
// rotation that zeroes c, applied to the two columns of A and of the accumulated V.  The
// rotation threshold is |c| <= tol * sqrt(a b) with tol = sqrt(rows) * eps (as in LAPACK's
// dgesvj), which gives the singular values to high relative accuracy.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

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

__global__ void __launch_bounds__(JT)
jacobi_round_kernel(double* __restrict__ A, int lda, int rows, double* __restrict__ V, int ldv, int vrows, int n,
                    int n_even, int round, double tol, SmallInfo* __restrict__ info) {
    __shared__ double red[JT / 32];
    const int b = blockIdx.x, N1 = n_even - 1;
    int p, q;
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
    if (q >= n) return;  // the bye of an odd tournament
    double* ap = A + (size_t)p * lda;
    double* aq = A + (size_t)q * lda;
    double sa = 0.0, sb = 0.0, sc = 0.0;
    for (int r = threadIdx.x; r < rows; r += JT) {
        const double x = ap[r], y = aq[r];
        sa = fma(x, x, sa);
        sb = fma(y, y, sb);
        sc = fma(x, y, sc);
    }
    const double a = block_sum(sa, red), bb = block_sum(sb, red), c = block_sum(sc, red);
    if (!(fabs(c) > tol * sqrt(a) * sqrt(bb))) return;  // also leaves NaN / zero columns alone
    const double zeta = (bb - a) / (2.0 * c);
    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
    for (int r = threadIdx.x; r < rows; r += JT) {
        const double x = ap[r], y = aq[r];
        ap[r] = cs * x - sn * y;
        aq[r] = sn * x + cs * y;
    }
    if (V != nullptr) {
        double* vp = V + (size_t)p * ldv;
        double* vq = V + (size_t)q * ldv;
        for (int r = threadIdx.x; r < vrows; r += JT) {
            const double x = vp[r], y = vq[r];
            vp[r] = cs * x - sn * y;
            vq[r] = sn * x + cs * y;
        }
    }
    if (threadIdx.x == 0) atomicAdd(&info->rotations, 1);
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
    const double tol = sqrt((double)rows) * 2.220446049250313e-16;
    int sweeps = 0;
    const int max_sweeps = 40;
    bool converged = (n == 1);
    while (!converged && sweeps < max_sweeps) {
        MOR_TRY(small_info_reset(d_info));
        for (int round = 0; round < n_even - 1; ++round) {
            jacobi_round_kernel<<<n_even / 2, JT, 0, cx.stream>>>(A, lda, rows, V, ldv, n, n, n_even, round, tol,
                                                                  d_info);
        }
        count_launch(n_even - 1);
        MOR_CUDA(cudaGetLastError());
        SmallInfo h;
        MOR_TRY(small_info_read(d_info, &h));
        ++sweeps;
        converged = (h.rotations == 0);
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
