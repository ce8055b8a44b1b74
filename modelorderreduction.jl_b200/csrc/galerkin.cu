// Galerkin projection of the linear part of the reduced model -- the second "next" row of the hot
// path, reference src/deim.jl:154-155 and 261:
//     A_hat = V' * A * V     (A: SparseMatrixCSC from separate_terms, utils.jl:126)
//     g_hat = V' * g ,   y_hat(0) = V' * snapshot[:, 1]
// A is taken in Julia's own CSC arrays (colptr / rowval 1-based Int64, nzval Float64).  CSC of A
// read column by column is CSR of A', so W = A' V needs no atomics: one thread per (column c of A,
// basis vector j) gathers W[c, j] = sum_p nzval[p] * V[rowval[p], j]; then A_hat = W' V runs on the
// DMMA contraction kernel (gemm_tn).  The projections of vectors are gemm_tn with one column.
// In the single-process multi-GPU mode the COLUMNS of A are sharded (multi.cpp: every GPU takes a column slice of the
// CSC arrays and the rows of V that slice touches -- its row halo comes straight from the caller's host matrix, so no
// inter-GPU exchange is needed -- and the k x k partial products are all-reduced).  With one process per GPU the
// sparse product stays single-rank.  The vector projection is row-sharded like everything else.
#include <algorithm>

#include "internal.h"

namespace mor {
namespace {

// One thread per column c of A and per tile of JT basis vectors: the index / value stream of the column is read once
// per tile (not once per basis vector), the JT gathers V[row, j] are coalesced across the threads of a warp whenever
// neighbouring columns of A touch neighbouring rows (every stencil / MOL operator).  p_base / row_lo shift the global
// 1-based positions of a column SLICE of A (multi-GPU: a rank holds columns [c0, c1) and the rows of V they touch).
constexpr int JT = 8;
__global__ void __launch_bounds__(256)
csc_t_times_dense_kernel(long long ncols, const long long* __restrict__ colptr, const long long* __restrict__ rowval,
                         const double* __restrict__ nzval, long long p_base, long long row_lo, long long nrows,
                         const double* __restrict__ V, long long ldv, int k, double* __restrict__ W, long long ldw,
                         int* __restrict__ bad) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int j0 = blockIdx.y * JT;
    if (c >= ncols) return;
    double s[JT];
#pragma unroll
    for (int u = 0; u < JT; ++u) s[u] = 0.0;
    const double* v = V + (size_t)j0 * ldv;
    const int nj = k - j0 < JT ? k - j0 : JT;
    for (long long p = colptr[c] - 1 - p_base; p < colptr[c + 1] - 1 - p_base; ++p) {
        const long long r = rowval[p] - 1 - row_lo;
        if (r < 0 || r >= nrows) {
            *bad = 1;
            continue;
        }
        const double a = nzval[p];
#pragma unroll
        for (int u = 0; u < JT; ++u)
            if (u < nj) s[u] = fma(a, v[r + (size_t)u * ldv], s[u]);
    }
#pragma unroll
    for (int u = 0; u < JT; ++u)
        if (u < nj) W[(size_t)(j0 + u) * ldw + c] = s[u];
}

}  // namespace

// Ahat (k x k, this rank's PARTIAL sum when the columns of A are sharded) = W' * V[c_off : c_off + ncols, :] with
// W = (A[:, c0 : c0 + ncols])' * V: the columns [c0, c0 + ncols) of A given as a slice of Julia's CSC arrays (colptr
// keeps its global 1-based positions, p_base = colptr[c0] - 1), V the rows [row_lo, row_lo + nrows) of the basis that
// those columns touch, c_off = c0 - row_lo.  A single rank passes the whole matrix (p_base = row_lo = c_off = 0).
int32_t galerkin_csc_device(int64_t ncols, const long long* d_colptr, const long long* d_rowval, const double* d_nzval,
                            int64_t p_base, int64_t row_lo, int64_t nrows, int64_t c_off, const double* V, int k, int64_t ldv,
                            double* Ahat, int lda) {
    Ctx& cx = ctx();
    MOR_ARG(ncols >= 0 && nrows >= 0 && k >= 1 && k <= 4096 && lda >= k && c_off >= 0 && c_off + ncols <= nrows,
            "galerkin projection: bad dimensions");
    const int64_t ldw = std::max<int64_t>((ncols + 1) & ~int64_t(1), 2);
    DevBuf W, bad;
    MOR_TRY(W.alloc(sizeof(double) * (size_t)ldw * k));
    MOR_TRY(bad.alloc(sizeof(int)));
    MOR_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), cx.stream));
    if (ldw != ncols) MOR_CUDA(cudaMemsetAsync(W.p, 0, W.bytes, cx.stream));
    Timer t_spmm(MOR_T_SPMM), t_proj(MOR_T_PROJ);
    if (ncols > 0) {
        dim3 grid((unsigned)((ncols + 255) / 256), (unsigned)((k + JT - 1) / JT));
        t_spmm.start();
        csc_t_times_dense_kernel<<<grid, 256, 0, cx.stream>>>(ncols, d_colptr, d_rowval, d_nzval, p_base, row_lo, nrows, V, ldv, k,
                                                             W.as<double>(), ldw, bad.as<int>());
        t_spmm.stop();
        count_launch();
        MOR_CUDA(cudaGetLastError());
    }
    t_proj.start();
    MOR_TRY(gemm_tn(W.as<double>(), ldw, k, V + c_off, ldv, k, ncols, Ahat, lda));  // (A'V)' V = V' A V
    t_proj.stop();
    int h_bad = 0;
    MOR_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, cx.stream));
    MOR_CUDA(cudaStreamSynchronize(cx.stream));
    t_spmm.collect();
    t_proj.collect();
    MOR_ARG(h_bad == 0, "galerkin projection: row index outside 1..m in the CSC structure");
    return MOR_OK;
}

// Host-pointer form for one rank's column slice (called by the public entry point with the whole matrix, and by the
// single-process multi-GPU mode with each GPU's slice).  Ahat_dev: k x k on the device (ld = k), partial when sliced.
int32_t galerkin_csc_slice(int64_t ncols, const int64_t* colptr, const int64_t* rowval, const double* nzval, int64_t row_lo,
                           int64_t nrows, int64_t c_off, const double* V, int64_t k, int64_t ldv, double* Ahat_dev) {
    Ctx& cx = ctx();
    const int64_t p_base = colptr[0] - 1, nnz = colptr[ncols] - colptr[0];
    for (double& t : cx.timings) t = 0.0;
    DevBuf dcp, drv, dnz, dV;
    MOR_TRY(dcp.alloc(sizeof(long long) * (size_t)(ncols + 1)));
    MOR_TRY(drv.alloc(sizeof(long long) * (size_t)(nnz > 0 ? nnz : 1)));
    MOR_TRY(dnz.alloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1)));
    const int64_t ldd = std::max<int64_t>((nrows + 1) & ~int64_t(1), 2);
    MOR_TRY(dV.alloc(sizeof(double) * (size_t)ldd * (size_t)k));
    if (ldd != nrows) MOR_CUDA(cudaMemset2DAsync(dV.as<double>() + nrows, (size_t)ldd * 8, 0, 8, (size_t)k, cx.stream));
    Timer t_h2d(MOR_T_H2D);
    t_h2d.start();
    MOR_CUDA(cudaMemcpyAsync(dcp.p, colptr, sizeof(long long) * (size_t)(ncols + 1), cudaMemcpyHostToDevice, cx.stream));
    if (nnz > 0) {
        MOR_CUDA(cudaMemcpyAsync(drv.p, rowval + p_base, sizeof(long long) * (size_t)nnz, cudaMemcpyHostToDevice, cx.stream));
        MOR_CUDA(cudaMemcpyAsync(dnz.p, nzval + p_base, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, cx.stream));
    }
    MOR_TRY(h2d_2d(dV.as<double>(), ldd, V + row_lo, ldv, nrows, k, cx.stream));
    t_h2d.stop();
    int32_t st = galerkin_csc_device(ncols, dcp.as<long long>(), drv.as<long long>(), dnz.as<double>(), p_base, row_lo, nrows,
                                     c_off, dV.as<double>(), (int)k, ldd, Ahat_dev, (int)k);
    cudaStreamSynchronize(cx.stream);
    t_h2d.collect();
    return st;
}

}  // namespace mor

using namespace mor;

extern "C" {

int32_t mor_b200_galerkin_csc(int64_t m, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                              const double* V, int64_t k, int64_t ldv, double* Ahat, int64_t lda) {
    MOR_API_LOCK;
    if (multi_on() && !is_worker()) return multi_galerkin_csc(m, colptr, rowval, nzval, V, k, ldv, Ahat, lda);
    MOR_ARG(colptr && V && Ahat && m >= 1 && k >= 1 && ldv >= m && lda >= k, "mor_b200_galerkin_csc: bad arguments");
    MOR_ARG(colptr[0] == 1 && colptr[m] >= 1, "mor_b200_galerkin_csc: colptr must be 1-based (Julia SparseMatrixCSC)");
    const int64_t nnz = colptr[m] - 1;
    MOR_ARG(nnz == 0 || (rowval && nzval), "mor_b200_galerkin_csc: null rowval / nzval");
    MOR_TRY(require_ctx());
    Ctx& cx = ctx();
    DevBuf dA;
    MOR_TRY(dA.alloc(sizeof(double) * (size_t)k * k));
    int32_t st = galerkin_csc_slice(m, colptr, rowval, nzval, 0, m, 0, V, k, ldv, dA.as<double>());
    if (st == MOR_OK) {
        cudaError_t e = cudaMemcpy2DAsync(Ahat, (size_t)lda * 8, dA.p, (size_t)k * 8, (size_t)k * 8, (size_t)k,
                                          cudaMemcpyDeviceToHost, cx.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(cx.stream);
        if (e != cudaSuccess) st = cuda_fail(e, "copy A_hat", __FILE__, __LINE__);
    }
    return st;
}

// out (k x nvec, ld = ldo) = V' * Xv for nvec vectors (columns of Xv, m x nvec): g_hat = V' g,
// y_hat(0) = V' snapshot[:, 1].  V and Xv are this rank's row shards; the result is replicated.
int32_t mor_b200_project(const double* V, int64_t m, int64_t k, int64_t ldv, const double* Xv, int64_t nvec,
                         int64_t ldx, double* out, int64_t ldo) {
    MOR_API_LOCK;
    if (multi_on() && !is_worker()) return multi_project(V, m, k, ldv, Xv, nvec, ldx, out, ldo);
    MOR_ARG(V && Xv && out && m >= 0 && k >= 1 && nvec >= 1, "mor_b200_project: bad arguments");
    MOR_ARG(ldv >= (m > 0 ? m : 1) && ldx >= (m > 0 ? m : 1) && ldo >= k && k <= 4096 && nvec <= 4096,
            "mor_b200_project: bad dimensions");
    MOR_TRY(require_ctx());
    Ctx& cx = ctx();
    mor_mat_t dV = nullptr, dX = nullptr;
    DevBuf dC;
    int32_t st = mor_b200_mat_alloc(m, k, &dV);
    if (st == MOR_OK) st = mor_b200_mat_alloc(m, nvec, &dX);
    if (st == MOR_OK) st = dC.alloc(sizeof(double) * (size_t)k * nvec);
    if (st == MOR_OK) st = mor_b200_mat_upload(dV, V, ldv);
    if (st == MOR_OK) st = mor_b200_mat_upload(dX, Xv, ldx);
    if (st == MOR_OK) st = gemm_tn(dV->p, dV->ld, (int)k, dX->p, dX->ld, (int)nvec, m, dC.as<double>(), (int)k);
    if (st == MOR_OK && cx.nranks > 1) st = comm_allreduce_sum(dC.as<double>(), (size_t)k * nvec);
    if (st == MOR_OK) {
        cudaError_t e = cudaMemcpy2DAsync(out, (size_t)ldo * 8, dC.p, (size_t)k * 8, (size_t)k * 8, (size_t)nvec,
                                          cudaMemcpyDeviceToHost, cx.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(cx.stream);
        if (e != cudaSuccess) st = cuda_fail(e, "copy projection", __FILE__, __LINE__);
    }
    mor_b200_mat_free(dV);
    mor_b200_mat_free(dX);
    return st;
}

}  // extern "C"
