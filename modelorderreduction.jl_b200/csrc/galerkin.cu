// Galerkin projection of the linear part of the reduced model -- the second "next" row of the hot
// path, reference src/deim.jl:154-155 and 261:
//     A_hat = V' * A * V     (A: SparseMatrixCSC from separate_terms, utils.jl:126)
//     g_hat = V' * g ,   y_hat(0) = V' * snapshot[:, 1]
// A is taken in Julia's own CSC arrays (colptr / rowval 1-based Int64, nzval Float64).  CSC of A
// read column by column is CSR of A', so W = A' V needs no atomics: one thread per (column c of A,
// basis vector j) gathers W[c, j] = sum_p nzval[p] * V[rowval[p], j]; then A_hat = W' V runs on the
// DMMA contraction kernel (gemm_tn).  The projections of vectors are gemm_tn with one column.
// Single rank for the sparse product (the row halo of a sharded V is not exchanged yet); the
// vector projection is row-sharded like everything else.
#include "internal.h"

namespace mor {
namespace {

__global__ void __launch_bounds__(256)
csc_t_times_dense_kernel(long long ncols, const long long* __restrict__ colptr, const long long* __restrict__ rowval,
                         const double* __restrict__ nzval, long long nrows, const double* __restrict__ V, long long ldv,
                         double* __restrict__ W, long long ldw, int* __restrict__ bad) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (c >= ncols) return;
    const double* v = V + (size_t)j * ldv;
    double s = 0.0;
    for (long long p = colptr[c] - 1; p < colptr[c + 1] - 1; ++p) {
        const long long r = rowval[p] - 1;
        if (r < 0 || r >= nrows) {
            *bad = 1;
            continue;
        }
        s = fma(nzval[p], v[r], s);
    }
    W[(size_t)j * ldw + c] = s;
}

}  // namespace

int32_t galerkin_csc_device(int64_t m, const long long* d_colptr, const long long* d_rowval, const double* d_nzval,
                            const double* V, int k, int64_t ldv, double* Ahat, int lda) {
    Ctx& cx = ctx();
    MOR_ARG(cx.nranks == 1, "galerkin projection of a sparse operator is single-rank for now (no halo exchange)");
    MOR_ARG(m >= 1 && k >= 1 && k <= 4096 && lda >= k, "galerkin projection: bad dimensions");
    const int64_t ldw = (m + 1) & ~int64_t(1);
    DevBuf W, bad;
    MOR_TRY(W.alloc(sizeof(double) * (size_t)ldw * k));
    MOR_TRY(bad.alloc(sizeof(int)));
    MOR_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), cx.stream));
    if (ldw != m) MOR_CUDA(cudaMemset2DAsync(W.as<double>() + m, (size_t)ldw * 8, 0, 8, (size_t)k, cx.stream));
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)k);
    csc_t_times_dense_kernel<<<grid, 256, 0, cx.stream>>>(m, d_colptr, d_rowval, d_nzval, m, V, ldv, W.as<double>(), ldw,
                                                         bad.as<int>());
    count_launch();
    MOR_CUDA(cudaGetLastError());
    MOR_TRY(gemm_tn(W.as<double>(), ldw, k, V, ldv, k, m, Ahat, lda));  // (A'V)' V = V' A V
    int h_bad = 0;
    MOR_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, cx.stream));
    MOR_CUDA(cudaStreamSynchronize(cx.stream));
    MOR_ARG(h_bad == 0, "galerkin projection: row index outside 1..m in the CSC structure");
    return MOR_OK;
}

}  // namespace mor

using namespace mor;

extern "C" {

int32_t mor_b200_galerkin_csc(int64_t m, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                              const double* V, int64_t k, int64_t ldv, double* Ahat, int64_t lda) {
    MOR_API_LOCK;
    MOR_ROUTE_SOLO(mor_b200_galerkin_csc(m, colptr, rowval, nzval, V, k, ldv, Ahat, lda));
    MOR_ARG(colptr && V && Ahat && m >= 1 && k >= 1 && ldv >= m && lda >= k, "mor_b200_galerkin_csc: bad arguments");
    MOR_ARG(colptr[0] == 1 && colptr[m] >= 1, "mor_b200_galerkin_csc: colptr must be 1-based (Julia SparseMatrixCSC)");
    const int64_t nnz = colptr[m] - 1;
    MOR_ARG(nnz == 0 || (rowval && nzval), "mor_b200_galerkin_csc: null rowval / nzval");
    MOR_TRY(require_ctx());
    Ctx& cx = ctx();
    DevBuf dcp, drv, dnz, dA;
    mor_mat_t dV = nullptr;
    MOR_TRY(dcp.alloc(sizeof(long long) * (size_t)(m + 1)));
    MOR_TRY(drv.alloc(sizeof(long long) * (size_t)(nnz > 0 ? nnz : 1)));
    MOR_TRY(dnz.alloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1)));
    MOR_TRY(dA.alloc(sizeof(double) * (size_t)k * k));
    MOR_CUDA(cudaMemcpyAsync(dcp.p, colptr, sizeof(long long) * (size_t)(m + 1), cudaMemcpyHostToDevice, cx.stream));
    if (nnz > 0) {
        MOR_CUDA(cudaMemcpyAsync(drv.p, rowval, sizeof(long long) * (size_t)nnz, cudaMemcpyHostToDevice, cx.stream));
        MOR_CUDA(cudaMemcpyAsync(dnz.p, nzval, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, cx.stream));
    }
    int32_t st = mor_b200_mat_alloc(m, k, &dV);
    if (st == MOR_OK) st = mor_b200_mat_upload(dV, V, ldv);
    if (st == MOR_OK)
        st = galerkin_csc_device(m, dcp.as<long long>(), drv.as<long long>(), dnz.as<double>(), dV->p, (int)k, dV->ld,
                                 dA.as<double>(), (int)k);
    if (st == MOR_OK) {
        cudaError_t e = cudaMemcpy2DAsync(Ahat, (size_t)lda * 8, dA.p, (size_t)k * 8, (size_t)k * 8, (size_t)k,
                                          cudaMemcpyDeviceToHost, cx.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(cx.stream);
        if (e != cudaSuccess) st = cuda_fail(e, "copy A_hat", __FILE__, __LINE__);
    }
    mor_b200_mat_free(dV);
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
