// Small dense (n x n, n <= a few thousand) FP64 kernels of the replicated "core" of the POD
// factorisation: scaling + blocked Cholesky of the Gram matrix, triangular inverse, a plain
// tiled GEMM and a few element-wise helpers.  Everything is column-major and lives in HBM/L2;
// these kernels are latency-, not throughput-critical (SURVEY.md kernels K2, K7b helpers).
#include <cmath>

#include "internal.h"

namespace mor {
namespace {

constexpr int NB = 32;  // Cholesky block size

// D[j] = sqrt(G[j][j]) (0 for an all-zero column); flags non-finite input
__global__ void diag_kernel(const double* __restrict__ G, int n, int ld, double* __restrict__ D,
                            SmallInfo* __restrict__ info) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double v = G[(size_t)j * ld + j];
    if (!isfinite(v)) {
        D[j] = 1.0;
        atomicExch(&info->nonfinite, 1);
    } else if (v <= 0.0) {
        D[j] = 0.0;
        atomicAdd(&info->nzero, 1);
    } else {
        D[j] = sqrt(v);
    }
}

// Gs = D^-1 G D^-1 (+ shift on the diagonal); zero columns become unit vectors e_j.
// delta2 accumulates ||D^-1 G D^-1 - I||_F^2 (without the shift).
__global__ void __launch_bounds__(256)
scale_kernel(const double* __restrict__ G, int n, int ld, const double* __restrict__ D, double shift,
             double* __restrict__ Gs, int lds, SmallInfo* __restrict__ info) {
    __shared__ double red[8];
    const int i = blockIdx.x * 32 + (threadIdx.x & 31);
    const int j0 = blockIdx.y * 32 + (threadIdx.x >> 5) * 4;
    double local = 0.0;
    bool bad = false;
    if (i < n) {
        const double di = D[i];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            if (j >= n) break;
            const double dj = D[j];
            const double g = G[(size_t)j * ld + i];
            double v;
            if (di > 0.0 && dj > 0.0) {
                v = g / (di * dj);
            } else {
                v = (i == j) ? 1.0 : 0.0;
            }
            if (!isfinite(g)) bad = true;
            const double e = v - (i == j ? 1.0 : 0.0);
            local += e * e;
            Gs[(size_t)j * lds + i] = v + (i == j ? shift : 0.0);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        atomicAdd(&info->delta2, s);
    }
    if (bad) atomicExch(&info->nonfinite, 1);
}

// ---- blocked Cholesky, upper: A = R'R, R overwrites the upper triangle ---------------------
__global__ void __launch_bounds__(1024)
chol_diag_kernel(double* __restrict__ A, int ld, int kb, int nb, SmallInfo* __restrict__ info) {
    __shared__ double S[NB][NB + 1];
    __shared__ int failed;
    const int i = threadIdx.x & 31, j = threadIdx.x >> 5;  // element (row i, col j)
    if (threadIdx.x == 0) failed = info->chol_fail;
    if (i < nb && j < nb) S[i][j] = A[(size_t)(kb + j) * ld + kb + i];
    __syncthreads();
    if (failed) return;
    for (int k = 0; k < nb; ++k) {
        if (threadIdx.x == 0) {
            const double piv = S[k][k];
            if (!(piv > 0.0) || !isfinite(piv)) {
                failed = kb + k + 1;
            } else {
                S[k][k] = sqrt(piv);
            }
        }
        __syncthreads();
        if (failed) break;
        if (i == k && j > k && j < nb) S[k][j] /= S[k][k];
        __syncthreads();
        if (i > k && j >= i && j < nb) S[i][j] -= S[k][i] * S[k][j];
        __syncthreads();
    }
    if (failed) {
        if (threadIdx.x == 0) info->chol_fail = failed;
        return;
    }
    if (i < nb && j < nb && j >= i) A[(size_t)(kb + j) * ld + kb + i] = S[i][j];
}

// rows kb..kb+nb of R to the right of the diagonal block: solve Rkk' * x = A[kb:kb+nb, j]
__global__ void __launch_bounds__(128)
chol_panel_kernel(double* __restrict__ A, int n, int ld, int kb, int nb, const SmallInfo* __restrict__ info) {
    __shared__ double S[NB][NB + 1];
    if (info->chol_fail) return;
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
        const int i = e % nb, j = e / nb;
        S[i][j] = A[(size_t)(kb + j) * ld + kb + i];
    }
    __syncthreads();
    const int j = kb + nb + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double x[NB];
    double* col = A + (size_t)j * ld + kb;
#pragma unroll
    for (int r = 0; r < NB; ++r) x[r] = r < nb ? col[r] : 0.0;
#pragma unroll
    for (int r = 0; r < NB; ++r) {
        if (r < nb) {
            double s = x[r];
#pragma unroll
            for (int l = 0; l < NB; ++l)
                if (l < r) s = fma(-S[l][r], x[l], s);
            x[r] = s / S[r][r];
        }
    }
#pragma unroll
    for (int r = 0; r < NB; ++r)
        if (r < nb) col[r] = x[r];
}

// trailing update A[i][j] -= sum_l R[kb+l][i] * R[kb+l][j] for i, j >= kb+nb (upper tiles only)
__global__ void __launch_bounds__(256)
chol_update_kernel(double* __restrict__ A, int n, int ld, int kb, int nb, const SmallInfo* __restrict__ info) {
    if (blockIdx.x > blockIdx.y) return;  // tile row > tile col: below the diagonal
    if (info->chol_fail) return;
    __shared__ double Ri[NB][NB + 1], Rj[NB][NB + 1];
    const int i0 = kb + nb + blockIdx.x * 32, j0 = kb + nb + blockIdx.y * 32;
    for (int e = threadIdx.x; e < NB * 32; e += blockDim.x) {
        const int l = e % NB, c = e / NB;
        Ri[l][c] = (l < nb && i0 + c < n) ? A[(size_t)(i0 + c) * ld + kb + l] : 0.0;
        Rj[l][c] = (l < nb && j0 + c < n) ? A[(size_t)(j0 + c) * ld + kb + l] : 0.0;
    }
    __syncthreads();
    const int ti = (threadIdx.x & 15) * 2, tj = (threadIdx.x >> 4) * 2;
    double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
#pragma unroll
    for (int l = 0; l < NB; ++l) {
        const double a0 = Ri[l][ti], a1 = Ri[l][ti + 1], b0 = Rj[l][tj], b1 = Rj[l][tj + 1];
        s00 = fma(a0, b0, s00);
        s01 = fma(a0, b1, s01);
        s10 = fma(a1, b0, s10);
        s11 = fma(a1, b1, s11);
    }
    const int i = i0 + ti, j = j0 + tj;
    if (i < n && j < n) A[(size_t)j * ld + i] -= s00;
    if (i < n && j + 1 < n) A[(size_t)(j + 1) * ld + i] -= s01;
    if (i + 1 < n && j < n) A[(size_t)j * ld + i + 1] -= s10;
    if (i + 1 < n && j + 1 < n) A[(size_t)(j + 1) * ld + i + 1] -= s11;
}

__global__ void zero_lower_kernel(double* __restrict__ A, int n, int ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i < n && j < n && i > j) A[(size_t)j * ld + i] = 0.0;
}

// column j of inv(R): back substitution with column-oriented updates, one CTA per column
__global__ void __launch_bounds__(256)
trinv_kernel(const double* __restrict__ R, int n, int ld, double* __restrict__ Rinv, int ldi) {
    extern __shared__ double x[];
    const int j = blockIdx.x;
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = (i == j) ? 1.0 : 0.0;
    __syncthreads();
    for (int i = j; i >= 0; --i) {
        const double xi = x[i] / R[(size_t)i * ld + i];
        __syncthreads();
        if (threadIdx.x == 0) x[i] = xi;
        const double* col = R + (size_t)i * ld;
        for (int r = threadIdx.x; r < i; r += blockDim.x) x[r] = fma(-col[r], xi, x[r]);
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) Rinv[(size_t)j * ldi + i] = (i <= j) ? x[i] : 0.0;
}

// C = alpha * op(A) * op(B) + beta * C, 32 x 32 tiles, 256 threads, 2 x 2 per thread
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
small_gemm_kernel(int m, int n, int k, double alpha, const double* __restrict__ A, int lda,
                  const double* __restrict__ B, int ldb, double beta, double* __restrict__ C, int ldc) {
    __shared__ double As[32][33], Bs[32][33];  // As[l][i], Bs[l][j]
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int ti = (threadIdx.x & 15) * 2, tj = (threadIdx.x >> 4) * 2;
    double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
    for (int l0 = 0; l0 < k; l0 += 32) {
        for (int e = threadIdx.x; e < 1024; e += 256) {
            // op(A)[i][l]: A[i + l*lda] or, transposed, A[l + i*lda] -- pick the contiguous index first
            int a_l, a_i, b_l, b_j;
            if (TA) { a_l = e & 31; a_i = e >> 5; } else { a_i = e & 31; a_l = e >> 5; }
            if (TB) { b_j = e & 31; b_l = e >> 5; } else { b_l = e & 31; b_j = e >> 5; }
            const int gi = i0 + a_i, gl = l0 + a_l;
            As[a_l][a_i] = (gi < m && gl < k) ? (TA ? A[(size_t)gi * lda + gl] : A[(size_t)gl * lda + gi]) : 0.0;
            const int gj = j0 + b_j, gl2 = l0 + b_l;
            Bs[b_l][b_j] = (gj < n && gl2 < k) ? (TB ? B[(size_t)gl2 * ldb + gj] : B[(size_t)gj * ldb + gl2]) : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int l = 0; l < 32; ++l) {
            const double a0 = As[l][ti], a1 = As[l][ti + 1], b0 = Bs[l][tj], b1 = Bs[l][tj + 1];
            s00 = fma(a0, b0, s00);
            s01 = fma(a0, b1, s01);
            s10 = fma(a1, b0, s10);
            s11 = fma(a1, b1, s11);
        }
        __syncthreads();
    }
    const int i = i0 + ti, j = j0 + tj;
    auto put = [&](int ii, int jj, double s) {
        if (ii < m && jj < n) {
            double* c = C + (size_t)jj * ldc + ii;
            *c = (beta == 0.0) ? alpha * s : alpha * s + beta * *c;
        }
    };
    put(i, j, s00);
    put(i, j + 1, s01);
    put(i + 1, j, s10);
    put(i + 1, j + 1, s11);
}

// A[i][j] *= (rows ? d[i] : d[j]) or /= when inv
__global__ void scale_rc_kernel(double* __restrict__ A, int m, int n, int ld, const double* __restrict__ d,
                                int rows, int inv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= m || j >= n) return;
    double s = rows ? d[i] : d[j];
    if (s == 0.0) s = 1.0;
    if (inv) s = 1.0 / s;
    A[(size_t)j * ld + i] *= s;
}

__global__ void set_identity_kernel(double* __restrict__ A, int n, int ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i < n && j < n) A[(size_t)j * ld + i] = (i == j) ? 1.0 : 0.0;
}

// zero the rows i of A with flag[i] == 0.0 (deflated, all-zero snapshot columns)
__global__ void zero_rows_kernel(double* __restrict__ A, int m, int n, int ld, const double* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i < m && j < n && flag[i] == 0.0) A[(size_t)j * ld + i] = 0.0;
}

// out[i][j] = in[j][i]
__global__ void transpose_kernel(const double* __restrict__ in, long long rows, long long cols, long long ldin,
                                 double* __restrict__ out, long long ldout) {
    __shared__ double tile[32][33];
    const long long r0 = (long long)blockIdx.x * 32, c0 = (long long)blockIdx.y * 32;
    for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
        const long long r = r0 + threadIdx.x, c = c0 + cc;
        tile[cc][threadIdx.x] = (r < rows && c < cols) ? in[c * ldin + r] : 0.0;
    }
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        const long long c = c0 + threadIdx.x, r = r0 + rr;
        if (r < rows && c < cols) out[r * ldout + c] = tile[threadIdx.x][rr];
    }
}

// Solve A X = B by LU with partial pivoting (what Julia's `\` does for a general square matrix:
// dgetrf + dgetrs), one CTA, A (n x n) and B (n x nrhs) column-major in L2, both overwritten.
// transpose != 0 solves A' X = B instead (A is read transposed).  *fail = 1-based zero pivot.
__global__ void __launch_bounds__(1024)
lu_solve_kernel(double* __restrict__ A, int n, int lda, int transpose, double* __restrict__ B, int nrhs, int ldb,
                int* __restrict__ fail) {
    __shared__ double sv[32];
    __shared__ int si[32];
    __shared__ int piv_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto at = [&](int r, int c) -> double& { return transpose ? A[(size_t)r * lda + c] : A[(size_t)c * lda + r]; };
    for (int j = 0; j < n; ++j) {
        // pivot: first row of maximal |A[r][j]|, r >= j (idamax semantics)
        double bv = -1.0;
        int bi = 0x7fffffff;
        for (int r = j + tid; r < n; r += 1024) {
            const double v = fabs(at(r, j));
            if (v > bv) {
                bv = v;
                bi = r;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ov > bv || (ov == bv && oi < bi)) {
                bv = ov;
                bi = oi;
            }
        }
        if (lane == 0) {
            sv[warp] = bv;
            si[warp] = bi;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 32; ++w)
                if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) {
                    bv = sv[w];
                    bi = si[w];
                }
            piv_s = bi;
            if (!(bv > 0.0) && *fail == 0) *fail = j + 1;
        }
        __syncthreads();
        const int pv = piv_s;
        if (pv != j && pv < n) {
            for (int c = tid; c < n; c += 1024) {
                const double t = at(j, c);
                at(j, c) = at(pv, c);
                at(pv, c) = t;
            }
            for (int c = tid; c < nrhs; c += 1024) {
                const double t = B[(size_t)c * ldb + j];
                B[(size_t)c * ldb + j] = B[(size_t)c * ldb + pv];
                B[(size_t)c * ldb + pv] = t;
            }
        }
        __syncthreads();
        const double d = at(j, j);
        for (int r = j + 1 + tid; r < n; r += 1024) at(r, j) = at(r, j) / d;
        __syncthreads();
        const int rem = n - j - 1;
        for (int e = tid; e < rem * rem; e += 1024) {
            const int r = j + 1 + e % rem, c = j + 1 + e / rem;
            at(r, c) = fma(-at(r, j), at(j, c), at(r, c));
        }
        for (int e = tid; e < rem * nrhs; e += 1024) {
            const int r = j + 1 + e % rem, c = e / rem;
            B[(size_t)c * ldb + r] = fma(-at(r, j), B[(size_t)c * ldb + j], B[(size_t)c * ldb + r]);
        }
        __syncthreads();
    }
    for (int j = n - 1; j >= 0; --j) {
        const double d = at(j, j);
        for (int c = tid; c < nrhs; c += 1024) B[(size_t)c * ldb + j] /= d;
        __syncthreads();
        for (int e = tid; e < j * nrhs; e += 1024) {
            const int r = e % j, c = e / j;
            B[(size_t)c * ldb + r] = fma(-at(r, j), B[(size_t)c * ldb + j], B[(size_t)c * ldb + r]);
        }
        __syncthreads();
    }
}

// out[i][j] = (row0 <= idx[i]-1 < row0 + m) ? U[idx[i]-1-row0][j] : 0   (k x k gather of the selected rows)
__global__ void gather_rows_kernel(const double* __restrict__ U, long long m, long long ld, long long row0,
                                   const long long* __restrict__ idx, int k, double* __restrict__ out, int ldo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= k) return;
    const long long r = idx[i] - 1 - row0;
    out[(size_t)j * ldo + i] = (r >= 0 && r < m) ? U[(size_t)j * ld + r] : 0.0;
}

}  // namespace

int32_t small_lu_solve(double* A, int n, int lda, bool transpose, double* B, int nrhs, int ldb, int* d_fail) {
    MOR_CUDA(cudaMemsetAsync(d_fail, 0, sizeof(int), ctx().stream));
    lu_solve_kernel<<<1, 1024, 0, ctx().stream>>>(A, n, lda, transpose ? 1 : 0, B, nrhs, ldb, d_fail);
    count_launch();
    MOR_CUDA(cudaGetLastError());
    return MOR_OK;
}

int32_t gather_rows(const double* U, int64_t m, int64_t ld, int64_t row0, const long long* d_idx, int k, double* out,
                    int ldo) {
    gather_rows_kernel<<<dim3((k + 127) / 128, k), 128, 0, ctx().stream>>>(U, m, ld, row0, d_idx, k, out, ldo);
    count_launch();
    MOR_CUDA(cudaGetLastError());
    return MOR_OK;
}

int32_t small_info_reset(SmallInfo* d_info) {
    MOR_CUDA(cudaMemsetAsync(d_info, 0, sizeof(SmallInfo), ctx().stream));
    return MOR_OK;
}

int32_t small_info_read(const SmallInfo* d_info, SmallInfo* h) {
    MOR_CUDA(cudaMemcpyAsync(h, d_info, sizeof(SmallInfo), cudaMemcpyDeviceToHost, ctx().stream));
    MOR_CUDA(cudaStreamSynchronize(ctx().stream));
    return MOR_OK;
}

int32_t small_diag(const double* G, int n, int ld, double* D, SmallInfo* d_info) {
    diag_kernel<<<(n + 127) / 128, 128, 0, ctx().stream>>>(G, n, ld, D, d_info);
    count_launch();
    MOR_CUDA(cudaGetLastError());
    return MOR_OK;
}

int32_t small_scale(const double* G, int n, int ld, const double* D, double shift, double* Gs, int lds,
                    SmallInfo* d_info) {
    dim3 grid((n + 31) / 32, (n + 31) / 32);
    scale_kernel<<<grid, 256, 0, ctx().stream>>>(G, n, ld, D, shift, Gs, lds, d_info);
    count_launch();
    MOR_CUDA(cudaGetLastError());
    return MOR_OK;
}

int32_t small_chol_upper(double* A, int n, int ld, SmallInfo* d_info) {
    cudaStream_t st = ctx().stream;
    for (int kb = 0; kb < n; kb += NB) {
        const int nb = n - kb < NB ? n - kb : NB;
        chol_diag_kernel<<<1, 1024, 0, st>>>(A, ld, kb, nb, d_info);
        const int rest = n - kb - nb;
        if (rest > 0) {
            chol_panel_kernel<<<(rest + 127) / 128, 128, 0, st>>>(A, n, ld, kb, nb, d_info);
            const int T = (rest + 31) / 32;
            chol_update_kernel<<<dim3(T, T), 256, 0, st>>>(A, n, ld, kb, nb, d_info);
            count_launch(2);
        }
        count_launch();
    }
    zero_lower_kernel<<<dim3((n + 127) / 128, n), 128, 0, st>>>(A, n, ld);
    count_launch();
    MOR_CUDA(cudaGetLastError());
    return MOR_OK;
}

int32_t small_trinv_upper(const double* R, int n, int ld, double* Rinv, int ldi) {
    trinv_kernel<<<n, 256, sizeof(double) * (size_t)n, ctx().stream>>>(R, n, ld, Rinv, ldi);
    count_launch();
    MOR_CUDA(cudaGetLastError());
    return MOR_OK;
}

int32_t small_gemm(bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int lda, const double* B,
                   int ldb, double beta, double* C, int ldc) {
    if (m <= 0 || n <= 0) return MOR_OK;
    dim3 grid((m + 31) / 32, (n + 31) / 32);
    cudaStream_t st = ctx().stream;
    if (!ta && !tb) small_gemm_kernel<false, false><<<grid, 256, 0, st>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
    else if (ta && !tb) small_gemm_kernel<true, false><<<grid, 256, 0, st>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
    else if (!ta && tb) small_gemm_kernel<false, true><<<grid, 256, 0, st>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
    else small_gemm_kernel<true, true><<<grid, 256, 0, st>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
    count_launch();
    MOR_CUDA(cudaGetLastError());
    return MOR_OK;
}

int32_t small_scale_rc(double* A, int m, int n, int ld, const double* d, bool rows, bool inv) {
    if (m <= 0 || n <= 0) return MOR_OK;
    scale_rc_kernel<<<dim3((m + 127) / 128, n), 128, 0, ctx().stream>>>(A, m, n, ld, d, rows ? 1 : 0, inv ? 1 : 0);
    count_launch();
    MOR_CUDA(cudaGetLastError());
    return MOR_OK;
}

int32_t small_identity(double* A, int n, int ld) {
    set_identity_kernel<<<dim3((n + 127) / 128, n), 128, 0, ctx().stream>>>(A, n, ld);
    count_launch();
    MOR_CUDA(cudaGetLastError());
    return MOR_OK;
}

int32_t small_zero_rows(double* A, int m, int n, int ld, const double* flag) {
    zero_rows_kernel<<<dim3((m + 127) / 128, n), 128, 0, ctx().stream>>>(A, m, n, ld, flag);
    count_launch();
    MOR_CUDA(cudaGetLastError());
    return MOR_OK;
}

int32_t transpose(const double* in, int64_t rows, int64_t cols, int64_t ldin, double* out, int64_t ldout) {
    if (rows <= 0 || cols <= 0) return MOR_OK;
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
    MOR_ARG(grid.y <= 65535u, "transpose: too many columns");
    transpose_kernel<<<grid, dim3(32, 8), 0, ctx().stream>>>(in, rows, cols, ldin, out, ldout);
    count_launch();
    MOR_CUDA(cudaGetLastError());
    return MOR_OK;
}

}  // namespace mor
