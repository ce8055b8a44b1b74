// POD driver: the numeric work behind reduce!(pod, SVD()/TSVD()/RSVD()) -- reference
// src/DataReduction/POD.jl:156-166 (svd), 195-202 (tsvd), 231-238 (rsvd) and the seams
// POD.jl:8-27 they call.
//
// Tall-skinny factorisation of T (rows x n, rows >= n, row-sharded over the ranks):
//   pass p:  G = T'T  (gemm_tn, DMMA)  -> all-reduce over ranks -> scale to unit diagonal ->
//            Cholesky R_p (shifted if it breaks down) -> R = R_p * R;
//            stop when the scaled Gram matrix is close enough to I, else T <- T * inv(R_p)
//            in place (gemm_nn, DMMA) and go again (CholeskyQR2 / shifted CholeskyQR3).
//   core:    one-sided Jacobi SVD of the n x n factor R = U_R S V'.
//   basis:   left vectors of T = T_final * M(:, 1:k), M = V S^-1 (one pass) or inv(R_last) U_R.
// A wide input (m < n) is factored through its transpose, so U comes from the small side.
// RSVD is Halko-Martinsson-Tropp: Y = X Omega, Q = orth(Y) by the same CholeskyQR passes,
// B' = X'Q, Jacobi SVD of B', U = Q U_B.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "internal.h"

namespace mor {
namespace {

constexpr double SINGLE_PASS_DELTA = 0.9;   // ||Gs - I||_2 <= ||Gs - I||_F < 0.9: kappa(Gs) < 19, one Gram pass is exact enough
constexpr double CONVERGED_DELTA = 0.05;    // later passes: T is orthonormal enough for the final correction
constexpr int MAX_PASSES = 6;

// flip column j of V (rows x r) and of M so that V's largest-|entry| (first on ties) is positive
__global__ void __launch_bounds__(256)
sign_fix_kernel(double* __restrict__ V, int vrows, int ldv, double* __restrict__ M, int mrows, int ldm) {
    __shared__ double sv[8];
    __shared__ int si[8];
    __shared__ int flip;
    const int j = blockIdx.x;
    double* v = V + (size_t)j * ldv;
    double best = -1.0;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < vrows; i += blockDim.x) {
        const double a = fabs(v[i]);
        if (a > best) {
            best = a;
            bi = i;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ob > best || (ob == best && oi < bi)) {
            best = ob;
            bi = oi;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        sv[threadIdx.x >> 5] = best;
        si[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (sv[w] > best || (sv[w] == best && si[w] < bi)) {
                best = sv[w];
                bi = si[w];
            }
        flip = (bi < vrows && v[bi] < 0.0) ? 1 : 0;
    }
    __syncthreads();
    if (!flip) return;
    for (int i = threadIdx.x; i < vrows; i += blockDim.x) v[i] = -v[i];
    if (M) {
        double* mc = M + (size_t)j * ldm;
        for (int i = threadIdx.x; i < mrows; i += blockDim.x) mc[i] = -mc[i];
    }
}

int32_t sign_fix(double* V, int vrows, int ldv, int r, double* M, int mrows, int ldm) {
    if (r <= 0) return MOR_OK;
    sign_fix_kernel<<<r, 256, 0, ctx().stream>>>(V, vrows, ldv, M, mrows, ldm);
    count_launch();
    MOR_CUDA(cudaGetLastError());
    return MOR_OK;
}

int32_t upload_perm(const std::vector<int>& order, DevBuf& d) {
    MOR_TRY(d.alloc(sizeof(int) * order.size()));
    MOR_CUDA(cudaMemcpyAsync(d.p, order.data(), sizeof(int) * order.size(), cudaMemcpyHostToDevice, ctx().stream));
    return MOR_OK;
}

int32_t upload_vec(const std::vector<double>& v, DevBuf& d) {
    MOR_TRY(d.alloc(sizeof(double) * v.size()));
    MOR_CUDA(cudaMemcpyAsync(d.p, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice, ctx().stream));
    return MOR_OK;
}

struct Timers {
    Timer total{MOR_T_TOTAL}, gram{MOR_T_GRAM}, apply{MOR_T_APPLY}, core{MOR_T_CORE}, basis{MOR_T_BASIS},
        comm{MOR_T_COMM}, sketch{MOR_T_SKETCH}, h2d{MOR_T_H2D}, d2h{MOR_T_D2H};
    void collect() {
        total.collect(); gram.collect(); apply.collect(); core.collect(); basis.collect();
        comm.collect(); sketch.collect(); h2d.collect(); d2h.collect();
    }
};

// One CholeskyQR pass on T (trows x n): Gram, all-reduce, scaled (shifted) Cholesky.
// Outputs: Rs (n x n upper, factor of the SCALED Gram matrix), D (column norms, 0 for zero
// columns), delta = ||Gs - I||_F, shift used.
struct PassOut {
    double delta = 0.0, shift = 0.0;
    int nzero = 0;
};
int32_t gram_pass(const double* T, int64_t trows, int n, int64_t ldt, bool replicated, double* G, double* Rs,
                  double* D, SmallInfo* d_info, Timers& tm, PassOut* out, const double* G_pre = nullptr) {
    tm.gram.start();
    if (G_pre) {  // this rank's Gram matrix was accumulated while the rows were still arriving
        MOR_CUDA(cudaMemcpyAsync(G, G_pre, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToDevice, ctx().stream));
    } else {
        MOR_TRY(gemm_tn(T, ldt, n, T, ldt, n, trows, G, n));
    }
    tm.gram.stop();
    if (!replicated && ctx().nranks > 1) {
        tm.comm.start();
        MOR_TRY(comm_allreduce_sum(G, (size_t)n * n));
        tm.comm.stop();
    }
    tm.core.start();
    DebugLaps dbg("gram_pass: scaled Cholesky");
    double shift = 0.0;
    SmallInfo h;
    for (int attempt = 0;; ++attempt) {
        MOR_TRY(small_info_reset(d_info));
        MOR_TRY(small_diag(G, n, n, D, d_info));
        MOR_TRY(small_scale(G, n, n, D, shift, Rs, n, d_info));
        dbg.lap("diag + scale");
        MOR_TRY(small_chol_upper(Rs, n, n, d_info));
        MOR_TRY(small_info_read(d_info, &h));
        dbg.lap("blocked Cholesky");
        if (h.nonfinite) {
            tm.core.stop();
            set_error("POD: the snapshot matrix contains Inf or NaN");
            return MOR_ERR_NONFINITE;
        }
        if (!h.chol_fail) break;
        // numerically rank deficient: shifted Cholesky (Fukaya et al., shifted CholeskyQR3)
        shift = (shift == 0.0) ? 1e-13 * n : shift * 100.0;
        if (shift > 1e-3) {
            tm.core.stop();
            set_error("POD: shifted Cholesky of the Gram matrix keeps breaking down");
            return MOR_ERR_NOCONV;
        }
    }
    tm.core.stop();
    out->delta = std::sqrt(h.delta2);
    out->shift = shift;
    out->nzero = h.nzero;
    return MOR_OK;
}

// T <- T * inv(Rs * diag(D))   (in place, upper triangular)
int32_t apply_rinv(double* T, int64_t trows, int n, int64_t ldt, const double* Rs, const double* D, double* Rinv,
                   Timers& tm) {
    tm.core.start();
    MOR_TRY(small_trinv_upper(Rs, n, n, Rinv, n));
    MOR_TRY(small_scale_rc(Rinv, n, n, n, D, /*rows=*/true, /*inv=*/true));
    tm.core.stop();
    tm.apply.start();
    MOR_TRY(gemm_nn(T, ldt, n, Rinv, n, n, trows, T, ldt, /*upper=*/1));
    tm.apply.stop();
    return MOR_OK;
}

// Orthonormalise the columns of Y (rows x l) in place by CholeskyQR passes.
int32_t orthonormalize(double* Y, int64_t rows, int l, int64_t ldy, bool replicated, Timers& tm, int* passes_out) {
    DevBuf G, Rs, Rinv, D, info;
    const size_t nn = sizeof(double) * (size_t)l * l;
    MOR_TRY(G.alloc(nn));
    MOR_TRY(Rs.alloc(nn));
    MOR_TRY(Rinv.alloc(nn));
    MOR_TRY(D.alloc(sizeof(double) * (size_t)l));
    MOR_TRY(info.alloc(sizeof(SmallInfo)));
    int passes = 0;
    for (;;) {
        PassOut po;
        MOR_TRY(gram_pass(Y, rows, l, ldy, replicated, G.as<double>(), Rs.as<double>(), D.as<double>(),
                          info.as<SmallInfo>(), tm, &po));
        // the Gram matrix just computed certifies the current Y: orthonormal to rounding?
        if (passes > 0 && po.shift == 0.0 && po.delta <= 64.0 * l * 2.3e-16) break;
        if (passes >= MAX_PASSES) {
            set_error("orthonormalisation did not converge in " + std::to_string(MAX_PASSES) + " CholeskyQR passes");
            return MOR_ERR_NOCONV;
        }
        MOR_TRY(apply_rinv(Y, rows, l, ldy, Rs.as<double>(), D.as<double>(), Rinv.as<double>(), tm));
        ++passes;
        // after a well-conditioned pass one more apply reaches rounding level; skip the
        // certifying Gram when the previous one was already close to I
        if (passes >= 2 && po.shift == 0.0 && po.delta < 1e-6) break;
    }
    if (passes_out) *passes_out = passes;
    return MOR_OK;
}

// Factor the tall matrix T (trows x n, trows_global >= n).  May overwrite T when
// `may_overwrite`; otherwise a private copy is made the first time an in-place pass is
// needed (returned through h->T_owned).  Fills h->M, h->Vs, h->sigma, h->T, passes, sweeps.
int32_t factor_tall(mor_pod_s* h, double* T, int64_t trows, int n, int64_t ldt, bool may_overwrite,
                    bool force_two_pass, bool replicated, Timers& tm, const double* G_first = nullptr) {
    const size_t nn = sizeof(double) * (size_t)n * n;
    DevBuf G, Rs, Rtot, Rtmp, Rinv, D, D1, info, A, V;
    MOR_TRY(G.alloc(nn));
    MOR_TRY(Rs.alloc(nn));
    MOR_TRY(Rtot.alloc(nn));
    MOR_TRY(Rtmp.alloc(nn));
    MOR_TRY(Rinv.alloc(nn));
    MOR_TRY(V.alloc(nn));
    MOR_TRY(D.alloc(sizeof(double) * (size_t)n));
    MOR_TRY(D1.alloc(sizeof(double) * (size_t)n));
    MOR_TRY(info.alloc(sizeof(SmallInfo)));
    cudaStream_t st = ctx().stream;

    bool single_pass = false;
    int passes = 0;
    for (;;) {
        PassOut po;
        MOR_TRY(gram_pass(T, trows, n, ldt, replicated, G.as<double>(), Rs.as<double>(), D.as<double>(),
                          info.as<SmallInfo>(), tm, &po, passes == 0 ? G_first : nullptr));
        ++passes;
        tm.core.start();
        // R_p = Rs * diag(D)  (zero columns act as D = 1)
        MOR_CUDA(cudaMemcpyAsync(Rtmp.p, Rs.p, nn, cudaMemcpyDeviceToDevice, st));
        MOR_TRY(small_scale_rc(Rtmp.as<double>(), n, n, n, D.as<double>(), /*rows=*/false, /*inv=*/false));
        if (passes == 1) {
            MOR_CUDA(cudaMemcpyAsync(Rtot.p, Rtmp.p, nn, cudaMemcpyDeviceToDevice, st));
            MOR_CUDA(cudaMemcpyAsync(D1.p, D.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
        } else {
            MOR_TRY(small_gemm(false, false, n, n, n, 1.0, Rtmp.as<double>(), n, Rtot.as<double>(), n, 0.0,
                               G.as<double>(), n));
            MOR_CUDA(cudaMemcpyAsync(Rtot.p, G.p, nn, cudaMemcpyDeviceToDevice, st));
        }
        tm.core.stop();
        if (passes == 1 && po.shift == 0.0 && po.delta < SINGLE_PASS_DELTA && !force_two_pass) {
            single_pass = true;
            break;
        }
        if (passes > 1 && po.shift == 0.0 && po.delta < CONVERGED_DELTA) break;
        if (passes >= MAX_PASSES) {
            set_error("POD: CholeskyQR did not converge in " + std::to_string(MAX_PASSES) + " passes");
            return MOR_ERR_NOCONV;
        }
        if (!may_overwrite) {  // private copy of T before the first in-place update
            DevBuf copy;
            MOR_TRY(copy.alloc(sizeof(double) * (size_t)ldt * (size_t)n));
            MOR_CUDA(cudaMemcpyAsync(copy.p, T, sizeof(double) * (size_t)ldt * (size_t)n, cudaMemcpyDeviceToDevice, st));
            h->T_owned.take(copy);
            T = h->T_owned.as<double>();
            may_overwrite = true;
        }
        MOR_TRY(apply_rinv(T, trows, n, ldt, Rs.as<double>(), D.as<double>(), Rinv.as<double>(), tm));
    }

    // ---- core SVD of the n x n factor -----------------------------------------------------
    // R = U_R S V'.  When the POD basis comes from the tall side (U = T_final * inv(R_last) * U_R)
    // the rotations need not be accumulated: the Jacobi sweeps then move half the data, and the
    // right vectors are recovered as V = R' U_R S^-1 (accurate to eps * s_1 / s_j -- they are
    // only an optional output there).  For a wide input the small-side vectors ARE the basis, so
    // V is accumulated and stays orthonormal to rounding for every column.
    tm.core.start();
    DebugLaps dbg("factor_tall: core SVD");
    const bool accumulate_v = replicated || n < 64;
    // One pass, tall side: only V and S are needed (U = X V S^-1), so the Jacobi runs on the columns of R' -- the
    // rotated columns converge to v_j s_j directly (no accumulation, no inv(R), no n^3 products afterwards), and
    // one-sided Jacobi on the transposed triangular factor converges at least as fast (Veselic-Hari, Drmac-Veselic).
    const bool on_transpose = single_pass && !accumulate_v;
    MOR_TRY(A.alloc(nn));
    MOR_CUDA(cudaMemcpyAsync(A.p, Rtot.p, nn, cudaMemcpyDeviceToDevice, st));
    // rows of R that belong to all-zero snapshot columns carry no energy
    MOR_TRY(small_zero_rows(A.as<double>(), n, n, n, D1.as<double>()));
    if (!accumulate_v && !on_transpose) MOR_CUDA(cudaMemcpyAsync(Rtot.p, A.p, nn, cudaMemcpyDeviceToDevice, st));  // keep R (deflated)
    if (on_transpose) {
        MOR_TRY(transpose(A.as<double>(), n, n, n, G.as<double>(), n));
        MOR_CUDA(cudaMemcpyAsync(A.p, G.p, nn, cudaMemcpyDeviceToDevice, st));
    }
    std::vector<int> order;
    int sweeps = 0;
    dbg.lap("copy R / deflate");
    int32_t js = jacobi_svd(A.as<double>(), n, n, n, accumulate_v ? V.as<double>() : nullptr, n, h->sigma, order, &sweeps,
                            /*well_scaled=*/on_transpose);
    if (js != MOR_OK) {
        tm.core.stop();
        return js;
    }
    dbg.lap("jacobi_svd");
    DevBuf perm, inv_sigma, inv_vn, inv_both;
    MOR_TRY(upload_perm(order, perm));
    std::vector<double> is((size_t)n), iv((size_t)n, 1.0), ib((size_t)n);
    if (accumulate_v) {
        // the accumulated rotations drift from unit norm by ~(rotations per column) * eps: renormalise
        std::vector<double> vn;
        MOR_TRY(col_norms(V.as<double>(), n, n, n, vn));
        for (int j = 0; j < n; ++j) iv[j] = vn[order[j]] > 0.0 ? 1.0 / vn[order[j]] : 1.0;
    }
    for (int j = 0; j < n; ++j) {
        is[j] = h->sigma[j] > 0.0 ? 1.0 / h->sigma[j] : 0.0;
        ib[j] = on_transpose ? is[j] * is[j] : is[j] * iv[j];
    }
    MOR_TRY(upload_vec(is, inv_sigma));
    MOR_TRY(upload_vec(iv, inv_vn));
    MOR_TRY(upload_vec(ib, inv_both));
    MOR_TRY(h->Vs.alloc(nn));
    MOR_TRY(h->M.alloc(nn));
    if (accumulate_v) {
        MOR_TRY(gather_cols(V.as<double>(), n, n, n, perm.as<int>(), inv_vn.as<double>(), h->Vs.as<double>(), n));
    }
    if (on_transpose) {
        // columns of A are v_j s_j:  V = A S^-1,  M = V S^-1 = A S^-2
        MOR_TRY(gather_cols(A.as<double>(), n, n, n, perm.as<int>(), inv_sigma.as<double>(), h->Vs.as<double>(), n));
        MOR_TRY(gather_cols(A.as<double>(), n, n, n, perm.as<int>(), inv_both.as<double>(), h->M.as<double>(), n));
    } else if (single_pass && accumulate_v) {
        // T = U S V'  =>  U = T V S^-1
        MOR_TRY(gather_cols(V.as<double>(), n, n, n, perm.as<int>(), inv_both.as<double>(), h->M.as<double>(), n));
    } else {
        // T_final (X itself after one pass, Q_{p-1} otherwise) times inv(R_last) is orthonormal:
        // U = T_final * inv(R_last) * U_R with U_R = A S^-1
        MOR_TRY(gather_cols(A.as<double>(), n, n, n, perm.as<int>(), inv_sigma.as<double>(), G.as<double>(), n));
        MOR_TRY(small_trinv_upper(Rs.as<double>(), n, n, Rinv.as<double>(), n));
        MOR_TRY(small_scale_rc(Rinv.as<double>(), n, n, n, D.as<double>(), /*rows=*/true, /*inv=*/true));
        MOR_TRY(small_gemm(false, false, n, n, n, 1.0, Rinv.as<double>(), n, G.as<double>(), n, 0.0, h->M.as<double>(), n));
        if (!accumulate_v) {  // V = R' U_R S^-1
            MOR_TRY(small_gemm(true, false, n, n, n, 1.0, Rtot.as<double>(), n, G.as<double>(), n, 0.0, h->Vs.as<double>(), n));
            MOR_TRY(small_scale_rc(h->Vs.as<double>(), n, n, n, inv_sigma.as<double>(), /*rows=*/false, /*inv=*/false));
        }
    }
    dbg.lap("basis maker M, right vectors");
    MOR_TRY(sign_fix(h->Vs.as<double>(), n, n, n, h->M.as<double>(), n, n));
    MOR_CUDA(cudaStreamSynchronize(st));  // perm / inv_sigma are released on return
    tm.core.stop();
    h->T = T;
    h->trows = trows;
    h->ldt = ldt;
    h->tn = n;
    h->vs_rows = n;
    h->r = n;
    h->passes = passes;
    h->sweeps = sweeps;
    return MOR_OK;
}

// Randomized SVD (reference POD.jl:22-27 -> RandomizedLinAlg.rsvd(X, k, p)).
int32_t factor_rsvd(mor_pod_s* h, const double* X, int64_t m, int n, int64_t ldx, int k, int p, const double* Omega_d,
                    int64_t ldo, uint64_t seed, Timers& tm) {
    cudaStream_t st = ctx().stream;
    int l = k + p;
    if ((int64_t)l > h->m_global) l = (int)h->m_global;  // thin QR of an m x l sketch has min(m, l) columns
    if (l > n) l = n;
    DevBuf Om;
    if (!Omega_d) {
        MOR_TRY(Om.alloc(sizeof(double) * (size_t)n * (k + p)));
        MOR_TRY(fill_normal(Om.as<double>(), n, k + p, n, seed, 0, 1.0));
        Omega_d = Om.as<double>();
        ldo = n;
    }
    // Y = X * Omega(:, 1:l)
    const int64_t ldy = ((m + 1) & ~int64_t(1)) > 0 ? ((m + 1) & ~int64_t(1)) : 2;
    MOR_TRY(h->T_owned.alloc(sizeof(double) * (size_t)ldy * (size_t)l));
    double* Y = h->T_owned.as<double>();
    MOR_CUDA(cudaMemsetAsync(Y, 0, h->T_owned.bytes, st));
    tm.sketch.start();
    MOR_TRY(gemm_nn(X, ldx, n, Omega_d, (int)ldo, l, m, Y, ldy, 0));
    tm.sketch.stop();
    int passes = 0;
    MOR_TRY(orthonormalize(Y, m, l, ldy, false, tm, &passes));
    const int kq = (p > 0) ? std::min(k, l) : l;  // rsvd keeps Q[:, 1:k] when p > 0
    // B' = X' Q  (n x kq), replicated after the all-reduce
    DevBuf Bt, V;
    MOR_TRY(Bt.alloc(sizeof(double) * (size_t)n * kq));
    tm.sketch.start();
    MOR_TRY(gemm_tn(X, ldx, n, Y, ldy, kq, m, Bt.as<double>(), n));
    tm.sketch.stop();
    if (ctx().nranks > 1) {
        tm.comm.start();
        MOR_TRY(comm_allreduce_sum(Bt.as<double>(), (size_t)n * kq));
        tm.comm.stop();
    }
    // B = U_B S V_B'  <=>  B' J = V_B S with J = U_B: one-sided Jacobi on the columns of B'
    tm.core.start();
    MOR_TRY(V.alloc(sizeof(double) * (size_t)kq * kq));
    std::vector<int> order;
    std::vector<double> sig;
    int sweeps = 0;
    int32_t js = jacobi_svd(Bt.as<double>(), n, kq, n, V.as<double>(), kq, sig, order, &sweeps);
    if (js != MOR_OK) {
        tm.core.stop();
        return js;
    }
    const int r = std::min(k, kq);
    h->sigma.assign(sig.begin(), sig.begin() + r);
    DevBuf perm, inv_sigma;
    MOR_TRY(upload_perm(order, perm));
    std::vector<double> is((size_t)kq);
    for (int j = 0; j < kq; ++j) is[j] = sig[j] > 0.0 ? 1.0 / sig[j] : 0.0;
    MOR_TRY(upload_vec(is, inv_sigma));
    MOR_TRY(h->M.alloc(sizeof(double) * (size_t)kq * r));
    MOR_TRY(h->Vs.alloc(sizeof(double) * (size_t)n * r));
    std::vector<double> vn;
    MOR_TRY(col_norms(V.as<double>(), kq, kq, kq, vn));
    std::vector<double> iv((size_t)kq);
    for (int j = 0; j < kq; ++j) iv[j] = vn[order[j]] > 0.0 ? 1.0 / vn[order[j]] : 1.0;
    DevBuf inv_vn;
    MOR_TRY(upload_vec(iv, inv_vn));
    MOR_TRY(gather_cols(V.as<double>(), kq, kq, r, perm.as<int>(), inv_vn.as<double>(), h->M.as<double>(), kq));
    MOR_TRY(gather_cols(Bt.as<double>(), n, n, r, perm.as<int>(), inv_sigma.as<double>(), h->Vs.as<double>(), n));
    MOR_TRY(sign_fix(h->Vs.as<double>(), n, n, r, h->M.as<double>(), kq, kq));
    MOR_CUDA(cudaStreamSynchronize(st));
    tm.core.stop();
    h->T = Y;
    h->trows = m;
    h->ldt = ldy;
    h->tn = kq;
    h->vs_rows = n;
    h->r = r;
    h->passes = passes;
    h->sweeps = sweeps;
    return MOR_OK;
}

int32_t global_rows(int64_t m, int64_t* m_global) {
    std::vector<int64_t> all;
    MOR_TRY(comm_allgather_i64(m, all));
    int64_t s = 0;
    for (int64_t v : all) s += v;
    *m_global = s;
    return MOR_OK;
}

// Common entry: X is a device matrix (this rank's shard).  `own` != nullptr transfers
// ownership of X's allocation to the handle (host-pointer entry points).
int32_t pod_factor_device(double* X, int64_t m, int64_t n, int64_t ldx, DevBuf* own, int32_t method, int64_t k,
                          int64_t p, const double* Omega_d, int64_t ldo, uint64_t seed, int32_t flags,
                          mor_pod_t* handle, double* S_out, int64_t* nS, Timers& tm, const double* G_first = nullptr) {
    MOR_ARG(handle && S_out && nS, "pod_factor: null output pointer");
    MOR_ARG(m >= 0 && n >= 1, "pod_factor: bad dimensions");
    MOR_ARG(method == MOR_SVD || method == MOR_TSVD || method == MOR_RSVD, "pod_factor: unknown method");
    MOR_ARG(n <= 4096 || m <= 4096, "pod_factor: the small dimension of the snapshot matrix must be <= 4096");
    int64_t m_global = 0;
    MOR_TRY(global_rows(m, &m_global));
    MOR_ARG(m_global >= 1, "pod_factor: empty snapshot matrix");
    const int64_t rmax = std::min(m_global, n);
    if (method != MOR_SVD) MOR_ARG(k >= 1 && k <= rmax, "pod_factor: nmodes must be in 1..min(m, n)");
    MOR_ARG(p >= 0, "pod_factor: negative oversampling");
    const bool wide = m_global < n;
    MOR_ARG(!(wide && ctx().nranks > 1), "pod_factor: a wide snapshot matrix (m < n) cannot be row-sharded");

    std::unique_ptr<mor_pod_s> h(new mor_pod_s);
    h->m = m;
    h->n = n;
    h->m_global = m_global;
    h->method = method;
    h->wide = false;
    if (own) h->T_owned.take(*own);
    const bool may_overwrite = own != nullptr || (flags & MOR_POD_MAY_OVERWRITE);
    const bool force2 = (flags & MOR_POD_FORCE_TWO_PASS) != 0;

    if (method == MOR_RSVD) {
        DevBuf keepX;  // X must stay alive through the call only (Q is a separate matrix)
        if (own) keepX.take(h->T_owned);
        MOR_TRY(factor_rsvd(h.get(), X, m, (int)n, ldx, (int)k, (int)p, Omega_d, ldo, seed, tm));
        MOR_CUDA(cudaStreamSynchronize(ctx().stream));
    } else if (!wide) {
        MOR_TRY(factor_tall(h.get(), X, m, (int)n, ldx, may_overwrite, force2, false, tm, G_first));
    } else {
        // T = X' (n x m), owned by the handle
        const int64_t ldt = (n + 1) & ~int64_t(1);
        DevBuf Tt;
        MOR_TRY(Tt.alloc(sizeof(double) * (size_t)ldt * (size_t)m));
        MOR_CUDA(cudaMemsetAsync(Tt.p, 0, Tt.bytes, ctx().stream));
        MOR_TRY(transpose(X, m, n, ldx, Tt.as<double>(), ldt));
        MOR_CUDA(cudaStreamSynchronize(ctx().stream));
        h->T_owned.take(Tt);   // the uploaded X (if any) is no longer needed
        h->wide = true;
        MOR_TRY(factor_tall(h.get(), h->T_owned.as<double>(), n, (int)m, ldt, true, force2, true, tm));
    }
    const int64_t count = (method == MOR_SVD) ? rmax : k;
    for (int64_t i = 0; i < count; ++i) S_out[i] = i < (int64_t)h->sigma.size() ? h->sigma[(size_t)i] : 0.0;
    *nS = count;
    *handle = h.release();
    return MOR_OK;
}

// tall-side vectors: out (trows x k, ldo) = T * M(:, 0:k)
int32_t tall_vectors(mor_pod_s* h, int k, double* out, int64_t ldo, Timers& tm) {
    tm.basis.start();
    MOR_TRY(gemm_nn(h->T, h->ldt, h->tn, h->M.as<double>(), h->tn, k, h->trows, out, ldo, 0));
    tm.basis.stop();
    return MOR_OK;
}

void reset_timings() {
    for (double& t : ctx().timings) t = 0.0;
}

}  // namespace

int32_t orthonormalize_device(double* A, int64_t m, int n, int64_t ld, int* passes) {
    reset_timings();
    Timers tm;
    tm.total.start();
    int32_t st = orthonormalize(A, m, n, ld, false, tm, passes);
    tm.total.stop();
    cudaStreamSynchronize(ctx().stream);
    tm.collect();
    return st;
}

}  // namespace mor

using namespace mor;

extern "C" {

int32_t mor_b200_pod_factor_dev(mor_mat_t X, int32_t method, int64_t k, int64_t p, mor_mat_t Omega, uint64_t seed,
                                int32_t flags, mor_pod_t* handle, double* S_out, int64_t* nS) {
    MOR_API_LOCK;
    MOR_ROUTE_SOLO(mor_b200_pod_factor_dev(X, method, k, p, Omega, seed, flags, handle, S_out, nS));
    MOR_ARG(X != nullptr, "mor_b200_pod_factor_dev: null matrix");
    MOR_TRY(require_ctx());
    MOR_ARG(X->ld % 2 == 0 && (uintptr_t)X->p % 16 == 0,
            "mor_b200_pod_factor_dev: matrix must be 16-byte aligned with an even leading dimension");
    if (Omega) MOR_ARG(Omega->m == X->n && Omega->n >= k + p, "mor_b200_pod_factor_dev: Omega must be n x (k+p)");
    reset_timings();
    Timers tm;
    tm.total.start();
    int32_t st = pod_factor_device(X->p, X->m, X->n, X->ld, nullptr, method, k, p, Omega ? Omega->p : nullptr,
                                   Omega ? Omega->ld : 0, seed, flags, handle, S_out, nS, tm);
    tm.total.stop();
    cudaStreamSynchronize(ctx().stream);
    tm.collect();
    return st;
}

static int32_t factor_from_host(const double* X, const double* const* cols, int64_t m, int64_t n, int64_t ldx,
                                int32_t method, int64_t k, int64_t p, const double* Omega, uint64_t seed,
                                mor_pod_t* handle, double* S_out, int64_t* nS) {
    MOR_ARG((X != nullptr || cols != nullptr) && handle && S_out && nS, "mor_b200_pod_factor: null pointer");
    MOR_ARG(m >= 0 && n >= 1 && (cols || ldx >= std::max<int64_t>(m, 1)), "mor_b200_pod_factor: bad dimensions");
    MOR_TRY(require_ctx());
    Ctx& cx = ctx();
    reset_timings();
    Timers tm;
    tm.total.start();
    const int64_t ld = std::max<int64_t>((m + 1) & ~int64_t(1), 2);
    DevBuf dX, dOm, Gacc;
    MOR_TRY(ingest_alloc(dX, sizeof(double) * (size_t)ld * (size_t)n));
    if (ld != m)  // keep the pad row finite
        MOR_CUDA(cudaMemset2DAsync(dX.as<double>() + m, (size_t)ld * 8, 0, 8, (size_t)n, cx.stream));
    // Pipelined ingest (tall SVD/TSVD of a Matrix): rows travel in ~1 GiB chunks on a copy
    // stream while the Gram matrix of the chunks that already landed accumulates on the DMMA
    // pipe -- the first CholeskyQR pass is hidden behind the PCIe transfer.
    const char* ch_env = getenv("MOR_B200_INGEST_CHUNK_ROWS");  // test hook: force small chunks
    const bool pipelined = !cols && method != MOR_RSVD && m >= n && n <= 4096 &&
                           (ch_env != nullptr || (size_t)m * n >= (size_t(1) << 24));
    // On every exit path (errors included): nothing may still be reading the caller's buffer or writing into dX
    // when they go away, and the events are released.
    struct IngestGuard {
        Ctx& cx;
        std::vector<cudaEvent_t> events;
        ~IngestGuard() {
            if (cx.copy_stream) cudaStreamSynchronize(cx.copy_stream);
            cudaStreamSynchronize(cx.stream);
            for (cudaEvent_t ev : events) cudaEventDestroy(ev);
        }
    } guard{cx, {}};
    std::vector<cudaEvent_t>& events = guard.events;
    if (m > 0 && pipelined) {
        if (!cx.copy_stream) MOR_CUDA(cudaStreamCreateWithFlags(&cx.copy_stream, cudaStreamNonBlocking));
        MOR_TRY(Gacc.alloc(sizeof(double) * (size_t)n * n));
        cudaEvent_t ready;
        MOR_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        MOR_CUDA(cudaEventRecord(ready, cx.stream));
        MOR_CUDA(cudaStreamWaitEvent(cx.copy_stream, ready, 0));
        events.push_back(ready);
        int64_t ch = ((int64_t(1) << 27) / n) & ~int64_t(4095);  // rows per chunk: ~1 GiB
        if (ch < 4096) ch = 4096;
        if (ch_env && atoll(ch_env) > 0) ch = (atoll(ch_env) + 1) & ~1LL;
        tm.gram.start();
        for (int64_t r0 = 0; r0 < m; r0 += ch) {
            const int64_t rc = std::min(ch, m - r0);
            MOR_TRY(h2d_2d(dX.as<double>() + r0, ld, X + r0, ldx, rc, n, cx.copy_stream));
            cudaEvent_t ev;
            MOR_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            MOR_CUDA(cudaEventRecord(ev, cx.copy_stream));
            MOR_CUDA(cudaStreamWaitEvent(cx.stream, ev, 0));
            events.push_back(ev);
            MOR_TRY(gemm_tn(dX.as<double>() + r0, ld, (int)n, dX.as<double>() + r0, ld, (int)n, rc, Gacc.as<double>(),
                            (int)n, r0 > 0 ? 1 : 0));
        }
        tm.gram.stop();
    } else if (m > 0) {
        tm.h2d.start();
        if (cols) {
            for (int64_t j = 0; j < n; ++j) {
                MOR_ARG(cols[j] != nullptr, "mor_b200_pod_factor_cols: null column pointer");
                MOR_TRY(h2d_2d(dX.as<double>() + (size_t)j * ld, ld, cols[j], m, m, 1, cx.stream));
            }
        } else {
            MOR_TRY(h2d_2d(dX.as<double>(), ld, X, ldx, m, n, cx.stream));
        }
        tm.h2d.stop();
    }
    int64_t ldo = 0;
    if (Omega && method == MOR_RSVD) {
        MOR_TRY(dOm.alloc(sizeof(double) * (size_t)n * (size_t)(k + p)));
        MOR_CUDA(cudaMemcpyAsync(dOm.p, Omega, sizeof(double) * (size_t)n * (size_t)(k + p), cudaMemcpyHostToDevice,
                                 cx.stream));
        ldo = n;
    }
    int32_t st = pod_factor_device(dX.as<double>(), m, n, ld, &dX, method, k, p, dOm.as<double>(), ldo, seed, 0, handle,
                                   S_out, nS, tm, (m > 0 && pipelined) ? Gacc.as<double>() : nullptr);
    tm.total.stop();
    cudaStreamSynchronize(cx.stream);
    tm.collect();
    return st;
}

int32_t mor_b200_pod_factor(const double* X, int64_t m, int64_t n, int64_t ldx, int32_t method, int64_t k, int64_t p,
                            const double* Omega, uint64_t seed, mor_pod_t* handle, double* S_out, int64_t* nS) {
    MOR_API_LOCK;
    if (multi_on() && !is_worker()) return multi_pod_factor(X, nullptr, m, n, ldx, method, k, p, Omega, seed, handle, S_out, nS);
    return factor_from_host(X, nullptr, m, n, ldx, method, k, p, Omega, seed, handle, S_out, nS);
}

int32_t mor_b200_pod_factor_cols(const double* const* cols, int64_t m, int64_t n, int32_t method, int64_t k,
                                 int64_t p, const double* Omega, uint64_t seed, mor_pod_t* handle, double* S_out,
                                 int64_t* nS) {
    MOR_API_LOCK;
    if (multi_on() && !is_worker()) return multi_pod_factor(nullptr, cols, m, n, m, method, k, p, Omega, seed, handle, S_out, nS);
    MOR_ARG(cols != nullptr, "mor_b200_pod_factor_cols: null pointer");
    return factor_from_host(nullptr, cols, m, n, m, method, k, p, Omega, seed, handle, S_out, nS);
}

int32_t mor_b200_pod_basis_dev(mor_pod_t h, int64_t k, mor_mat_t U_out) {
    MOR_API_LOCK;
    MOR_ROUTE_SOLO(mor_b200_pod_basis_dev(h, k, U_out));
    MOR_ARG(h != nullptr && U_out != nullptr, "mor_b200_pod_basis_dev: null argument");
    MOR_ARG(k >= 1 && k <= h->r, "mor_b200_pod_basis_dev: k must be in 1..number of computed triplets");
    MOR_ARG(U_out->m == h->m && U_out->n >= k, "mor_b200_pod_basis_dev: U_out must be m x >= k");
    MOR_TRY(require_ctx());
    reset_timings();
    Timers tm;
    tm.total.start();
    int32_t st = MOR_OK;
    if (!h->wide) {
        st = tall_vectors(h, (int)k, U_out->p, U_out->ld, tm);
    } else {  // U is the small side: m x k block of Vs
        cudaError_t e = cudaMemcpy2DAsync(U_out->p, (size_t)U_out->ld * 8, h->Vs.p, (size_t)h->vs_rows * 8,
                                          (size_t)h->m * 8, (size_t)k, cudaMemcpyDeviceToDevice, ctx().stream);
        if (e != cudaSuccess) st = cuda_fail(e, "cudaMemcpy2DAsync", __FILE__, __LINE__);
    }
    tm.total.stop();
    cudaStreamSynchronize(ctx().stream);
    tm.collect();
    return st;
}

int32_t mor_b200_pod_right_dev(mor_pod_t h, int64_t k, mor_mat_t V_out) {
    MOR_API_LOCK;
    MOR_ROUTE_SOLO(mor_b200_pod_right_dev(h, k, V_out));
    MOR_ARG(h != nullptr && V_out != nullptr, "mor_b200_pod_right_dev: null argument");
    MOR_ARG(k >= 1 && k <= h->r, "mor_b200_pod_right_dev: k must be in 1..number of computed triplets");
    MOR_ARG(V_out->m == h->n && V_out->n >= k, "mor_b200_pod_right_dev: V_out must be n x >= k");
    MOR_TRY(require_ctx());
    reset_timings();
    Timers tm;
    int32_t st = MOR_OK;
    if (h->wide) {  // the right vectors of a wide X are the tall side of its transpose
        st = tall_vectors(h, (int)k, V_out->p, V_out->ld, tm);
    } else {
        cudaError_t e = cudaMemcpy2DAsync(V_out->p, (size_t)V_out->ld * 8, h->Vs.p, (size_t)h->vs_rows * 8,
                                          (size_t)h->vs_rows * 8, (size_t)k, cudaMemcpyDeviceToDevice, ctx().stream);
        if (e != cudaSuccess) st = cuda_fail(e, "cudaMemcpy2DAsync", __FILE__, __LINE__);
    }
    cudaStreamSynchronize(ctx().stream);
    tm.collect();
    return st;
}

int32_t mor_b200_pod_basis(mor_pod_t h, int64_t k, double* U_out, int64_t ldu, double* V_out, int64_t ldv) {
    MOR_API_LOCK;
    if (h != nullptr && multi_on() && !is_worker()) return multi_pod_basis(h, k, U_out, ldu, V_out, ldv);
    MOR_ARG(h != nullptr && U_out != nullptr, "mor_b200_pod_basis: null argument");
    MOR_ARG(k >= 1 && k <= h->r, "mor_b200_pod_basis: k must be in 1..number of computed triplets");
    MOR_ARG(ldu >= std::max<int64_t>(h->m, 1) && (!V_out || ldv >= h->n), "mor_b200_pod_basis: leading dimension too small");
    MOR_TRY(require_ctx());
    Ctx& cx = ctx();
    reset_timings();
    Timers tm;
    tm.total.start();
    // tall-side vectors (trows x k) computed on the device, then copied out
    double* tall_host = h->wide ? V_out : U_out;
    const int64_t tall_ld = h->wide ? ldv : ldu;
    double* small_host = h->wide ? U_out : V_out;
    const int64_t small_ld = h->wide ? ldu : ldv;
    if (tall_host && h->trows > 0) {
        // U = T * M in blocks of 16 columns through two device buffers: block c+1 is computed while block c travels
        // to the host on the copy stream (the whole m x k result need not fit next to a 137 GB snapshot matrix, and
        // its 8.6 GB no longer cost a cudaMalloc / cudaFree pair per call)
        const int64_t ldd = std::max<int64_t>((h->trows + 1) & ~int64_t(1), 2);
        constexpr int CB = 16;
        const int nblk = (int)((k + CB - 1) / CB);
        DevBuf dU[2];
        for (int b = 0; b < 2 && b < nblk; ++b) MOR_TRY(dU[b].alloc(sizeof(double) * (size_t)ldd * (size_t)std::min<int64_t>(CB, k)));
        if (!cx.copy_stream) MOR_CUDA(cudaStreamCreateWithFlags(&cx.copy_stream, cudaStreamNonBlocking));
        struct Ev {
            cudaEvent_t done[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
            ~Ev() {
                for (int b = 0; b < 2; ++b) {
                    if (done[b]) cudaEventDestroy(done[b]);
                    if (copied[b]) cudaEventDestroy(copied[b]);
                }
            }
        } ev;
        for (int b = 0; b < 2; ++b) {
            MOR_CUDA(cudaEventCreateWithFlags(&ev.done[b], cudaEventDisableTiming));
            MOR_CUDA(cudaEventCreateWithFlags(&ev.copied[b], cudaEventDisableTiming));
        }
        auto compute = [&](int blk) -> int32_t {
            const int b = blk & 1, c0 = blk * CB, nc = (int)std::min<int64_t>(CB, k - c0);
            if (blk >= 2) MOR_CUDA(cudaStreamWaitEvent(cx.stream, ev.copied[b], 0));   // the buffer's previous block has left
            tm.basis.start();
            MOR_TRY(gemm_nn(h->T, h->ldt, h->tn, h->M.as<double>() + (size_t)c0 * h->tn, h->tn, nc, h->trows, dU[b].as<double>(), ldd, 0));
            tm.basis.stop();
            MOR_CUDA(cudaEventRecord(ev.done[b], cx.stream));
            return MOR_OK;
        };
        MOR_TRY(compute(0));
        for (int blk = 0; blk < nblk; ++blk) {
            if (blk + 1 < nblk) MOR_TRY(compute(blk + 1));
            const int b = blk & 1, c0 = blk * CB, nc = (int)std::min<int64_t>(CB, k - c0);
            MOR_CUDA(cudaStreamWaitEvent(cx.copy_stream, ev.done[b], 0));
            MOR_TRY(d2h_2d(tall_host + (size_t)c0 * tall_ld, tall_ld, dU[b].as<double>(), ldd, h->trows, nc, cx.copy_stream, /*sync=*/false));
            MOR_CUDA(cudaEventRecord(ev.copied[b], cx.copy_stream));
        }
        MOR_CUDA(cudaStreamSynchronize(cx.copy_stream));
        MOR_CUDA(cudaStreamSynchronize(cx.stream));
    }
    if (small_host) {
        tm.d2h.start();
        MOR_CUDA(cudaMemcpy2DAsync(small_host, (size_t)small_ld * 8, h->Vs.p, (size_t)h->vs_rows * 8,
                                   (size_t)h->vs_rows * 8, (size_t)k, cudaMemcpyDeviceToHost, cx.stream));
        tm.d2h.stop();
        MOR_CUDA(cudaStreamSynchronize(cx.stream));
    }
    tm.total.stop();
    cudaStreamSynchronize(cx.stream);
    tm.collect();
    return MOR_OK;
}

int32_t mor_b200_pod_free(mor_pod_t h) {
    MOR_API_LOCK;
    if (h != nullptr && multi_on() && !is_worker()) return multi_pod_free(h);
    if (!h) return MOR_OK;
    if (ctx().ready) {
        cudaSetDevice(ctx().device);
        cudaStreamSynchronize(ctx().stream);
    }
    delete h;
    return MOR_OK;
}

int32_t mor_b200_pod_info(mor_pod_t h, int32_t* passes, int32_t* sweeps) {
    MOR_ARG(h != nullptr, "mor_b200_pod_info: null handle");
    if (!h->parts.empty()) h = h->parts[0];
    if (passes) *passes = h->passes;
    if (sweeps) *sweeps = h->sweeps;
    return MOR_OK;
}

int32_t mor_b200_orthonormalize_dev(mor_mat_t A, int32_t* passes) {
    MOR_API_LOCK;
    MOR_ROUTE_SOLO(mor_b200_orthonormalize_dev(A, passes));
    MOR_ARG(A != nullptr, "mor_b200_orthonormalize_dev: null matrix");
    MOR_TRY(require_ctx());
    MOR_ARG(A->n >= 1 && A->n <= 4096, "mor_b200_orthonormalize_dev: column count must be in 1..4096");
    int p = 0;
    int32_t st = orthonormalize_device(A->p, A->m, (int)A->n, A->ld, &p);
    if (passes) *passes = p;
    return st;
}

}  // extern "C"
