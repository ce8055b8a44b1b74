// Kernel-level test hooks (NOT part of the public ABI in include/mor_b200.h): they let
// tests/test_kernels_gpu.py compare each hand-written kernel with numpy on its own.
#include "internal.h"

using namespace mor;

extern "C" {

// R (m x bw) = U[:, l0 : l0+bw] - U[:, 0 : l0] * Cm  (Cm: l0 x bw on the device) through deim_panel_kernel
int32_t mor_dbg_deim_panel(mor_mat_t U, int32_t l0, int32_t bw, mor_mat_t Cm, mor_mat_t R) {
    MOR_API_LOCK;
    MOR_ARG(U && Cm && R && l0 >= 16 && l0 % 16 == 0 && bw >= 1 && bw <= 16 && U->n >= l0 + bw && Cm->m == l0 &&
                Cm->n == bw && R->m == U->m && R->n >= bw,
            "mor_dbg_deim_panel: shapes");
    MOR_TRY(require_ctx());
    std::vector<double> ch((size_t)Cm->ld * bw), cp((size_t)(l0 / 16) * 320, 0.0);
    MOR_CUDA(cudaMemcpy(ch.data(), Cm->p, sizeof(double) * ch.size(), cudaMemcpyDeviceToHost));
    for (int c = 0; c < l0 / 16; ++c)
        for (int q = 0; q < bw; ++q)
            for (int kk = 0; kk < 16; ++kk) cp[(size_t)c * 320 + q * 20 + kk] = ch[(size_t)q * Cm->ld + 16 * c + kk];
    DevBuf d;
    MOR_TRY(d.alloc(sizeof(double) * cp.size()));
    MOR_CUDA(cudaMemcpyAsync(d.p, cp.data(), sizeof(double) * cp.size(), cudaMemcpyHostToDevice, ctx().stream));
    MOR_TRY(deim_panel(U->p, U->ld, l0, bw, d.as<double>(), U->m, R->p, R->ld));
    MOR_CUDA(cudaStreamSynchronize(ctx().stream));
    return MOR_OK;
}

int32_t mor_dbg_gemm_tn(mor_mat_t A, mor_mat_t B, mor_mat_t C) {
    MOR_API_LOCK;
    MOR_ARG(A && B && C && A->m == B->m && C->m == A->n && C->n == B->n, "mor_dbg_gemm_tn: shapes");
    MOR_TRY(require_ctx());
    MOR_TRY(gemm_tn(A->p, A->ld, (int)A->n, B->p, B->ld, (int)B->n, A->m, C->p, (int)C->ld));
    MOR_CUDA(cudaStreamSynchronize(ctx().stream));
    return MOR_OK;
}

int32_t mor_dbg_gemm_nn(mor_mat_t X, mor_mat_t W, mor_mat_t Y, int32_t upper) {
    MOR_API_LOCK;
    MOR_ARG(X && W && Y && W->m == X->n && Y->m == X->m && Y->n == W->n, "mor_dbg_gemm_nn: shapes");
    MOR_TRY(require_ctx());
    MOR_TRY(gemm_nn(X->p, X->ld, (int)X->n, W->p, (int)W->ld, (int)W->n, X->m, Y->p, Y->ld, upper));
    MOR_CUDA(cudaStreamSynchronize(ctx().stream));
    return MOR_OK;
}

// in place: G (n x n SPD) -> R upper with R'R = G; returns the failing pivot (1-based) or 0
int32_t mor_dbg_chol(mor_mat_t G, int32_t* fail) {
    MOR_API_LOCK;
    MOR_ARG(G && G->m == G->n && fail, "mor_dbg_chol: shapes");
    MOR_TRY(require_ctx());
    DevBuf info;
    MOR_TRY(info.alloc(sizeof(SmallInfo)));
    MOR_TRY(small_info_reset(info.as<SmallInfo>()));
    MOR_TRY(small_chol_upper(G->p, (int)G->n, (int)G->ld, info.as<SmallInfo>()));
    SmallInfo h;
    MOR_TRY(small_info_read(info.as<SmallInfo>(), &h));
    *fail = h.chol_fail;
    return MOR_OK;
}

int32_t mor_dbg_trinv(mor_mat_t R, mor_mat_t Rinv) {
    MOR_API_LOCK;
    MOR_ARG(R && Rinv && R->m == R->n && Rinv->m == R->m && Rinv->n == R->n, "mor_dbg_trinv: shapes");
    MOR_TRY(require_ctx());
    MOR_TRY(small_trinv_upper(R->p, (int)R->n, (int)R->ld, Rinv->p, (int)Rinv->ld));
    MOR_CUDA(cudaStreamSynchronize(ctx().stream));
    return MOR_OK;
}

int32_t mor_dbg_small_gemm(int32_t ta, int32_t tb, mor_mat_t A, mor_mat_t B, mor_mat_t C) {
    MOR_API_LOCK;
    MOR_ARG(A && B && C, "mor_dbg_small_gemm: null");
    MOR_TRY(require_ctx());
    const int k = (int)(ta ? A->m : A->n);
    MOR_TRY(small_gemm(ta != 0, tb != 0, (int)C->m, (int)C->n, k, 1.0, A->p, (int)A->ld, B->p, (int)B->ld, 0.0, C->p,
                       (int)C->ld));
    MOR_CUDA(cudaStreamSynchronize(ctx().stream));
    return MOR_OK;
}

// in place one-sided Jacobi on A (rows x n); V (n x n) accumulated; sigma/order length n
int32_t mor_dbg_jacobi(mor_mat_t A, mor_mat_t V, double* sigma, int32_t* order, int32_t* sweeps) {
    MOR_API_LOCK;
    MOR_ARG(A && V && sigma && order && V->m == A->n && V->n == A->n, "mor_dbg_jacobi: shapes");
    MOR_TRY(require_ctx());
    std::vector<double> s;
    std::vector<int> o;
    int sw = 0;
    MOR_TRY(jacobi_svd(A->p, (int)A->m, (int)A->n, (int)A->ld, V->p, (int)V->ld, s, o, &sw));
    for (size_t i = 0; i < s.size(); ++i) {
        sigma[i] = s[i];
        order[i] = o[i];
    }
    if (sweeps) *sweeps = sw;
    MOR_CUDA(cudaStreamSynchronize(ctx().stream));
    return MOR_OK;
}

}  // extern "C"

// ---- live roofline denominators for bench.py (MEASURED_PEAKS.json has no FP64 figure) ------
namespace {
__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, int iters, double a0, double b0) {
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc[i][0] = threadIdx.x;
        acc[i][1] = i;
    }
    const double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[i][0]), "+d"(acc[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) peak_read_kernel(const double2* __restrict__ p, size_t n2, double* out) {
    double s = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t st = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * st < n2; i += 4 * st) {
        const double2 a = p[i], b = p[i + st], c = p[i + 2 * st], d = p[i + 3 * st];
        s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
    }
    for (; i < n2; i += st) {
        const double2 a = p[i];
        s += a.x + a.y;
    }
    if (s == 1.2345e300) out[0] = s;
}
}  // namespace

extern "C" {

// best-of-`reps` DMMA.8x8x4 issue-rate peak (TFLOP/s) and streaming read bandwidth (GB/s over
// a `gib`-GiB buffer), measured with CUDA events on the library stream
int32_t mor_dbg_peaks(int32_t reps, int32_t gib, double* dmma_tflops, double* read_gbs) {
    MOR_API_LOCK;
    MOR_ARG(dmma_tflops && read_gbs && reps >= 1 && gib >= 1, "mor_dbg_peaks: bad arguments");
    MOR_TRY(require_ctx());
    Ctx& cx = ctx();
    DevBuf out, buf;
    MOR_TRY(out.alloc(sizeof(double) * (size_t)cx.num_sms * 4 * 256));
    const size_t bytes = (size_t)gib << 30;
    MOR_TRY(buf.alloc(bytes));
    MOR_CUDA(cudaMemsetAsync(buf.p, 0, bytes, cx.stream));
    cudaEvent_t a, b;
    MOR_CUDA(cudaEventCreate(&a));
    MOR_CUDA(cudaEventCreate(&b));
    const int iters = 8192;
    double best_t = 0.0, best_g = 0.0;
    for (int r = 0; r < reps + 1; ++r) {
        float ms = 0.f;
        MOR_CUDA(cudaEventRecord(a, cx.stream));
        peak_dmma_kernel<<<cx.num_sms * 4, 256, 0, cx.stream>>>(out.as<double>(), iters, 1.0, 1e-3);
        MOR_CUDA(cudaEventRecord(b, cx.stream));
        MOR_CUDA(cudaEventSynchronize(b));
        MOR_CUDA(cudaEventElapsedTime(&ms, a, b));
        const double tf = (double)cx.num_sms * 4 * 8 * iters * 8.0 * 512.0 / (ms * 1e-3) * 1e-12;
        if (r > 0 && tf > best_t) best_t = tf;
        MOR_CUDA(cudaEventRecord(a, cx.stream));
        peak_read_kernel<<<cx.num_sms * 16, 256, 0, cx.stream>>>(buf.as<double2>(), bytes / 16, out.as<double>());
        MOR_CUDA(cudaEventRecord(b, cx.stream));
        MOR_CUDA(cudaEventSynchronize(b));
        MOR_CUDA(cudaEventElapsedTime(&ms, a, b));
        const double gbs = (double)bytes / (ms * 1e-3) * 1e-9;
        if (r > 0 && gbs > best_g) best_g = gbs;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *dmma_tflops = best_t;
    *read_gbs = best_g;
    return MOR_OK;
}

}  // extern "C"
