// Synthetic snapshot generators: counter-based Philox4x32-10, so the
// value of entry (global row, column) depends only on the seed -- the same matrix is
// produced for any number of row shards.  HBM write-bound.
#include <algorithm>
#include <cmath>
#include <vector>

#include "internal.h"

namespace mor {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

// four N(0,1) variates from one Philox block (Box-Muller in fp32; the column scale applied
// afterwards is a full FP64 number, so entries carry 53-bit mantissas)
__device__ __forceinline__ void normal4(uint4 r, float z[4]) {
    const float inv32 = 2.3283064365386963e-10f;  // 2^-32
    float u0 = ((float)r.x + 0.5f) * inv32, u1 = ((float)r.y + 0.5f) * inv32;
    float u2 = ((float)r.z + 0.5f) * inv32, u3 = ((float)r.w + 0.5f) * inv32;
    u0 = fminf(fmaxf(u0, 1e-12f), 1.0f);
    u2 = fminf(fmaxf(u2, 1e-12f), 1.0f);
    float ra = sqrtf(-2.0f * __logf(u0)), rb = sqrtf(-2.0f * __logf(u2));
    float sa, ca, sb, cb;
    __sincosf(6.283185307179586f * u1, &sa, &ca);
    __sincosf(6.283185307179586f * u3, &sb, &cb);
    z[0] = ra * ca;
    z[1] = ra * sa;
    z[2] = rb * cb;
    z[3] = rb * sb;
}

// one thread = one (block of 4 global rows, column)
__global__ void __launch_bounds__(256)
synth_kernel(double* __restrict__ X, int64_t m, int64_t n, int64_t ld, uint64_t seed, int64_t row0,
             const double* __restrict__ colscale) {
    const int64_t gb0 = row0 >> 2;                         // first global 4-row block touched
    const int64_t nblk = ((row0 + m + 3) >> 2) - gb0;      // blocks overlapping [row0, row0+m)
    const int64_t total = nblk * n;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t col = t / nblk;
        const int64_t gb = gb0 + (t - col * nblk);
        uint4 r = philox4x32_10(make_uint4((uint32_t)gb, (uint32_t)(gb >> 32), (uint32_t)col, 0x6d6f72u), key);
        float z[4];
        normal4(r, z);
        const double s = colscale[col];
        double* dst = X + col * ld;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t lr = (gb << 2) + q - row0;  // local row
            if (lr >= 0 && lr < m) dst[lr] = (double)z[q] * s;
        }
    }
}

static int32_t launch_synth(double* X, int64_t m, int64_t n, int64_t ld, uint64_t seed, int64_t row0,
                            const std::vector<double>& scale) {
    if (m == 0 || n == 0) return MOR_OK;
    Ctx& c = ctx();
    DevBuf ds;
    MOR_TRY(ds.alloc(sizeof(double) * (size_t)n));
    MOR_CUDA(cudaMemcpyAsync(ds.p, scale.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c.stream));
    int64_t total = (((row0 + m + 3) >> 2) - (row0 >> 2)) * n;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)c.num_sms * 32;
    if (blocks > cap) blocks = cap;
    synth_kernel<<<(unsigned)blocks, 256, 0, c.stream>>>(X, m, n, ld, seed, row0, ds.as<double>());
    count_launch();
    MOR_CUDA(cudaGetLastError());
    MOR_CUDA(cudaStreamSynchronize(c.stream));  // ds is released on return
    return MOR_OK;
}

int32_t kat_fill(double* X, int64_t m, int64_t n, int64_t ld, uint64_t seed, int64_t row0, double decades);

// ---- 2-D viscous Burgers snapshots ---------------------------
// Cole-Hopf closed form: with phi = sum_i exp(theta_i),
//   theta_i = (-(p_i x + q_i y) + (p_i^2 + q_i^2) t / 2 + d_i) / (2 nu)      (phi_t = nu lap(phi)),
// (u, v) = -2 nu grad(phi) / phi = softmax(theta)-weighted mean of the states (p_i, q_i) solves
//   u_t + u u_x + v u_y = nu lap(u)  (same for v):
// three constant states separated by travelling fronts of width ~ 4 nu / |p_i - p_j| that merge in a
// moving triple point.  Row r of the snapshot matrix is grid node (ix, iy) = (r mod nx, r div nx) of the
// u component, x = (ix + 1/2) / nx, y = (iy + 1/2) / nx; column j is time t_j = j / (n - 1).
// The same closed form is restated in oracle/mor_oracle.py:burgers_snapshots.
struct BurgersCoef {
    double ax[3], ay[3], p[3];
};

static void burgers_states(uint64_t seed, double p[3], double q[3], double d[3]) {
    // irrational front directions (correctly rounded square roots, the same doubles as in the oracle): with
    // few-digit decimals the front normals are commensurate with a 4096-wide grid, distant nodes along a front
    // then see histories equal to the last bit or two and the DEIM argmax is decided by rounding noise
    const double P[3] = {sqrt(0.878), sqrt(0.0008), -sqrt(0.5805)}, Q[3] = {sqrt(0.0633), -sqrt(0.2641), sqrt(0.2379)};
    const double D[3] = {0.0, -0.45, -0.55};
    for (int i = 0; i < 3; ++i) {
        p[i] = P[i];
        q[i] = Q[i];
        // the seed shifts the initial front positions by at most +-0.01 (multiples of 1/3200)
        const uint64_t h = (seed * (uint64_t)(2 * i + 3)) & 63u;
        d[i] = D[i] + ((double)h - 32.0) / 3200.0;
    }
}

// one thread = one row, blockIdx.y = a group of 16 columns; tcoef[3 * j + i] = time part of theta_i in column j
__global__ void __launch_bounds__(256)
burgers_kernel(double* __restrict__ X, int64_t m, int64_t n, int64_t ld, int64_t row0, int64_t nx, BurgersCoef cf,
               const double* __restrict__ tcoef) {
    const int64_t j0 = (int64_t)blockIdx.y * 16, j1 = j0 + 16 < n ? j0 + 16 : n;
    const double h = 1.0 / (double)nx;
    for (int64_t lr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; lr < m; lr += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = row0 + lr, iy = g / nx, ix = g - iy * nx;
        const double x = ((double)ix + 0.5) * h, y = ((double)iy + 0.5) * h;
        double s[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) s[i] = cf.ax[i] * x + cf.ay[i] * y;
        for (int64_t j = j0; j < j1; ++j) {
            double th[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) th[i] = s[i] + __ldg(&tcoef[3 * j + i]);
            const double mx = fmax(th[0], fmax(th[1], th[2]));
            double num = 0.0, den = 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double e = exp(th[i] - mx);
                num += cf.p[i] * e;
                den += e;
            }
            X[j * ld + lr] = num / den;
        }
    }
}

static int32_t burgers_fill(double* X, int64_t m, int64_t n, int64_t ld, uint64_t seed, int64_t row0, double nu,
                            int64_t nx) {
    MOR_ARG(nu > 0.0 && nx >= 1, "mor_b200_synth(BURGERS): param = viscosity > 0, iparam = grid width >= 1");
    if (m == 0 || n == 0) return MOR_OK;
    Ctx& c = ctx();
    double p[3], q[3], d[3];
    burgers_states(seed, p, q, d);
    BurgersCoef cf;
    std::vector<double> tc((size_t)(3 * n));
    for (int i = 0; i < 3; ++i) {
        cf.ax[i] = -p[i] / (2.0 * nu);
        cf.ay[i] = -q[i] / (2.0 * nu);
        cf.p[i] = p[i];
    }
    for (int64_t j = 0; j < n; ++j) {
        const double t = n > 1 ? (double)j / (double)(n - 1) : 0.0;
        for (int i = 0; i < 3; ++i) tc[(size_t)(3 * j + i)] = (0.5 * (p[i] * p[i] + q[i] * q[i]) * t + d[i]) / (2.0 * nu);
    }
    DevBuf dt;
    MOR_TRY(dt.alloc(sizeof(double) * tc.size()));
    MOR_CUDA(cudaMemcpyAsync(dt.p, tc.data(), sizeof(double) * tc.size(), cudaMemcpyHostToDevice, c.stream));
    int64_t bx = (m + 255) / 256;
    const int64_t cap = (int64_t)c.num_sms * 8;
    if (bx > cap) bx = cap;
    dim3 grid((unsigned)bx, (unsigned)((n + 15) / 16));
    burgers_kernel<<<grid, 256, 0, c.stream>>>(X, m, n, ld, row0, nx, cf, dt.as<double>());
    count_launch();
    MOR_CUDA(cudaGetLastError());
    MOR_CUDA(cudaStreamSynchronize(c.stream));  // dt and tc are released on return
    return MOR_OK;
}

int32_t synth_fill(double* X, int64_t m, int64_t n, int64_t ld, int32_t kind, uint64_t seed,
                   int64_t row0, double param, int64_t iparam) {
    std::vector<double> scale((size_t)(n > 0 ? n : 1), 1.0);
    switch (kind) {
        case MOR_SYNTH_GRADED:
            for (int64_t j = 0; j < n; ++j)
                scale[j] = n > 1 ? std::pow(10.0, -param * (double)j / (double)(n - 1)) : 1.0;
            break;
        case MOR_SYNTH_GAUSS:
            for (int64_t j = 0; j < n; ++j) scale[j] = param;
            break;
        case MOR_SYNTH_LOWRANK:
            for (int64_t j = 0; j < n; ++j) {
                if (j < iparam) scale[j] = 1.0;
                else {
                    int64_t rest = n - iparam;
                    scale[j] = 1e-3 * (rest > 1 ? std::pow(10.0, -param * (double)(j - iparam) / (double)(rest - 1)) : 1.0);
                }
            }
            break;
        case MOR_SYNTH_KAT:
            return kat_fill(X, m, n, ld, seed, row0, param);
        case MOR_SYNTH_BURGERS:
            return burgers_fill(X, m, n, ld, seed, row0, param, iparam);
        default:
            set_error("mor_b200_synth: unknown generator kind");
            return MOR_ERR_ARG;
    }
    return launch_synth(X, m, n, ld, seed, row0, scale);
}

// ---- known-answer matrix (SURVEY.md section 8d "KAT at full size") -------------------------------
//   X = H_u * [S V'; 0],  H_u = I - 2 u u' (u = z / ||z||, z Philox N(0,1) over the M global rows),
//   V' = I - 2 v v' (v_j ~ cos(j + 1), normalised),  S = diag(10^(-param j / (n - 1))).
// Its singular values are exactly S (up to the rounding of the entries), its left vectors the
// columns of H_u, its right vectors the columns of V -- at any size and for any row sharding.
// Not a column scaling of a well-conditioned matrix: it exercises the multi-pass CholeskyQR path.
__device__ __forceinline__ double kat_z(uint64_t seed, int64_t grow) {
    const int64_t gb = grow >> 2;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    uint4 r = philox4x32_10(make_uint4((uint32_t)gb, (uint32_t)(gb >> 32), 0xffffffffu, 0x6b6174u), key);
    float z[4];
    normal4(r, z);
    return (double)z[grow & 3];
}

__global__ void __launch_bounds__(256)
kat_sumsq_kernel(uint64_t seed, int64_t row0, int64_t m, double* __restrict__ out) {
    __shared__ double red[8];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const double z = kat_z(seed, row0 + i);
        s = fma(z, z, s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(out, t);
    }
}

__global__ void kat_top_kernel(uint64_t seed, int n, double* __restrict__ ztop) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ztop[i] = kat_z(seed, i);
}

// reference left vector entry: (H_u e_j)[gi] = delta - 2 u_gi u_j
__device__ __forceinline__ double kat_uref(uint64_t seed, int64_t gi, int j, double inv_norm, const double* utop) {
    return (gi == j ? 1.0 : 0.0) - 2.0 * (kat_z(seed, gi) * inv_norm) * utop[j];
}

__global__ void __launch_bounds__(256)
kat_fill_kernel(double* __restrict__ X, int64_t m, int64_t n, int64_t ld, uint64_t seed, int64_t row0, double inv_norm,
                const double* __restrict__ sig, const double* __restrict__ v, const double* __restrict__ w) {
    const int64_t total = m * n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t col = t / m, lr = t - col * m, gi = row0 + lr;
        double b = 0.0;
        if (gi < n) b = sig[gi] * ((gi == col ? 1.0 : 0.0) - 2.0 * v[gi] * v[col]);
        X[col * ld + lr] = b - 2.0 * (kat_z(seed, gi) * inv_norm) * w[col];
    }
}

// d[j] += sum_i U[i][j] * uref(i, j)   (pass 0);   e[j] = max_i |U[i][j] - sgn[j] * uref(i, j)|  (pass 1)
__global__ void __launch_bounds__(256)
kat_check_kernel(const double* __restrict__ U, int64_t m, int64_t ld, uint64_t seed, int64_t row0, double inv_norm,
                 const double* __restrict__ utop, int pass, const double* __restrict__ sgn, double* __restrict__ out) {
    __shared__ double red[8];
    const int j = blockIdx.y;
    const double* u = U + (size_t)j * ld;
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const double r = kat_uref(seed, row0 + i, j, inv_norm, utop);
        if (pass == 0) s = fma(u[i], r, s);
        else s = fmax(s, fabs(u[i] - sgn[j] * r));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, s, off);
        s = pass == 0 ? s + o : fmax(s, o);
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = red[0];
        for (int w = 1; w < 8; ++w) t = pass == 0 ? t + red[w] : fmax(t, red[w]);
        if (pass == 0) atomicAdd(&out[j], t);
        else atomicMax(reinterpret_cast<unsigned long long*>(&out[j]), (unsigned long long)__double_as_longlong(t));
    }
}

struct KatSetup {
    double inv_norm;
    std::vector<double> utop, sig, v, w;
};

// collective over the ranks (one all-reduce of ||z||^2)
static int32_t kat_setup(int64_t m_local, int64_t row0, int64_t n, uint64_t seed, double decades, KatSetup* ks) {
    Ctx& c = ctx();
    DevBuf acc, ztop;
    MOR_TRY(acc.alloc(sizeof(double)));
    MOR_TRY(ztop.alloc(sizeof(double) * (size_t)n));
    MOR_CUDA(cudaMemsetAsync(acc.p, 0, sizeof(double), c.stream));
    if (m_local > 0) kat_sumsq_kernel<<<c.num_sms * 8, 256, 0, c.stream>>>(seed, row0, m_local, acc.as<double>());
    kat_top_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(seed, (int)n, ztop.as<double>());
    count_launch(2);
    MOR_CUDA(cudaGetLastError());
    MOR_TRY(comm_allreduce_sum(acc.as<double>(), 1));
    double zz = 0.0;
    std::vector<double> z((size_t)n);
    MOR_CUDA(cudaMemcpyAsync(&zz, acc.p, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    MOR_CUDA(cudaMemcpyAsync(z.data(), ztop.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c.stream));
    MOR_CUDA(cudaStreamSynchronize(c.stream));
    MOR_ARG(zz > 0.0, "known-answer matrix: empty");
    ks->inv_norm = 1.0 / std::sqrt(zz);
    ks->utop.resize((size_t)n);
    ks->sig.resize((size_t)n);
    ks->v.resize((size_t)n);
    ks->w.resize((size_t)n);
    double vn = 0.0;
    for (int64_t j = 0; j < n; ++j) {
        ks->utop[j] = z[j] * ks->inv_norm;
        ks->sig[j] = n > 1 ? std::pow(10.0, -decades * (double)j / (double)(n - 1)) : 1.0;
        ks->v[j] = std::cos((double)(j + 1));
        vn += ks->v[j] * ks->v[j];
    }
    vn = 1.0 / std::sqrt(vn);
    double t = 0.0;
    for (int64_t j = 0; j < n; ++j) {
        ks->v[j] *= vn;
        t += ks->sig[j] * ks->utop[j] * ks->v[j];
    }
    for (int64_t j = 0; j < n; ++j) ks->w[j] = ks->sig[j] * ks->utop[j] - 2.0 * ks->v[j] * t;
    return MOR_OK;
}

static int32_t upload(const std::vector<double>& h, DevBuf& d) {
    MOR_TRY(d.alloc(sizeof(double) * h.size()));
    MOR_CUDA(cudaMemcpyAsync(d.p, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, ctx().stream));
    return MOR_OK;
}

int32_t kat_fill(double* X, int64_t m, int64_t n, int64_t ld, uint64_t seed, int64_t row0, double decades) {
    Ctx& c = ctx();
    KatSetup ks;
    MOR_TRY(kat_setup(m, row0, n, seed, decades, &ks));
    DevBuf ds, dv, dw;
    MOR_TRY(upload(ks.sig, ds));
    MOR_TRY(upload(ks.v, dv));
    MOR_TRY(upload(ks.w, dw));
    if (m > 0 && n > 0) {
        kat_fill_kernel<<<c.num_sms * 16, 256, 0, c.stream>>>(X, m, n, ld, seed, row0, ks.inv_norm, ds.as<double>(),
                                                             dv.as<double>(), dw.as<double>());
        count_launch();
        MOR_CUDA(cudaGetLastError());
    }
    MOR_CUDA(cudaStreamSynchronize(c.stream));
    return MOR_OK;
}

// max_i,j<k |U[i][j] -+ (H_u e_j)[i]| over ALL ranks (collective); U is this rank's row shard
int32_t kat_check_left(const double* U, int64_t m, int64_t ld, int k, int64_t n, uint64_t seed, int64_t row0,
                       double decades, double* max_err) {
    Ctx& c = ctx();
    MOR_ARG(k >= 1 && k <= n, "kat_check: k must be in 1..n");
    KatSetup ks;
    MOR_TRY(kat_setup(m, row0, n, seed, decades, &ks));
    DevBuf du, dd, de;
    MOR_TRY(upload(ks.utop, du));
    MOR_TRY(dd.alloc(sizeof(double) * (size_t)k));
    MOR_TRY(de.alloc(sizeof(double) * (size_t)k));
    MOR_CUDA(cudaMemsetAsync(dd.p, 0, sizeof(double) * (size_t)k, c.stream));
    MOR_CUDA(cudaMemsetAsync(de.p, 0, sizeof(double) * (size_t)k, c.stream));
    dim3 grid(c.num_sms, (unsigned)k);
    if (m > 0) kat_check_kernel<<<grid, 256, 0, c.stream>>>(U, m, ld, seed, row0, ks.inv_norm, du.as<double>(), 0, nullptr, dd.as<double>());
    MOR_TRY(comm_allreduce_sum(dd.as<double>(), (size_t)k));
    std::vector<double> d((size_t)k);
    MOR_CUDA(cudaMemcpyAsync(d.data(), dd.p, sizeof(double) * (size_t)k, cudaMemcpyDeviceToHost, c.stream));
    MOR_CUDA(cudaStreamSynchronize(c.stream));
    for (double& x : d) x = x < 0.0 ? -1.0 : 1.0;
    MOR_CUDA(cudaMemcpyAsync(dd.p, d.data(), sizeof(double) * (size_t)k, cudaMemcpyHostToDevice, c.stream));
    if (m > 0) kat_check_kernel<<<grid, 256, 0, c.stream>>>(U, m, ld, seed, row0, ks.inv_norm, du.as<double>(), 1, dd.as<double>(), de.as<double>());
    count_launch(2);
    MOR_CUDA(cudaGetLastError());
    std::vector<double> e((size_t)k);
    MOR_CUDA(cudaMemcpyAsync(e.data(), de.p, sizeof(double) * (size_t)k, cudaMemcpyDeviceToHost, c.stream));
    MOR_CUDA(cudaStreamSynchronize(c.stream));
    double mx = 0.0;
    for (double x : e) mx = std::max(mx, x);
    // max over ranks through a sum all-reduce of a one-hot vector (NCCL has MAX, but this keeps one collective type)
    DevBuf slots;
    MOR_TRY(slots.alloc(sizeof(double) * (size_t)c.nranks));
    std::vector<double> one((size_t)c.nranks, 0.0);
    one[(size_t)c.rank] = mx;
    MOR_CUDA(cudaMemcpyAsync(slots.p, one.data(), sizeof(double) * (size_t)c.nranks, cudaMemcpyHostToDevice, c.stream));
    MOR_TRY(comm_allreduce_sum(slots.as<double>(), (size_t)c.nranks));
    MOR_CUDA(cudaMemcpyAsync(one.data(), slots.p, sizeof(double) * (size_t)c.nranks, cudaMemcpyDeviceToHost, c.stream));
    MOR_CUDA(cudaStreamSynchronize(c.stream));
    for (double x : one) mx = std::max(mx, x);
    *max_err = mx;
    return MOR_OK;
}

int32_t fill_normal(double* X, int64_t m, int64_t n, int64_t ld, uint64_t seed, int64_t row0, double scale) {
    std::vector<double> s((size_t)(n > 0 ? n : 1), scale);
    return launch_synth(X, m, n, ld, seed, row0, s);
}

}  // namespace mor
