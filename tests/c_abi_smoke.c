/* Plain-C consumer of include/mor_b200.h: proves the boundary is a C ABI (no C++ / torch types) and
 * that the library fails loudly without a device.  Built and run by tests/test_abi_symbols.py:
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude tests/c_abi_smoke.c -o c_abi_smoke \
 *       -Lmodelorderreduction.jl_b200/lib -lmor_b200 -Wl,-rpath,<libdir> -lm
 * Exit code: 0 = POD + DEIM ran on a GPU and passed the checks below; 3 = no CUDA device (the
 * library reported MOR_ERR_CUDA, as it must: there is no CPU fallback); anything else = failure. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "mor_b200.h"

int main(void) {
    const int64_t m = 2000, n = 6, k = 2;
    double* X = (double*)malloc(sizeof(double) * (size_t)(m * n));
    double S[6], *U = (double*)malloc(sizeof(double) * (size_t)(m * k));
    int64_t nS = 0, idx[2] = {0, 0};
    mor_pod_t h = 0;
    int32_t st;
    int64_t i, j;
    if (mor_b200_version() != MOR_B200_VERSION) return 10;
    /* two exactly orthogonal dominant directions with known singular values 3 sqrt(m/2) and 1.5 sqrt(m/2) */
    for (j = 0; j < n; ++j)
        for (i = 0; i < m; ++i) X[j * m + i] = 0.0;
    for (i = 0; i < m; ++i) {
        X[0 * m + i] = (i % 2 == 0) ? 3.0 : 0.0;
        X[1 * m + i] = (i % 2 == 1) ? 1.5 : 0.0;
        X[2 * m + i] = 1e-3 * sin(0.37 * (double)i);
    }
    st = mor_b200_init(0);
    if (st == MOR_ERR_CUDA) {
        printf("no CUDA device: %s\n", mor_b200_last_error());
        st = mor_b200_pod_factor(X, m, n, m, MOR_SVD, k, 0, 0, 0, &h, S, &nS);
        return st == MOR_ERR_CUDA ? 3 : 11;
    }
    if (st != MOR_OK) return 12;
    st = mor_b200_pod_factor(X, m, n, m, MOR_SVD, k, 0, 0, 0, &h, S, &nS);
    if (st != MOR_OK) { printf("pod_factor: %s\n", mor_b200_last_error()); return 13; }
    if (nS != n) return 14;
    if (fabs(S[0] - 3.0 * sqrt(m / 2.0)) > 1e-6 * S[0] || fabs(S[1] - 1.5 * sqrt(m / 2.0)) > 1e-6 * S[1]) return 15;
    st = mor_b200_pod_basis(h, k, U, m, 0, 0);
    if (st != MOR_OK) return 16;
    mor_b200_pod_free(h);
    st = mor_b200_deim_indices(U, m, k, m, idx);
    if (st != MOR_OK) { printf("deim: %s\n", mor_b200_last_error()); return 17; }
    if (idx[0] < 1 || idx[0] > m || idx[1] < 1 || idx[1] > m || idx[0] == idx[1]) return 18;
    if ((idx[0] % 2) == (idx[1] % 2)) return 19; /* one point on each of the two interleaved supports */
    printf("c_abi_smoke ok: s = %.6f %.6f, DEIM indices %lld %lld\n", S[0], S[1], (long long)idx[0], (long long)idx[1]);
    mor_b200_shutdown();
    free(X);
    free(U);
    return 0;
}
