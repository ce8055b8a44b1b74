// POD driver -- placeholder entry points until the factorisation kernels land.
#include "internal.h"
using namespace mor;
extern "C" {
#define NOTYET(name) set_error(name ": not implemented yet"); return MOR_ERR_CUDA;
int32_t mor_b200_pod_factor(const double*, int64_t, int64_t, int64_t, int32_t, int64_t, int64_t, const double*, uint64_t, mor_pod_t*, double*, int64_t*) { NOTYET("pod_factor") }
int32_t mor_b200_pod_factor_cols(const double* const*, int64_t, int64_t, int32_t, int64_t, int64_t, const double*, uint64_t, mor_pod_t*, double*, int64_t*) { NOTYET("pod_factor_cols") }
int32_t mor_b200_pod_basis(mor_pod_t, int64_t, double*, int64_t, double*, int64_t) { NOTYET("pod_basis") }
int32_t mor_b200_pod_free(mor_pod_t) { return MOR_OK; }
int32_t mor_b200_pod_info(mor_pod_t, int32_t*, int32_t*) { NOTYET("pod_info") }
int32_t mor_b200_pod_factor_dev(mor_mat_t, int32_t, int64_t, int64_t, mor_mat_t, uint64_t, int32_t, mor_pod_t*, double*, int64_t*) { NOTYET("pod_factor_dev") }
int32_t mor_b200_pod_basis_dev(mor_pod_t, int64_t, mor_mat_t) { NOTYET("pod_basis_dev") }
int32_t mor_b200_orthonormalize_dev(mor_mat_t, int32_t*) { NOTYET("orthonormalize_dev") }
}
