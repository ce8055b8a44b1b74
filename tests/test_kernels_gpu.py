"""Kernel-level checks of the hand-written POD kernels against numpy (fp64): the DMMA
tall-skinny GEMMs, the blocked Cholesky, the triangular inverse and the Jacobi SVD.
These go through test hooks (mor_dbg_*) that call the kernels exactly as the POD driver does."""
import ctypes as C

import numpy as np
import pytest

from mor_b200 import _lib
from mor_b200.device import DeviceMatrix

pytestmark = pytest.mark.gpu


def _dbg(lib, name, *args):
    fn = getattr(_lib.load_test(), name)        # libmor_b200_test.so: the hooks are not in the product library
    fn.restype = C.c_int32
    _lib.check(fn(*args))


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("m,p,q", [(1, 1, 1), (7, 3, 5), (20, 64, 64), (1000, 10, 10), (4099, 130, 70),
                                   (50_000, 256, 256), (33_333, 300, 64), (20_001, 64, 512), (8192, 1024, 1024)])
def test_gemm_tn_general(gpu, m, p, q):
    rng = np.random.default_rng(m + p + q)
    A = rng.standard_normal((m, p)); B = rng.standard_normal((m, q))
    dA, dB, dC = DeviceMatrix.from_host(A), DeviceMatrix.from_host(B), DeviceMatrix.alloc(p, q)
    _dbg(gpu, "mor_dbg_gemm_tn", dA.handle, dB.handle, dC.handle)
    got = dC.to_host()
    assert rel(got, A.T @ B) < 1e-13


@pytest.mark.parametrize("m,n", [(2, 2), (50, 7), (3000, 64), (10_000, 65), (100_000, 256), (30_011, 384), (9000, 1024)])
def test_gemm_tn_gram_is_symmetric_and_exact(gpu, m, n):
    rng = np.random.default_rng(m * 7 + n)
    X = rng.standard_normal((m, n)) * np.logspace(0, -3, n)
    dX, dG = DeviceMatrix.from_host(X), DeviceMatrix.alloc(n, n)
    _dbg(gpu, "mor_dbg_gemm_tn", dX.handle, dX.handle, dG.handle)
    G = dG.to_host()
    assert np.array_equal(G, G.T)                     # mirrored, bit-exact symmetry
    ref = X.T @ X
    d = np.sqrt(np.diag(ref))
    assert np.max(np.abs(G - ref) / np.outer(d, d)) < 1e-13   # componentwise in the column scaling


@pytest.mark.parametrize("m,n,q", [(1, 1, 1), (5, 3, 2), (128, 16, 64), (1000, 20, 7), (4097, 100, 64), (20_000, 256, 32),
                                   (3001, 1024, 64), (12_345, 70, 130)])
def test_gemm_nn(gpu, m, n, q):
    rng = np.random.default_rng(m + 3 * n + q)
    X = rng.standard_normal((m, n)); W = rng.standard_normal((n, q))
    dX, dW, dY = DeviceMatrix.from_host(X), DeviceMatrix.from_host(W), DeviceMatrix.alloc(m, q)
    _dbg(gpu, "mor_dbg_gemm_nn", dX.handle, dW.handle, dY.handle, 0)
    assert rel(dY.to_host(), X @ W) < 1e-13


@pytest.mark.parametrize("m,n", [(3, 2), (1000, 64), (5000, 65), (4099, 200), (2000, 1024)])
def test_gemm_nn_in_place_triangular(gpu, m, n):
    rng = np.random.default_rng(m + n)
    X = rng.standard_normal((m, n))
    W = np.triu(rng.standard_normal((n, n))) + 3 * np.eye(n)
    Wdirty = W + np.tril(np.full((n, n), np.nan), -1)      # the kernel must ignore the lower triangle
    dX, dW = DeviceMatrix.from_host(X), DeviceMatrix.from_host(Wdirty)
    _dbg(gpu, "mor_dbg_gemm_nn", dX.handle, dW.handle, dX.handle, 1)
    assert rel(dX.to_host(), X @ W) < 1e-13


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 100, 256, 1000])
def test_cholesky_and_triangular_inverse(gpu, n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n + 5, n))
    G = A.T @ A + 0.1 * np.eye(n)
    dG = DeviceMatrix.from_host(G)
    fail = C.c_int32(-1)
    _dbg(gpu, "mor_dbg_chol", dG.handle, C.byref(fail))
    assert fail.value == 0
    R = dG.to_host()
    assert np.array_equal(R, np.triu(R))
    assert rel(R.T @ R, G) < 1e-13
    dRi = DeviceMatrix.alloc(n, n)
    _dbg(gpu, "mor_dbg_trinv", dG.handle, dRi.handle)
    Ri = dRi.to_host()
    assert np.max(np.abs(Ri @ R - np.eye(n))) < 1e-10
    # breakdown is reported, not hidden
    Gbad = G.copy(); Gbad[n // 2, n // 2] = -1.0
    dB = DeviceMatrix.from_host(Gbad)
    _dbg(gpu, "mor_dbg_chol", dB.handle, C.byref(fail))
    assert fail.value == n // 2 + 1


@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_small_gemm(gpu, ta, tb):
    rng = np.random.default_rng(5)
    m, n, k = 70, 45, 100
    A = rng.standard_normal((k, m) if ta else (m, k)); B = rng.standard_normal((n, k) if tb else (k, n))
    dA, dB, dC = DeviceMatrix.from_host(A), DeviceMatrix.from_host(B), DeviceMatrix.alloc(m, n)
    _dbg(gpu, "mor_dbg_small_gemm", ta, tb, dA.handle, dB.handle, dC.handle)
    ref = (A.T if ta else A) @ (B.T if tb else B)
    assert rel(dC.to_host(), ref) < 1e-14


@pytest.mark.parametrize("rows,n,kappa", [(2, 2, 1e0), (5, 5, 1e3), (64, 64, 1e6), (200, 33, 1e10), (256, 256, 1e8),
                                          (1000, 7, 1e2)])
def test_jacobi_svd(gpu, rows, n, kappa):
    rng = np.random.default_rng(rows + n)
    U, _ = np.linalg.qr(rng.standard_normal((rows, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    s = np.logspace(0, -np.log10(kappa), n) if n > 1 else np.ones(1)
    A = (U * s) @ V.T
    dA, dV = DeviceMatrix.from_host(A), DeviceMatrix.alloc(n, n)
    sig = np.zeros(n); order = np.zeros(n, dtype=np.int32); sweeps = C.c_int32(0)
    _dbg(gpu, "mor_dbg_jacobi", dA.handle, dV.handle, sig.ctypes.data, order.ctypes.data, C.byref(sweeps))
    s_ref = np.linalg.svd(A, compute_uv=False)
    assert np.max(np.abs(sig - s_ref) / s_ref[0]) < 4e-13
    # one-sided Jacobi resolves small singular values to high RELATIVE accuracy when the
    # matrix is a well-conditioned basis times a column scaling; here just check against s
    assert np.max(np.abs(sig - s) / s) < 1e-6 * kappa * 1e-8 + 1e-9
    Vh = dV.to_host(); Ah = dA.to_host()
    assert np.max(np.abs(Vh.T @ Vh - np.eye(n))) < 2e-12
    assert rel(A @ Vh, Ah) < 1e-13
    assert 1 <= sweeps.value <= 40   # no de Rijk / Drmac-Veselic preconditioning yet: generic kappa=1e8 input takes ~30 sweeps


@pytest.mark.parametrize("m,l0,bw", [(1, 16, 16), (127, 16, 3), (128, 32, 16), (129, 48, 16), (4099, 240, 16),
                                     (50_001, 112, 9), (300_000, 64, 16)])
@pytest.mark.parametrize("tile", [128, 256, 512, 25616])
def test_deim_panel_residual(gpu, monkeypatch, m, l0, bw, tile):
    """deim_panel_kernel: R = U[:, l0:l0+bw] - U[:, :l0] C (the residual panel of the blocked DEIM path),
    for every tile shape of the kernel (rows per tile x columns per stage = 128x16, 256x8, 512x4, default 256x16)."""
    monkeypatch.setenv("MOR_B200_PANEL_TILE", str(tile))
    rng = np.random.default_rng(m + l0 + bw)
    U = rng.standard_normal((m, l0 + bw)); Cm = rng.standard_normal((l0, bw))
    dU, dC, dR = DeviceMatrix.from_host(U), DeviceMatrix.from_host(Cm), DeviceMatrix.alloc(m, 16)
    _dbg(gpu, "mor_dbg_deim_panel", dU.handle, l0, bw, dC.handle, dR.handle)
    got = dR.to_host()[:, :bw]
    ref = U[:, l0:] - U[:, :l0] @ Cm
    assert np.max(np.abs(got - ref)) < 1e-12 * np.sqrt(l0)
