"""`deim_interpolation_indices(basis)` -- /root/reference/src/deim.jl:9-28 -- rerouted to
mor_b200_deim_indices.  Pure function of `basis` (m x k dense Float64, "should not be a
sparse matrix", deim.jl:6); returns k 1-based Int64 indices like the Julia function."""
from __future__ import annotations

import numpy as np

from . import _lib


def deim_interpolation_indices(basis: np.ndarray) -> np.ndarray:
    if not isinstance(basis, np.ndarray) or basis.ndim != 2:
        raise TypeError("basis must be a dense 2-D array (deim.jl:6)")
    if basis.dtype != np.float64:
        raise TypeError("the B200 path handles Float64 (the shim defers other types to the reference)")
    m, k = basis.shape
    if k > 1024:
        from .pod import DeferToReference
        raise DeferToReference(f"{k} basis vectors > 1024 (the Julia shim runs the reference's deim_interpolation_indices)")
    strided_ok = (m <= 1 or basis.strides[0] == 8) and (
        k <= 1 or (basis.strides[1] % 8 == 0 and basis.strides[1] // 8 >= m))
    B = basis if strided_ok else np.asfortranarray(basis)
    ldb = B.strides[1] // 8 if k > 1 else max(m, 1)
    idx = np.empty(k, dtype=np.int64)
    _lib.check(_lib.load().mor_b200_deim_indices(B.ctypes.data, m, k, max(ldb, 1), idx.ctypes.data))
    return idx


def deim_last_margins(with_values: bool = False):
    """(margins, min_margin, min_step) of the last `deim_interpolation_indices` call: margins[l] is the relative gap
    between the best and the second-best residual at greedy step l+1 (mor_b200_deim_last_margins); the minimum is
    over steps >= 2.  `with_values`: also the winning residuals.  Not in the reference: it tells a caller how far
    above FP64 rounding each choice was."""
    import ctypes as C
    lib = _lib.load()
    n = C.c_int64(0)
    _lib.check(lib.mor_b200_deim_last_margins(None, None, 0, C.byref(n), None, None))
    m, v = np.zeros(max(n.value, 1)), np.zeros(max(n.value, 1))
    mn, at = C.c_double(1.0), C.c_int64(0)
    pd = C.POINTER(C.c_double)
    _lib.check(lib.mor_b200_deim_last_margins(m.ctypes.data_as(pd), v.ctypes.data_as(pd), n.value, None, C.byref(mn), C.byref(at)))
    if with_values:
        return m[:n.value], mn.value, at.value, v[:n.value]
    return m[:n.value], mn.value, at.value


def deim_last_margin() -> float:
    return deim_last_margins()[1]


def deim_projector(U: np.ndarray, V: np.ndarray, indices: np.ndarray) -> np.ndarray:
    """`((@view U[indices, :])' \\ (U' * V))'` -- the DEIM projector assembled inside the numeric
    `deim(...)` at /root/reference/src/deim.jl:150, rerouted to mor_b200_deim_projector.  U is the
    n x m DEIM basis, V the n x k POD basis, `indices` the 1-based interpolation indices; returns
    the k x m projector with F(y) ~= projector * F(V * y_hat)[indices]."""
    U = np.asfortranarray(U, dtype=np.float64)
    V = np.asfortranarray(V, dtype=np.float64)
    if U.ndim != 2 or V.ndim != 2 or U.shape[0] != V.shape[0]:
        raise ValueError("U and V must be matrices with the same number of rows")
    idx = np.ascontiguousarray(indices, dtype=np.int64)
    if idx.shape != (U.shape[1],):
        raise ValueError("need one interpolation index per DEIM basis vector")
    n, kd = U.shape
    kp = V.shape[1]
    P = np.empty((kp, kd), dtype=np.float64, order="F")
    _lib.check(_lib.load().mor_b200_deim_projector(U.ctypes.data, n, kd, max(n, 1), V.ctypes.data, kp, max(n, 1),
                                                   idx.ctypes.data, P.ctypes.data, max(kp, 1)))
    return P


def galerkin_project(A, V: np.ndarray) -> np.ndarray:
    """`A_hat = V' * A * V` (/root/reference/src/deim.jl:154) for the sparse linear operator A that
    `separate_terms` returns (utils.jl:126).  `A` is a scipy.sparse matrix (any format; converted to
    CSC = Julia's SparseMatrixCSC layout and handed over with 1-based indices)."""
    import scipy.sparse as sp
    A = sp.csc_matrix(A)
    V = np.asfortranarray(V, dtype=np.float64)
    m, k = V.shape
    if A.shape != (m, m):
        raise ValueError("A must be square with as many rows as V")
    colptr = (A.indptr.astype(np.int64) + 1)
    rowval = (A.indices.astype(np.int64) + 1)
    nzval = np.ascontiguousarray(A.data, dtype=np.float64)
    out = np.empty((k, k), dtype=np.float64, order="F")
    _lib.check(_lib.load().mor_b200_galerkin_csc(m, colptr.ctypes.data, rowval.ctypes.data, nzval.ctypes.data,
                                                 V.ctypes.data, k, max(m, 1), out.ctypes.data, max(k, 1)))
    return out


def project(V: np.ndarray, x: np.ndarray) -> np.ndarray:
    """`V' * x` for a vector or the columns of a matrix: `g_hat = V' * g` (deim.jl:155) and the reduced
    initial condition `V' * snapshot[:, 1]` (deim.jl:261)."""
    V = np.asfortranarray(V, dtype=np.float64)
    xv = np.asfortranarray(np.asarray(x, dtype=np.float64).reshape(V.shape[0], -1))
    m, k = V.shape
    nvec = xv.shape[1]
    out = np.empty((k, nvec), dtype=np.float64, order="F")
    _lib.check(_lib.load().mor_b200_project(V.ctypes.data, m, k, max(m, 1), xv.ctypes.data, nvec, max(m, 1),
                                            out.ctypes.data, max(k, 1)))
    return out[:, 0].copy() if np.ndim(x) == 1 else out
