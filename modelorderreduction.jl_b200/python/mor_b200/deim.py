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
    strided_ok = (m <= 1 or basis.strides[0] == 8) and (
        k <= 1 or (basis.strides[1] % 8 == 0 and basis.strides[1] // 8 >= m))
    B = basis if strided_ok else np.asfortranarray(basis)
    ldb = B.strides[1] // 8 if k > 1 else max(m, 1)
    idx = np.empty(k, dtype=np.int64)
    _lib.check(_lib.load().mor_b200_deim_indices(B.ctypes.data, m, k, max(ldb, 1), idx.ctypes.data))
    return idx
