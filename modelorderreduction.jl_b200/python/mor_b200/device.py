"""Device-resident side of the C ABI (mor_b200_mat_*, *_dev): inputs already in HBM.
Used by bench.py and by GPU-side callers; mirrors nothing in the reference (which has no
device path) -- it is the measurement surface for `value` in the bench contract."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib


class DeviceMatrix:
    """Column-major FP64 matrix in HBM owned (or borrowed) by libmor_b200."""

    def __init__(self, handle, owned: bool = True):
        self.handle = handle
        self._owned = owned

    @classmethod
    def alloc(cls, m: int, n: int) -> "DeviceMatrix":
        h = C.c_void_p()
        _lib.check(_lib.load().mor_b200_mat_alloc(m, n, C.byref(h)))
        return cls(h)

    @classmethod
    def wrap(cls, device_ptr: int, m: int, n: int, ld: int) -> "DeviceMatrix":
        """Borrow caller-owned device memory, e.g. `t.data_ptr()` of a torch tensor holding
        the matrix column-major (a (n, ld) row-major tensor)."""
        h = C.c_void_p()
        _lib.check(_lib.load().mor_b200_mat_wrap(device_ptr, m, n, ld, C.byref(h)))
        return cls(h)

    @classmethod
    def from_host(cls, a: np.ndarray) -> "DeviceMatrix":
        a = np.asfortranarray(a, dtype=np.float64)
        d = cls.alloc(a.shape[0], a.shape[1])
        _lib.check(_lib.load().mor_b200_mat_upload(d.handle, a.ctypes.data, max(a.shape[0], 1)))
        return d

    @property
    def info(self):
        m, n, ld, p = C.c_int64(), C.c_int64(), C.c_int64(), C.c_void_p()
        _lib.check(_lib.load().mor_b200_mat_info(self.handle, C.byref(m), C.byref(n), C.byref(ld), C.byref(p)))
        return m.value, n.value, ld.value, p.value

    @property
    def shape(self):
        m, n, _, _ = self.info
        return m, n

    def to_host(self) -> np.ndarray:
        m, n = self.shape
        out = np.empty((m, n), dtype=np.float64, order="F")
        _lib.check(_lib.load().mor_b200_mat_download(self.handle, out.ctypes.data, max(m, 1)))
        return out

    def synth(self, kind: int, seed: int, global_row0: int = 0, param: float = 3.0, iparam: int = 0) -> "DeviceMatrix":
        _lib.check(_lib.load().mor_b200_synth(self.handle, kind, seed, global_row0, param, iparam))
        return self

    def kat_check(self, k: int, n: int, seed: int, global_row0: int = 0, param: float = 3.0) -> float:
        """For left vectors computed from a SYNTH_KAT matrix: max |U[i][j] -+ (H_u e_j)[i]| (collective)."""
        err = C.c_double(0.0)
        _lib.check(_lib.load().mor_b200_kat_check(self.handle, k, n, seed, global_row0, param, C.byref(err)))
        return err.value

    def free(self):
        if self.handle:
            _lib.load().mor_b200_mat_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceFactor:
    """mor_pod_t obtained from a device-resident snapshot matrix."""

    def __init__(self, handle, s: np.ndarray, keep=None):
        self.handle, self.s = handle, s
        self._keep = keep      # the handle may borrow X's device memory (include/mor_b200.h): keep it alive

    def basis(self, k: int, out: DeviceMatrix) -> DeviceMatrix:
        _lib.check(_lib.load().mor_b200_pod_basis_dev(self.handle, k, out.handle))
        return out

    def right(self, k: int, out: DeviceMatrix) -> DeviceMatrix:
        _lib.check(_lib.load().mor_b200_pod_right_dev(self.handle, k, out.handle))
        return out

    def info(self):
        passes, sweeps = C.c_int32(0), C.c_int32(0)
        _lib.check(_lib.load().mor_b200_pod_info(self.handle, C.byref(passes), C.byref(sweeps)))
        return passes.value, sweeps.value

    def free(self):
        if self.handle:
            _lib.load().mor_b200_pod_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def pod_factor_dev(X: DeviceMatrix, method: int, k: int, p: int = 0,
                   omega: Optional[DeviceMatrix] = None, seed: int = 0, flags: int = 0,
                   m_global: Optional[int] = None) -> DeviceFactor:
    m, n = X.shape
    cap = min(m_global if m_global is not None else m, n)
    s = np.empty(max(cap, 1), dtype=np.float64)
    nS = C.c_int64(0)
    h = C.c_void_p()
    _lib.check(_lib.load().mor_b200_pod_factor_dev(
        X.handle, method, k, p, omega.handle if omega is not None else None, seed, flags,
        C.byref(h), s.ctypes.data, C.byref(nS)))
    return DeviceFactor(h, s[:nS.value].copy(), keep=X)


def orthonormalize_dev(A: DeviceMatrix) -> int:
    passes = C.c_int32(0)
    _lib.check(_lib.load().mor_b200_orthonormalize_dev(A.handle, C.byref(passes)))
    return passes.value


def deim_indices_dev(B: DeviceMatrix) -> np.ndarray:
    _, k = B.shape
    idx = np.empty(k, dtype=np.int64)
    _lib.check(_lib.load().mor_b200_deim_indices_dev(B.handle, idx.ctypes.data))
    return idx
