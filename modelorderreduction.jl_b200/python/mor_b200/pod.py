"""Host-side mirror of ModelOrderReduction.jl's data-reduction API, numeric calls rerouted
to libmor_b200.so.

Reference (paths relative to /root/reference/src):
  Types.jl:61-124            SVD / TSVD / RSVD backend tags
  ErrorHandle.jl:1-18        errorhandle (AssertionError, raised BEFORE any library call)
  DataReduction/POD.jl:79-119   POD struct + 4 constructors
  DataReduction/POD.jl:121-132  determine_truncation (host scalar loop, quirks preserved)
  DataReduction/POD.jl:156-238  reduce!(pod, SVD()/TSVD()/RSVD())

What is rerouted: the `_svd/_tsvd/_rsvd` seams (POD.jl:8-27) -> mor_b200_pod_factor[_cols]
and the `u[:, 1:nmodes]` copy (POD.jl:163,199,235) -> mor_b200_pod_basis.  Everything else
is the reference's own host logic restated in Python (Julia is not available in the build
image; julia/MORB200.jl is the same shim for a Julia host).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib

Snapshots = Union[np.ndarray, Sequence[np.ndarray]]

MAX_SMALL_DIM = 4096       # include/mor_b200.h: the small side of the snapshot matrix
_SVD_KWARGS_OK = {"alg"}                              # LAPACK driver choice: same result
_TSVD_KWARGS_OK = {"initvec", "tolconv", "maxiter"}   # TSVD.jl's Lanczos knobs: the exact top-k is returned


class DeferToReference(NotImplementedError):
    """The call is outside what libmor_b200 handles; the Julia shim `invoke`s the reference's own method here
    (julia/MORB200.jl: `defers`).  This Python mirror has no reference implementation to fall back to, so it
    says so instead of guessing."""


def defer_reason(snaps: Snapshots, alg) -> Optional[str]:
    """The shim's rule (MORB200.jl `defers`): why this call must run the reference's own method, or None.
    /root/reference/src/Types.jl:61-95 forwards `alg.kwargs` verbatim (POD.jl:157,196); ErrorHandle.jl:1-7 accepts
    any size."""
    m, n = (len(snaps[0]), len(snaps)) if _is_vov(snaps) else snaps.shape
    if min(m, n) > MAX_SMALL_DIM:
        return f"small dimension {min(m, n)} > {MAX_SMALL_DIM}"
    kw = getattr(alg, "kwargs", {}) or {}
    for key, val in kw.items():
        if key == "full" and val is False:
            continue
        if isinstance(alg, SVD) and key in _SVD_KWARGS_OK:
            continue
        if isinstance(alg, TSVD) and key in _TSVD_KWARGS_OK:
            continue
        return f"backend keyword argument {key}={val!r} is not understood by libmor_b200"
    if not _is_vov(snaps) and snaps.dtype != np.float64:
        return f"element type {snaps.dtype} (the library is Float64 only)"
    return None


# ---- backend tags (Types.jl:61-124) ---------------------------------------------------------
class SVD:
    """Dense SVD backend (Types.jl:61-70).  kwargs are accepted for source compatibility;
    `alg`/`full=false` need no action on this path."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs


class TSVD:
    """Truncated SVD backend (Types.jl:86-95).  `initvec/tolconv/maxiter` steer TSVD.jl's
    Lanczos iteration; the B200 path computes the exact top-k triplets directly, so they
    are accepted and ignored."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs


class RSVD:
    """Randomized SVD backend with oversampling p (Types.jl:119-124); `RSVD(2).p == 2`."""

    def __init__(self, p: int = 0):
        self.p = int(p)


def _is_vov(snaps) -> bool:
    return not isinstance(snaps, np.ndarray)


def errorhandle(snaps: Snapshots, nmodes: int, max_energy: float, min_nmodes: int,
                max_nmodes: int) -> None:
    """ErrorHandle.jl:1-18."""
    if _is_vov(snaps):
        m, n = len(snaps[0]), len(snaps)
    else:
        m, n = snaps.shape
    assert m > 1, "State vector is expected to be vector valued."
    s = min(m, n)
    assert 0 < nmodes <= s, f"Number of modes should be in {{1,2,...,{s}}}."
    assert min_nmodes <= max_nmodes, \
        "Minimum number of modes must lie below maximum number of modes"
    assert 0.0 <= max_energy <= 1.0, "Maximum relative energy must be in [0,1]"


class POD:
    """POD.jl:79-119.  Positional form `POD(snaps, k)`; keyword form
    `POD(snaps, min_renergy=1.0, min_nmodes=1, max_nmodes=length(snaps[1]))` -- for a
    Matrix `snaps[1]` is a scalar, so the default max_nmodes is 1 (reference behaviour)."""

    def __init__(self, snaps: Snapshots, nmodes: Optional[int] = None, *,
                 min_renergy: float = 1.0, min_nmodes: int = 1,
                 max_nmodes: Optional[int] = None):
        if nmodes is not None:
            errorhandle(snaps, nmodes, 0.0, nmodes, nmodes)
            self.min_renergy, self.min_nmodes, self.max_nmodes = 0.0, nmodes, nmodes
            self.nmodes = nmodes
        else:
            if max_nmodes is None:
                max_nmodes = len(snaps[0]) if _is_vov(snaps) else 1
            errorhandle(snaps, min_nmodes, min_renergy, min_nmodes, max_nmodes)
            self.min_renergy, self.min_nmodes, self.max_nmodes = min_renergy, min_nmodes, max_nmodes
            self.nmodes = min_nmodes
        self.snapshots = snaps          # kept by reference, never modified
        self.rbasis = None              # `missing`
        self.renergy = 1.0
        self.spectrum = None

    def __repr__(self):                 # POD.jl:240-249
        m = len(self.snapshots[0]) if _is_vov(self.snapshots) else self.snapshots.shape[0]
        return (f"POD \nReduction Order = {self.nmodes}\nSnapshot size = ({m},...)\n"
                f"Relative Energy = {self.renergy}\n")


def determine_truncation(s: np.ndarray, min_nmodes: int, max_nmodes: int,
                         min_renergy: float) -> Tuple[int, float]:
    """POD.jl:121-132 verbatim, including the plain-sigma energy, the off-by-one
    `s[nmodes + 1]` and the out-of-bounds error when the loop reaches the end."""
    nmodes = min_nmodes
    overall_energy = float(np.sum(s))
    energy = float(np.sum(s[:nmodes])) / overall_energy
    while energy < min_renergy and nmodes < max_nmodes:
        nmodes += 1
        if nmodes + 1 > len(s):
            raise IndexError(f"BoundsError: attempt to access {len(s)}-element Vector at index [{nmodes + 1}]")
        energy += float(s[nmodes]) / overall_energy
    return nmodes, energy


# ---- the rerouted seams -----------------------------------------------------------------------
class _Factor:
    """Owns a mor_pod_t handle."""

    def __init__(self, handle, m: int, n: int, s: np.ndarray):
        self.handle, self.m, self.n, self.s = handle, m, n, s

    def basis(self, k: int, want_v: bool = False):
        lib = _lib.load()
        U = np.empty((self.m, k), dtype=np.float64, order="F")
        V = np.empty((self.n, k), dtype=np.float64, order="F") if want_v else None
        _lib.check(lib.mor_b200_pod_basis(self.handle, k, U.ctypes.data, max(self.m, 1),
                                          V.ctypes.data if want_v else None, max(self.n, 1)))
        return (U, V) if want_v else U

    def info(self):
        passes, sweeps = C.c_int32(0), C.c_int32(0)
        _lib.check(_lib.load().mor_b200_pod_info(self.handle, C.byref(passes), C.byref(sweeps)))
        return passes.value, sweeps.value

    def close(self):
        if self.handle:
            _lib.load().mor_b200_pod_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _factor(snaps: Snapshots, method: int, k: int, p: int = 0,
            omega: Optional[np.ndarray] = None, seed: int = 0, m_global: Optional[int] = None) -> _Factor:
    """mor_b200_pod_factor / _cols: Matrix input goes by pointer + ld, a vector of vectors
    goes as an array of column pointers (no hcat copy, POD.jl:4-6)."""
    lib = _lib.load()
    handle = C.c_void_p()
    nS = C.c_int64(0)
    om_ptr = None
    if omega is not None:
        omega = np.asfortranarray(omega, dtype=np.float64)
        om_ptr = omega.ctypes.data
    if _is_vov(snaps):
        cols = [np.ascontiguousarray(c, dtype=np.float64) for c in snaps]
        m, n = len(cols[0]), len(cols)
        for c in cols:
            if c.shape != (m,):
                raise ValueError("all snapshots must have the same length")
        cap = n                                     # min(M_global, n) <= n, also for a row shard
        s = np.empty(cap, dtype=np.float64)
        ptrs = (C.c_void_p * n)(*[c.ctypes.data for c in cols])
        st = lib.mor_b200_pod_factor_cols(ptrs, m, n, method, k, p, om_ptr, seed,
                                          C.byref(handle), s.ctypes.data, C.byref(nS))
    else:
        X = snaps
        if X.dtype != np.float64 or X.ndim != 2:
            raise TypeError("the B200 path handles Float64 matrices (the shim defers other types to the reference)")
        m, n = X.shape
        # StridedMatrix with unit row stride goes zero-copy; anything else is made column-major
        strided_ok = (m <= 1 or X.strides[0] == 8) and (n <= 1 or (X.strides[1] % 8 == 0 and X.strides[1] // 8 >= m))
        if not strided_ok:
            X = np.asfortranarray(X)
        ldx = X.strides[1] // 8 if n > 1 else max(m, 1)
        cap = n
        s = np.empty(cap, dtype=np.float64)
        st = lib.mor_b200_pod_factor(X.ctypes.data, m, n, max(ldx, 1), method, k, p, om_ptr, seed,
                                     C.byref(handle), s.ctypes.data, C.byref(nS))
    _lib.check(st)
    return _Factor(handle, m, n, s[:nS.value].copy())


def reduce(pod: POD, alg, *, omega: Optional[np.ndarray] = None, seed: int = 0) -> None:
    """`reduce!(pod, alg)` -- mutates `pod`, returns None (POD.jl:156-166,195-202,231-238).
    `omega`/`seed` exist only for RSVD: the reference draws Omega from Julia's global RNG;
    here it is either supplied (shared with the oracle) or drawn by the library from `seed`."""
    if isinstance(alg, (SVD, TSVD, RSVD)):
        why = defer_reason(pod.snapshots, alg)
        if why is not None:
            raise DeferToReference(why)
    if isinstance(alg, SVD):
        f = _factor(pod.snapshots, _lib.MOR_SVD, pod.max_nmodes)
        try:
            s = f.s                                                       # full spectrum
            pod.nmodes, pod.renergy = determine_truncation(
                s, pod.min_nmodes, pod.max_nmodes, pod.min_renergy)      # POD.jl:158-162
            pod.rbasis = f.basis(pod.nmodes)                              # POD.jl:163
            pod.spectrum = s                                              # POD.jl:164
        finally:
            f.close()
    elif isinstance(alg, (TSVD, RSVD)):
        is_r = isinstance(alg, RSVD)
        f = _factor(pod.snapshots, _lib.MOR_RSVD if is_r else _lib.MOR_TSVD, pod.nmodes,
                    alg.p if is_r else 0, omega if is_r else None, seed)
        try:
            s = f.s
            n_max = min(f.m, f.n)                                         # POD.jl:197,233
            pod.renergy = float(np.sum(s) / (np.sum(s) + (n_max - pod.nmodes) * s[-1]))
            pod.rbasis = f.basis(pod.nmodes)                              # POD.jl:199,235
            pod.spectrum = s                                              # k values
        finally:
            f.close()
    else:
        raise TypeError(f"no method matching reduce!(::POD, ::{type(alg).__name__})")
    return None


reduce_bang = reduce
