"""ctypes binding of libmor_b200.so -- the same C ABI the Julia shim `ccall`s
(include/mor_b200.h).  There is no fallback: if the shared library is missing or a call
fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG_ROOT = Path(__file__).resolve().parents[2]          # modelorderreduction.jl_b200/
DEFAULT_LIB = _PKG_ROOT / "lib" / "libmor_b200.so"

# status codes (include/mor_b200.h)
OK, ERR_ARG, ERR_NONFINITE, ERR_CUDA, ERR_NOMEM, ERR_NOCONV, ERR_SINGULAR = range(7)
MOR_SVD, MOR_TSVD, MOR_RSVD = 0, 1, 2
SYNTH_GRADED, SYNTH_GAUSS, SYNTH_LOWRANK, SYNTH_KAT, SYNTH_BURGERS = 0, 1, 2, 3, 4
POD_MAY_OVERWRITE, POD_FORCE_TWO_PASS = 1, 2
T_SLOTS = ("total", "h2d", "gram", "apply", "core", "basis", "comm", "deim_step", "deim_tail",
           "d2h", "sketch", "deim_panel", "deim_blocked", "deim_update", "deim_fallback", "deim_exchange", "spmm", "proj")
T_NSLOTS = 24


class MorError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libmor_b200 status {status}: {message}")
        self.status = status


class SingularError(MorError, ArithmeticError):
    """The reference raises LinearAlgebra.SingularException here (deim.jl:21)."""


_i32, _i64, _u64, _dbl = C.c_int32, C.c_int64, C.c_uint64, C.c_double
_p = C.c_void_p
_pd = C.POINTER(C.c_double)
_pi64 = C.POINTER(C.c_int64)
_pi32 = C.POINTER(C.c_int32)

# name -> (restype, argtypes); this table is also what tests/test_abi_symbols.py checks
# against include/mor_b200.h
SIGNATURES = {
    "mor_b200_version": (_i32, []),
    "mor_b200_init": (_i32, [_i32]),
    "mor_b200_shutdown": (_i32, []),
    "mor_b200_last_error": (C.c_char_p, []),
    "mor_b200_set_stream": (_i32, [_p]),
    "mor_b200_sync": (_i32, []),
    "mor_b200_trim": (_i32, []),
    "mor_b200_host_alloc": (_i32, [_u64, C.POINTER(_p)]),
    "mor_b200_host_free": (_i32, [_p]),
    "mor_b200_init_multi": (_i32, [_i32]),
    "mor_b200_multi_info": (_i32, [_pi32]),
    "mor_b200_comm_unique_id": (_i32, [_p]),
    "mor_b200_comm_init": (_i32, [_i32, _i32, _p]),
    "mor_b200_comm_info": (_i32, [_pi32, _pi32]),
    "mor_b200_comm_destroy": (_i32, []),
    "mor_b200_pod_factor": (_i32, [_p, _i64, _i64, _i64, _i32, _i64, _i64, _p, _u64,
                                   C.POINTER(_p), _p, _pi64]),
    "mor_b200_pod_factor_cols": (_i32, [_p, _i64, _i64, _i32, _i64, _i64, _p, _u64,
                                        C.POINTER(_p), _p, _pi64]),
    "mor_b200_pod_basis": (_i32, [_p, _i64, _p, _i64, _p, _i64]),
    "mor_b200_pod_free": (_i32, [_p]),
    "mor_b200_pod_info": (_i32, [_p, _pi32, _pi32]),
    "mor_b200_deim_indices": (_i32, [_p, _i64, _i64, _i64, _p]),
    "mor_b200_deim_last_margins": (_i32, [_pd, _pd, _i64, _pi64, _pd, _pi64]),
    "mor_b200_deim_projector": (_i32, [_p, _i64, _i64, _i64, _p, _i64, _i64, _p, _p, _i64]),
    "mor_b200_deim_projector_dev": (_i32, [_p, _p, _p, _p]),
    "mor_b200_galerkin_csc": (_i32, [_i64, _p, _p, _p, _p, _i64, _i64, _p, _i64]),
    "mor_b200_project": (_i32, [_p, _i64, _i64, _i64, _p, _i64, _i64, _p, _i64]),
    "mor_b200_mat_alloc": (_i32, [_i64, _i64, C.POINTER(_p)]),
    "mor_b200_mat_wrap": (_i32, [_p, _i64, _i64, _i64, C.POINTER(_p)]),
    "mor_b200_mat_free": (_i32, [_p]),
    "mor_b200_mat_info": (_i32, [_p, _pi64, _pi64, _pi64, C.POINTER(_p)]),
    "mor_b200_mat_upload": (_i32, [_p, _p, _i64]),
    "mor_b200_mat_download": (_i32, [_p, _p, _i64]),
    "mor_b200_synth": (_i32, [_p, _i32, _u64, _i64, _dbl, _i64]),
    "mor_b200_kat_check": (_i32, [_p, _i64, _i64, _u64, _i64, _dbl, _pd]),
    "mor_b200_pod_factor_dev": (_i32, [_p, _i32, _i64, _i64, _p, _u64, _i32, C.POINTER(_p), _p,
                                       _pi64]),
    "mor_b200_pod_basis_dev": (_i32, [_p, _i64, _p]),
    "mor_b200_pod_right_dev": (_i32, [_p, _i64, _p]),
    "mor_b200_orthonormalize_dev": (_i32, [_p, _pi32]),
    "mor_b200_deim_indices_dev": (_i32, [_p, _p]),
    "mor_b200_last_timings": (_i32, [_pd, _i32]),
    "mor_b200_launch_count": (_i64, [_i32]),
}

_lib = None
_test_lib = None
TEST_LIB = _PKG_ROOT / "lib" / "libmor_b200_test.so"


def lib_path() -> Path:
    return Path(os.environ.get("MOR_B200_LIB", str(DEFAULT_LIB)))


def load() -> C.CDLL:
    """Load libmor_b200.so (once) and attach the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise FileNotFoundError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C modelorderreduction.jl_b200/csrc`).  There is no CPU fallback.")
    lib = C.CDLL(str(path), mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def load_test() -> C.CDLL:
    """libmor_b200_test.so: kernel-level test hooks and peak probes (mor_dbg_*).  Not part of the product
    ABI; used by tests/ and bench.py only."""
    global _test_lib
    if _test_lib is None:
        load()
        if not TEST_LIB.exists():
            raise FileNotFoundError(f"{TEST_LIB} not found: build it with make -C modelorderreduction.jl_b200/csrc")
        _test_lib = C.CDLL(str(TEST_LIB), mode=C.RTLD_GLOBAL)
    return _test_lib


def check(status: int) -> None:
    if status == OK:
        return
    msg = load().mor_b200_last_error().decode("utf-8", "replace")
    if status in (ERR_ARG, ERR_NONFINITE):
        raise ValueError(f"libmor_b200 status {status}: {msg}")     # shim: ArgumentError
    if status == ERR_SINGULAR:
        raise SingularError(status, msg)
    if status == ERR_NOMEM:
        raise MemoryError(f"libmor_b200 status {status}: {msg}")
    raise MorError(status, msg)


def last_timings() -> dict:
    buf = (C.c_double * T_NSLOTS)()
    check(load().mor_b200_last_timings(buf, T_NSLOTS))
    return {name: buf[i] for i, name in enumerate(T_SLOTS)}
