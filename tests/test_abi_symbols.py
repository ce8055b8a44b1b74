"""The C-ABI library loads and exports exactly what include/mor_b200.h declares, and the
ctypes/Julia bindings name the same symbols (CPU only: no compute calls)."""
import os
import re

import pytest

from conftest import ROOT
from mor_b200 import _lib


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mor_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mor_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in mor_b200.h but not exported"


def test_product_library_exports_only_the_header_abi():
    """nm -D: the extern "C" surface of libmor_b200.so is exactly include/mor_b200.h; the kernel test hooks
    (mor_dbg_*) live in libmor_b200_test.so."""
    import subprocess
    libdir = os.path.join(ROOT, "modelorderreduction.jl_b200", "lib")
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(libdir, "libmor_b200.so")], capture_output=True, text=True, check=True).stdout
    exported = sorted({line.split()[-1] for line in out.splitlines() if line.split()[-1].startswith(("mor_b200_", "mor_dbg_"))})
    assert exported == header_symbols()
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(libdir, "libmor_b200_test.so")], capture_output=True, text=True, check=True).stdout
    assert any(line.endswith("mor_dbg_peaks") for line in out.splitlines())


def test_bindings_cover_header():
    syms = header_symbols()
    assert sorted(_lib.SIGNATURES) == syms
    shim = open(os.path.join(ROOT, "modelorderreduction.jl_b200", "julia", "MORB200.jl")).read()
    for s in ("mor_b200_init", "mor_b200_pod_factor", "mor_b200_pod_factor_cols", "mor_b200_pod_basis",
              "mor_b200_pod_free", "mor_b200_deim_indices", "mor_b200_last_error",
              "mor_b200_comm_unique_id", "mor_b200_comm_init"):
        assert s in shim, f"{s} not bound by the Julia shim"


def test_version_and_error_channel():
    lib = _lib.load()
    assert lib.mor_b200_version() == 100
    # argument errors are reported without touching a device
    assert lib.mor_b200_mat_info(None, None, None, None, None) == _lib.ERR_ARG
    assert b"null" in lib.mor_b200_last_error()


def test_no_cpu_fallback_without_device():
    """On a box without a GPU every compute entry point must fail loudly (status 3)."""
    import numpy as np
    lib = _lib.load()
    st = lib.mor_b200_init(0)
    if st == _lib.OK:
        pytest.skip("a CUDA device is present")
    assert st == _lib.ERR_CUDA
    B = np.eye(4, 2, order="F")
    idx = np.zeros(2, dtype=np.int64)
    assert lib.mor_b200_deim_indices(B.ctypes.data, 4, 2, 4, idx.ctypes.data) == _lib.ERR_CUDA
    import mor_b200
    with pytest.raises(_lib.MorError):
        mor_b200.deim_interpolation_indices(B)


def _build_c_consumer(tmp_path):
    import subprocess
    libdir = os.path.join(ROOT, "modelorderreduction.jl_b200", "lib")
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe, "-L" + libdir, "-lmor_b200",
                    "-Wl,-rpath," + libdir, "-lm"], check=True)
    return exe


def test_header_is_plain_c_and_links(tmp_path):
    """include/mor_b200.h compiles as strict C99 and a plain-C program links against the library;
    without a device it gets MOR_ERR_CUDA from the compute entry points (exit code 3)."""
    import subprocess
    exe = _build_c_consumer(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode in (0, 3), out.stdout + out.stderr
    if out.returncode == 3:
        assert "no CUDA device" in out.stdout


@pytest.mark.gpu
def test_plain_c_consumer_on_gpu(tmp_path):
    import subprocess
    out = subprocess.run([_build_c_consumer(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c_abi_smoke ok" in out.stdout
