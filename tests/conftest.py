import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "modelorderreduction.jl_b200", "python"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


def _unhex(rows):
    import struct
    return np.asfortranarray([[struct.unpack(">d", bytes.fromhex(v))[0] for v in row] for row in rows])


@pytest.fixture(scope="session")
def mp_golden():
    """50-digit mpmath answers (oracle/make_golden_mpmath.py): the pin that is independent of LAPACK."""
    with open(os.path.join(ROOT, "tests", "golden", "mpmath_golden.json")) as f:
        g = json.load(f)
    g["F2"]["U"] = _unhex(g["F2"]["U_hex"])
    g["F2"]["Q5"] = _unhex(g["F2"]["Q5_hex"])
    for c in g["deim_random"]:
        c["basis"] = _unhex(c["basis_hex"])
    return g


@pytest.fixture(scope="session")
def gpu():
    """Initialises libmor_b200 on cuda:0.  A missing library is an error (never a skip);
    only the absence of any CUDA device skips, so `pytest tests/` stays usable on a CPU box."""
    from mor_b200 import _lib
    lib = _lib.load()
    st = lib.mor_b200_init(0)
    if st == _lib.ERR_CUDA and b"no CUDA device" in lib.mor_b200_last_error():
        pytest.skip("no CUDA device in this container")
    _lib.check(st)
    return lib


def sin_fixture():
    n, m = 20, 10
    return np.array([[np.sin(i * j / n) for j in range(1, m + 1)] for i in range(1, n + 1)], order="F")
