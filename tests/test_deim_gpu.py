"""DEIM parity on the GPU through the C ABI: indices must equal the oracle's exactly
(Int64 equality), including first-index tie-break and NaN-is-maximal."""
import numpy as np
import pytest
import scipy.linalg as sla

import mor_b200
import mor_oracle as O
from conftest import sin_fixture
from mor_b200 import _lib
from mor_b200.device import DeviceMatrix, deim_indices_dev

pytestmark = pytest.mark.gpu


def test_sin_fixture_goldens(gpu, golden):
    S = sin_fixture()
    Q, _ = sla.qr(S, mode="economic")
    assert mor_b200.deim_interpolation_indices(np.asfortranarray(Q[:, :5])).tolist() == golden["F2"]["deim_Q5"]
    U, _, _ = O._svd(S)
    assert mor_b200.deim_interpolation_indices(U).tolist() == golden["F2"]["deim_U10"]


def test_golden_random(gpu, golden):
    for case in golden["deim_random"]:
        rng = np.random.Generator(np.random.PCG64(case["seed"]))
        Q, _ = np.linalg.qr(rng.standard_normal((case["m"], case["k"])))
        got = mor_b200.deim_interpolation_indices(Q)
        assert got.dtype == np.int64
        assert got.tolist() == case["indices"]


@pytest.mark.parametrize("m,k,seed", [(2, 2, 0), (3, 1, 1), (513, 7, 2), (2048, 16, 3), (2049, 16, 4),
                                      (10_000, 40, 5), (65_537, 32, 6), (300_001, 64, 7)])
def test_vs_oracle_random(gpu, m, k, seed):
    rng = np.random.Generator(np.random.PCG64(100 + seed))
    Q, _ = np.linalg.qr(rng.standard_normal((m, k)))
    assert mor_b200.deim_interpolation_indices(np.asfortranarray(Q)).tolist() == \
        O.deim_interpolation_indices(Q).tolist()


def test_layouts(gpu):
    rng = np.random.Generator(np.random.PCG64(11))
    Q, _ = np.linalg.qr(rng.standard_normal((4100, 12)))
    want = O.deim_interpolation_indices(Q).tolist()
    assert mor_b200.deim_interpolation_indices(np.ascontiguousarray(Q)).tolist() == want   # row-major -> copied
    big = np.zeros((4133, 12), order="F"); big[:4100] = Q
    assert mor_b200.deim_interpolation_indices(big[:4100]).tolist() == want                # ld > m (odd ld)
    off = np.zeros((4101, 12), order="F"); off[1:] = Q
    assert mor_b200.deim_interpolation_indices(off[1:]).tolist() == want                   # 8-byte aligned only


def test_ties_and_nan(gpu):
    B = np.zeros((8, 2), order="F")
    B[:, 0] = [1, -2, 2, 0.5, 2, -2, 1, 0]
    B[:, 1] = [0, 1, 1, 4, -3, 4, 4, 4]
    assert mor_b200.deim_interpolation_indices(B).tolist() == O.deim_interpolation_indices(B).tolist() == [2, 7]
    Bn = B.copy(order="F"); Bn[5, 0] = np.nan; Bn[6, 0] = np.nan
    assert mor_b200.deim_interpolation_indices(Bn)[0] == 6
    # exact ties far apart (different CTAs / chunks): integer-valued data, all arithmetic exact
    m = 50_000
    T = np.zeros((m, 3), order="F")
    T[:, 0] = 1.0; T[40_000, 0] = -4.0; T[123, 0] = 4.0; T[45_001, 0] = 4.0
    T[:, 1] = 2.0; T[777, 1] = 8.0; T[30_000, 1] = -8.0
    T[:, 2] = (np.arange(m) % 7) - 3.0; T[49_999, 2] = 64.0; T[9, 2] = 64.0
    assert mor_b200.deim_interpolation_indices(T).tolist() == O.deim_interpolation_indices(T).tolist()


def test_singular_raises(gpu):
    Z = np.zeros((4, 3), order="F"); Z[:, 1] = 1; Z[:, 2] = [1, 2, 3, 4]
    with pytest.raises(np.linalg.LinAlgError):
        O.deim_interpolation_indices(Z)
    with pytest.raises(mor_b200.SingularError):
        mor_b200.deim_interpolation_indices(Z)
    with pytest.raises(ValueError):
        mor_b200.deim_interpolation_indices(np.zeros((4, 0), order="F"))


def test_device_resident_and_properties(gpu):
    """Device-generated basis (no host copy of the input): the selected rows must make P'U
    nonsingular, be distinct, and equal the oracle's on the downloaded matrix."""
    m, k = 200_000, 48
    B = DeviceMatrix.alloc(m, k).synth(_lib.SYNTH_GAUSS, seed=5, param=1.0 / np.sqrt(m))
    idx = deim_indices_dev(B)
    assert len(set(idx.tolist())) == k and idx.min() >= 1 and idx.max() <= m
    host = B.to_host()
    assert idx.tolist() == O.deim_interpolation_indices(host).tolist()
    t = _lib.last_timings()
    assert t["deim_step"] > 0 and t["total"] >= t["deim_step"]
