"""Host-side mirror (mor_b200) logic that runs without a GPU: constructors, assertions,
truncation quirks, shard partitioning -- checked against the oracle / reference behaviour."""
import os

import numpy as np
import pytest

import mor_b200
import mor_oracle as O
from mor_b200.dist import shard_rows


def test_tags_and_ctor_match_reference_semantics():
    assert mor_b200.RSVD(2).p == 2 and mor_b200.RSVD().p == 0          # Types.jl:113-117
    X = np.array([[3.0, 0.0], [0.0, 1.0]])
    pod = mor_b200.POD(X, 1)
    assert pod.nmodes == 1 and pod.min_renergy == 0.0 and pod.max_nmodes == 1   # POD.jl:70-77,111-114
    assert pod.rbasis is None and pod.spectrum is None and pod.renergy == 1.0
    assert mor_b200.POD(np.ones((5, 3))).max_nmodes == 1
    assert mor_b200.POD([np.ones(5)] * 3).max_nmodes == 5
    for bad in (dict(snaps=np.ones((1, 3)), nmodes=1), dict(snaps=np.ones((5, 3)), nmodes=4),
                dict(snaps=np.ones((5, 3)), nmodes=0)):
        with pytest.raises(AssertionError):
            mor_b200.POD(bad["snaps"], bad["nmodes"])
        with pytest.raises(AssertionError):
            O.POD(bad["snaps"], bad["nmodes"])
    with pytest.raises(AssertionError):
        mor_b200.POD(np.ones((5, 3)), min_renergy=1.5)
    with pytest.raises(AssertionError):
        mor_b200.POD(np.ones((5, 3)), min_nmodes=2, max_nmodes=1)


def test_determine_truncation_matches_oracle(golden):
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(2, 12))
        s = np.sort(rng.random(n))[::-1] + 1e-3
        lo = int(rng.integers(1, n + 1)); hi = int(rng.integers(lo, n + 1))
        mre = float(rng.random())
        try:
            want = O.determine_truncation(s, lo, hi, mre)
        except IndexError:
            with pytest.raises(IndexError):
                mor_b200.determine_truncation(s, lo, hi, mre)
            continue
        assert mor_b200.determine_truncation(s, lo, hi, mre) == want
    s = np.array([4.0, 3.0, 2.0, 1.0])
    for case in golden["F5"]:
        assert list(mor_b200.determine_truncation(s, *case["args"])) == case["out"]


def test_shard_rows_partition():
    for m in (0, 1, 7, 1000, 16_777_216, 12345677):
        for g in (1, 2, 3, 4, 8):
            pos = 0
            for r in range(g):
                r0, ml = shard_rows(m, g, r)
                assert r0 == pos and ml >= 0
                assert r0 % 4 == 0 or ml == 0
                pos += ml
            assert pos == m
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def test_reduce_rejects_unknown_backend():
    pod = mor_b200.POD(np.ones((4, 2)), 1)
    with pytest.raises(TypeError):
        mor_b200.reduce(pod, object())


def test_shim_defers_outside_the_library_limits():
    """MORB200.jl `defers` / pod.py `defer_reason`: the shim never errors where the reference would succeed
    (Types.jl:61-95 kwargs forwarded at POD.jl:157,196; ErrorHandle.jl accepts any size) -- it runs the reference's own
    method instead.  The Python mirror has no reference to fall back to and raises DeferToReference BEFORE any
    library call (so this runs without a GPU)."""
    import mor_b200
    X = np.zeros((5000, 4200))
    assert "small dimension" in mor_b200.defer_reason(X, mor_b200.SVD())
    with pytest.raises(mor_b200.DeferToReference):
        mor_b200.reduce(mor_b200.POD(X, 3), mor_b200.SVD())
    Y = np.ones((10, 4))
    assert mor_b200.defer_reason(Y, mor_b200.SVD()) is None
    assert mor_b200.defer_reason(Y, mor_b200.SVD(full=False)) is None
    assert mor_b200.defer_reason(Y, mor_b200.SVD(alg="QRIteration")) is None
    assert "full" in mor_b200.defer_reason(Y, mor_b200.SVD(full=True))
    assert mor_b200.defer_reason(Y, mor_b200.TSVD(tolconv=1e-12, maxiter=50)) is None
    assert "stepsize" in mor_b200.defer_reason(Y, mor_b200.TSVD(stepsize=3))
    assert "float32" in mor_b200.defer_reason(Y.astype(np.float32), mor_b200.SVD())
    with pytest.raises(mor_b200.DeferToReference):
        mor_b200.deim_interpolation_indices(np.zeros((2000, 1025), order="F"))
    shim = open(os.path.join(os.path.dirname(__file__), "..", "modelorderreduction.jl_b200", "julia", "MORB200.jl")).read()
    for needle in ("invoke(reduce!, Tuple{POD, SVD}", "invoke(reduce!, Tuple{POD, TSVD}", "invoke(reduce!, Tuple{POD, RSVD}",
                   "invoke(deim_interpolation_indices", "mor_b200_init_multi", "Vector{Float64}(undef, n)"):
        assert needle in shim
