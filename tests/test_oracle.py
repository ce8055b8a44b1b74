"""Pins the oracle: analytic answers of the reference's fixtures, the reference's own
range assertions, and the committed golden vectors (CPU only)."""
import numpy as np
import pytest
import scipy.linalg as sla

import mor_oracle as O
from conftest import sin_fixture


def test_f1_interface_fixture_analytic():
    # test/interface.jl:3-17 -- analytic: s=[3,1], rbasis=+-[1,0], renergy .75 (SVD) / .5 (TSVD,RSVD)
    X = np.array([[3.0, 0.0], [0.0, 1.0]], order="F")
    vov = [np.array([3.0, 0.0]), np.array([0.0, 1.0])]
    for snaps in (X, vov):
        for alg, re in ((O.SVD(), 0.75), (O.TSVD(), 0.5), (O.RSVD(), 0.5)):
            pod = O.POD(snaps, 1)
            assert O.reduce(pod, alg, omega=np.array([[1.0], [0.25]])) is None
            assert pod.rbasis.shape == (2, 1)
            assert len(pod.spectrum) >= pod.nmodes
            assert pod.nmodes == 1
            assert 0.0 < pod.renergy <= 1.0
            assert pod.renergy == pytest.approx(re, abs=1e-14)
            if isinstance(alg, O.RSVD):     # rank-1 sketch of a rank-2 matrix: approximate
                assert 1.0 <= pod.spectrum[0] <= 3.0 + 1e-14
            else:
                assert pod.spectrum[0] == pytest.approx(3.0, abs=1e-14)
                np.testing.assert_allclose(np.abs(pod.rbasis[:, 0]), [1.0, 0.0], atol=1e-14)


def test_f2_sin_fixture_golden(golden):
    S = sin_fixture()
    pod = O.POD(S, 3)
    O.reduce(pod, O.SVD())
    np.testing.assert_allclose(pod.spectrum, golden["F2"]["spectrum"], rtol=1e-12, atol=1e-15)
    assert pod.renergy == pytest.approx(golden["F2"]["renergy_svd_k3"], rel=1e-13)
    pod = O.POD(S, 3)
    O.reduce(pod, O.TSVD())
    assert pod.renergy == pytest.approx(golden["F2"]["renergy_tsvd_k3"], rel=1e-13)
    assert pod.spectrum.shape == (3,)
    Q, _ = sla.qr(S, mode="economic")
    assert O.deim_interpolation_indices(Q[:, :5]).tolist() == golden["F2"]["deim_Q5"] == [20, 11, 16, 5, 19]
    U, _, _ = O._svd(S)
    assert O.deim_interpolation_indices(U).tolist() == golden["F2"]["deim_U10"]


def test_oracle_pinned_by_mpmath_goldens(mp_golden, golden):
    """The independent pin (missing #1): the oracle's LAPACK answers on the reference's own fixture
    (src/precompile.jl:5-12) against 50-digit mpmath values computed without LAPACK/BLAS, and the DEIM chain of
    deim.jl:13-24 against its 50-digit evaluation -- with the margin by which each argmax is decided."""
    f2 = mp_golden["F2"]
    S = sin_fixture()
    _, s, _ = O._svd(S)
    s_mp = np.array(f2["sigma"])
    # kappa = 2e10: dgesdd is accurate to ~eps * sigma_1 in absolute terms, so the gate is SURVEY 8d's
    # (relative above 1e-5 sigma_1, absolute / sigma_1 below)
    assert O.sigma_error(s_mp, s) <= 1e-10
    assert np.max(np.abs(s - s_mp)) <= 8 * np.finfo(float).eps * s_mp[0]
    assert O.sigma_error(s_mp, np.array(golden["F2"]["spectrum"])) <= 1e-10      # the committed LAPACK golden too
    pod = O.POD(S, 3); O.reduce(pod, O.SVD())
    assert pod.renergy == pytest.approx(f2["renergy_svd_k3"], rel=1e-13)
    pod = O.POD(S, 3); O.reduce(pod, O.TSVD())
    assert pod.renergy == pytest.approx(f2["renergy_tsvd_k3"], rel=1e-13)
    u, _, _ = O._svd(S)
    # left vectors: sigma_1 - sigma_2 = 1.1e-3, so columns 1, 2 are only determined to ~eps / gap
    assert O.subspace_sine(f2["U"][:, :4], u[:, :4]) <= 1e-12
    for basis, chain in ((f2["U"], f2["deim_U10"]), (f2["Q5"], f2["deim_Q5"])) + \
            tuple((c["basis"], c["chain"]) for c in mp_golden["deim_random"]):
        assert chain["min_relative_margin"] > 1e-3           # decided far above any FP64 rounding error
        assert O.deim_interpolation_indices(basis).tolist() == chain["indices"]
        assert O.deim_interpolation_indices_blocked(basis, block=4, sub=2).tolist() == chain["indices"]
    # the LAPACK-made bases of golden.json select the same points as the mpmath-made ones
    assert golden["F2"]["deim_U10"] == f2["deim_U10"]["indices"]
    assert golden["F2"]["deim_Q5"] == f2["deim_Q5"]["indices"]
    Q, _ = sla.qr(S, mode="economic")
    assert np.max(np.abs(np.abs(Q[:, :5]) - np.abs(f2["Q5"]))) <= 1e-14


def test_f5_truncation_quirks(golden):
    s = np.array([4.0, 3.0, 2.0, 1.0])
    # probe from SURVEY 8a: mre=0.5 -> (2, 0.6) although the true 2-mode energy is 0.7
    assert O.determine_truncation(s, 1, 4, 0.5) == (2, pytest.approx(0.6))
    for case in golden["F5"]:
        n, e = O.determine_truncation(s, *case["args"])
        assert n == case["out"][0] and e == pytest.approx(case["out"][1], rel=1e-15)
    with pytest.raises(IndexError):      # Julia: BoundsError
        O.determine_truncation(s, 1, 4, 1.0)


def test_pod_ctor_defaults_and_asserts():
    X = np.ones((5, 3))
    assert O.POD(X).max_nmodes == 1                       # length(snaps[1]) of a Matrix
    assert O.POD([np.ones(5)] * 3).max_nmodes == 5        # ... of a Vector{Vector}
    with pytest.raises(AssertionError):
        O.POD(np.ones((1, 3)), 1)
    with pytest.raises(AssertionError):
        O.POD(X, 4)
    with pytest.raises(AssertionError):
        O.POD(X, 0)
    with pytest.raises(AssertionError):
        O.POD(X, min_renergy=1.5)
    with pytest.raises(AssertionError):
        O.POD(X, min_nmodes=2, max_nmodes=1)


def test_lorenz_ranges():
    # test/DataReduction.jl:27-67 (range assertions only; trajectory via scipy instead of Tsit5)
    from scipy.integrate import solve_ivp

    def lorenz(t, u):
        return [10.0 * (u[1] - u[0]), u[0] * (28.0 - u[2]) - u[1], u[0] * u[1] - (8 / 3) * u[2]]
    sol = solve_ivp(lorenz, (0, 100), [1.0, 0.0, 0.0], rtol=1e-8, atol=1e-8, max_step=0.05)
    X = np.asfortranarray(sol.y)
    vov = [X[:, j].copy() for j in range(X.shape[1])]
    assert O.matricize(vov).shape == X.shape
    p1, p2 = O.POD(X, 2), O.POD(vov, 2)
    O.reduce(p1, O.SVD()); O.reduce(p2, O.SVD())
    np.testing.assert_allclose(p1.rbasis, p2.rbasis)
    assert p1.renergy == pytest.approx(p2.renergy)
    assert p1.rbasis.shape == (3, 2) and p1.renergy > 0.9
    pe = O.POD(X, min_renergy=0.1, min_nmodes=1, max_nmodes=2)
    O.reduce(pe, O.SVD())
    assert pe.nmodes == 1 and pe.renergy > 0.1
    pt = O.POD(X, 2)
    O.reduce(pt, O.TSVD())
    assert pt.rbasis.shape == (3, 2) and pt.renergy > 0.7


def test_deim_golden_random(golden):
    for case in golden["deim_random"]:
        rng = np.random.Generator(np.random.PCG64(case["seed"]))
        Q, _ = np.linalg.qr(rng.standard_normal((case["m"], case["k"])))
        assert O.deim_interpolation_indices(Q).tolist() == case["indices"]


def test_deim_ties_and_nan():
    # exact ties: duplicated rows with exactly representable entries -> first index wins
    B = np.zeros((8, 2))
    B[:, 0] = [1, -2, 2, 0.5, 2, -2, 1, 0]      # |.| max = 2 first at row 2 (1-based)
    B[:, 1] = [0, 1, 1, 4, -3, 4, 4, 4]
    idx = O.deim_interpolation_indices(B)
    assert idx[0] == 2
    # second step: c = B[1,1]/B[1,0] = -0.5 ; r = |u2 + 0.5 u1| -> [.5,0,2,4.25,2,3,4.5,4] -> row 7
    assert idx[1] == 7
    Bn = B.copy(); Bn[5, 0] = np.nan; Bn[6, 0] = np.nan
    assert O.deim_interpolation_indices(Bn)[0] == 6      # first NaN is maximal
    # l=2 uses the 1x1 divide branch of the generic backslash; singular -> exception
    Z = np.zeros((4, 2)); Z[:, 1] = 1
    with pytest.raises(np.linalg.LinAlgError):
        O.deim_interpolation_indices(Z)


def test_sharded_restatements_match_unsharded():
    rng = np.random.Generator(np.random.PCG64(7))
    Q, _ = np.linalg.qr(rng.standard_normal((301, 12)))
    Q[150] = Q[17]                                   # exact tie across a shard boundary
    full = O.deim_interpolation_indices(Q)
    bounds = [0, 100, 100, 250, 301]                 # includes an empty shard
    shards = [(Q[a:b], a) for a, b in zip(bounds[:-1], bounds[1:])]
    # lock-step emulation of the all-gather: run the ranks as coroutines over shared state
    import threading
    n = len(shards)
    barrier = threading.Barrier(n)
    slots = [None] * n
    out = [None] * n

    def run(r):
        def allgather(x):
            slots[r] = x
            barrier.wait()
            res = list(slots)
            barrier.wait()
            return res
        out[r] = O.deim_interpolation_indices_sharded(shards[r][0], shards[r][1], allgather)
    ts = [threading.Thread(target=run, args=(r,)) for r in range(n)]
    [t.start() for t in ts]; [t.join() for t in ts]
    for r in range(n):
        assert out[r].tolist() == full.tolist()


def test_parity_metrics():
    s = np.array([1.0, 1e-3, 1e-9])
    assert O.sigma_error(s, s * (1 + 1e-12)) < 2e-12
    assert O.sigma_error(s, s + np.array([0, 0, 1e-12])) == pytest.approx(1e-12)
    Q, _ = np.linalg.qr(np.random.default_rng(0).standard_normal((50, 4)))
    assert O.subspace_sine(Q, -Q[:, ::-1]) < 1e-14


def test_burgers_closed_form_solves_the_pde():
    """BASELINE config 5 input: the Cole-Hopf field the MOR_SYNTH_BURGERS generator
    evaluates must satisfy u_t + u u_x + v u_y = nu lap(u) (central differences, nu large enough for FD)."""
    nu, h, t = 0.02, 1e-4, 0.37
    x, y = np.meshgrid(np.linspace(0.05, 0.95, 41), np.linspace(0.05, 0.95, 37))
    f = lambda dx=0.0, dy=0.0, dt=0.0: O.burgers_field(x + dx, y + dy, t + dt, nu, seed=5)
    for comp in (0, 1):
        w = f()[comp]
        u, v = f()
        wx = (f(h)[comp] - f(-h)[comp]) / (2 * h)
        wy = (f(0, h)[comp] - f(0, -h)[comp]) / (2 * h)
        wt = (f(0, 0, h)[comp] - f(0, 0, -h)[comp]) / (2 * h)
        lap = (f(h)[comp] + f(-h)[comp] + f(0, h)[comp] + f(0, -h)[comp] - 4 * w) / h**2
        assert np.abs(wt + u * wx + v * wy - nu * lap).max() < 1e-4 * np.abs(u * wx).max()


def test_burgers_snapshots_layout_and_shard_invariance():
    nx, n, nu = 48, 12, 2e-3
    X = O.burgers_snapshots(nx * 40, n, nx, nu, seed=3)
    assert X.shape == (nx * 40, n) and X.flags.f_contiguous
    # (numpy's SIMD exp may differ in the last bit between the body and the tail of an array)
    assert np.allclose(X[777:1500], O.burgers_snapshots(723, n, nx, nu, seed=3, row0=777), rtol=0, atol=1e-14)
    p, _, d = O.burgers_states(3)
    assert p.min() - 1e-12 <= X.min() and X.max() <= p.max() + 1e-12      # a convex combination of the states
    assert len({round(v, 6) for v in X[:, 0]} & {round(v, 6) for v in p}) == 3   # all three plateaus are in the domain
    assert np.all(np.abs(d - np.array([0.0, -0.45, -0.55])) <= 0.01 + 1e-15)
    # the fronts travel: snapshots are linearly independent (slow singular value decay)
    s = sla.svd(X, compute_uv=False)
    assert s[-1] / s[0] > 1e-6


def test_blocked_deim_restatement_equals_the_reference_formulation(golden):
    """The blocked left-looking-LU form of DEIM that the library runs (panels of 16, sub-panels of 4) picks the
    same points as deim.jl:9-28 restated line for line -- on the golden cases, the sin fixture, exact ties,
    the singular fixture and the Burgers POD basis."""
    for case in golden["deim_random"]:
        rng = np.random.Generator(np.random.PCG64(case["seed"]))
        Q, _ = np.linalg.qr(rng.standard_normal((case["m"], case["k"])))
        assert O.deim_interpolation_indices_blocked(Q).tolist() == case["indices"]
        assert O.deim_interpolation_indices_blocked(Q, block=8, sub=2).tolist() == case["indices"]
    U, _, _ = O._svd(sin_fixture())
    assert O.deim_interpolation_indices_blocked(U, block=4, sub=2).tolist() == golden["F2"]["deim_U10"]
    rng = np.random.Generator(np.random.PCG64(77))
    A = rng.standard_normal((3000, 40)); A[2000:3000] = A[500:1500]          # exact ties: the lower index wins
    want = O.deim_interpolation_indices(A)
    assert O.deim_interpolation_indices_blocked(A).tolist() == want.tolist() and not np.any(want > 2000)
    Z = np.zeros((4, 3)); Z[:, 1] = 1; Z[:, 2] = [1, 2, 3, 4]
    with pytest.raises(np.linalg.LinAlgError):
        O.deim_interpolation_indices_blocked(Z)
    S = O.burgers_snapshots(96 * 96, 40, 96, 2e-3, seed=1)
    Ub = O._svd(S)[0]
    assert O.deim_interpolation_indices_blocked(Ub).tolist() == O.deim_interpolation_indices(Ub).tolist()
