"""POD parity on the GPU through the C ABI (ctypes host mirror == the Julia shim's calls):
the reference's own test assertions (test/interface.jl, test/DataReduction.jl, doctests)
plus value-level parity with the oracle: singular values <= 1e-10 relative (absolute
1e-10*s1 below 1e-5*s1), leading subspace sine <= 1e-8."""
import numpy as np
import pytest
import scipy.linalg as sla

import mor_b200
import mor_oracle as O
from conftest import sin_fixture
from mor_b200 import _lib
from mor_b200.device import DeviceMatrix, pod_factor_dev
from mor_b200.pod import _factor

pytestmark = pytest.mark.gpu

SIGMA_TOL = 1e-10
ANGLE_TOL = 1e-8


def graded(m, n, decades, seed):
    rng = np.random.default_rng(seed)
    return np.asfortranarray(rng.standard_normal((m, n)) * np.logspace(0, -decades, n))


def check_against_oracle(X, k, alg_b, alg_o, omega=None, vov=False):
    snaps_b = [np.ascontiguousarray(X[:, j]) for j in range(X.shape[1])] if vov else X
    pb = mor_b200.POD(snaps_b, k); po = O.POD(X, k)
    assert mor_b200.reduce(pb, alg_b, omega=omega) is None
    O.reduce(po, alg_o, omega=omega)
    assert pb.nmodes == po.nmodes
    assert pb.rbasis.shape == po.rbasis.shape and pb.spectrum.shape == po.spectrum.shape
    assert O.sigma_error(po.spectrum, pb.spectrum) <= SIGMA_TOL
    assert pb.renergy == pytest.approx(po.renergy, rel=1e-9)
    assert O.subspace_sine(po.rbasis, pb.rbasis) <= ANGLE_TOL
    assert np.max(np.abs(pb.rbasis.T @ pb.rbasis - np.eye(pb.rbasis.shape[1]))) < 1e-9
    return pb, po


def test_interface_fixture(gpu, golden):
    """test/interface.jl:3-17 on [3 0;0 1] (Matrix and vector of vectors) x {SVD,TSVD,RSVD}."""
    X = np.array([[3.0, 0.0], [0.0, 1.0]], order="F")
    vov = [np.array([3.0, 0.0]), np.array([0.0, 1.0])]
    for snaps in (X, vov):
        for name, alg in (("svd", mor_b200.SVD()), ("tsvd", mor_b200.TSVD()), ("rsvd", mor_b200.RSVD())):
            pod = mor_b200.POD(snaps, 1)
            kw = dict(omega=np.array([[1.0], [0.5]])) if name == "rsvd" else {}
            assert mor_b200.reduce(pod, alg, **kw) is None
            assert pod.rbasis.shape == (2, 1)
            assert len(pod.spectrum) >= pod.nmodes
            assert pod.nmodes == 1
            assert 0.0 < pod.renergy <= 1.0
            g = golden["F1"][name]
            np.testing.assert_allclose(pod.spectrum, g["spectrum"], rtol=1e-13, atol=1e-15)
            assert pod.renergy == pytest.approx(g["renergy"], rel=1e-13)
            np.testing.assert_allclose(np.abs(pod.rbasis), g["rbasis_abs"], atol=1e-13)


def test_sin_fixture_goldens(gpu, golden):
    """src/precompile.jl fixture sin(i*j/n), 20 x 10, kappa = 2e10: needs the shifted path."""
    S = sin_fixture()
    pod = mor_b200.POD(S, 3)
    mor_b200.reduce(pod, mor_b200.SVD())
    s_ref = np.array(golden["F2"]["spectrum"])
    assert O.sigma_error(s_ref, pod.spectrum) <= SIGMA_TOL
    assert pod.renergy == pytest.approx(golden["F2"]["renergy_svd_k3"], rel=1e-10)
    pod = mor_b200.POD(S, 3)
    mor_b200.reduce(pod, mor_b200.TSVD())
    assert pod.spectrum.shape == (3,)
    assert pod.renergy == pytest.approx(golden["F2"]["renergy_tsvd_k3"], rel=1e-10)
    check_against_oracle(S, 3, mor_b200.SVD(), O.SVD())
    # downstream: DEIM on the library's own basis gives the oracle's indices
    U, _, _ = O._svd(S)
    pod = mor_b200.POD(S, 5); mor_b200.reduce(pod, mor_b200.SVD())
    assert mor_b200.deim_interpolation_indices(pod.rbasis).tolist() == \
        O.deim_interpolation_indices(U[:, :5]).tolist()


@pytest.mark.parametrize("m,n,k,dec,seed", [(200, 10, 3, 1, 0), (5000, 64, 8, 3, 1), (20_000, 65, 16, 3, 2),
                                            (100_000, 256, 32, 3, 3), (30_011, 384, 20, 2, 4), (262_144, 256, 32, 3, 5)])
def test_svd_tall_vs_oracle(gpu, m, n, k, dec, seed):
    X = graded(m, n, dec, seed)
    pb, _ = check_against_oracle(X, k, mor_b200.SVD(), O.SVD())
    assert pb.spectrum.shape == (n,)                      # SVD(): FULL spectrum, POD.jl:164


def test_tsvd_and_vov_and_energy_form(gpu):
    X = graded(8000, 40, 2, 7)
    check_against_oracle(X, 5, mor_b200.TSVD(), O.TSVD())
    check_against_oracle(X, 5, mor_b200.SVD(), O.SVD(), vov=True)      # test/DataReduction.jl:38-39
    vov = [np.ascontiguousarray(X[:, j]) for j in range(X.shape[1])]
    pb = mor_b200.POD(vov, min_renergy=0.3, min_nmodes=1, max_nmodes=10)
    po = O.POD(vov, min_renergy=0.3, min_nmodes=1, max_nmodes=10)
    mor_b200.reduce(pb, mor_b200.SVD()); O.reduce(po, O.SVD())
    assert pb.nmodes == po.nmodes and pb.renergy == pytest.approx(po.renergy, rel=1e-10)
    assert pb.rbasis.shape == (8000, pb.nmodes)


def test_wide_lorenz_like(gpu):
    """test/DataReduction.jl: 3 x nt Lorenz snapshots (wide), k = 2, all three backends."""
    from scipy.integrate import solve_ivp
    sol = solve_ivp(lambda t, u: [10 * (u[1] - u[0]), u[0] * (28 - u[2]) - u[1], u[0] * u[1] - 8 / 3 * u[2]],
                    (0, 100), [1.0, 0.0, 0.0], rtol=1e-8, atol=1e-8)
    X = np.asfortranarray(sol.y)
    assert X.shape[0] == 3 and X.shape[1] > 1000
    pb, po = check_against_oracle(X, 2, mor_b200.SVD(), O.SVD())
    assert pb.rbasis.shape == (3, 2) and pb.renergy > 0.9            # DataReduction.jl:41-43
    pb, _ = check_against_oracle(X, 2, mor_b200.TSVD(), O.TSVD())
    assert pb.renergy > 0.7                                           # DataReduction.jl:55-57
    omega = np.random.default_rng(3).standard_normal((X.shape[1], 2))
    pb, _ = check_against_oracle(X, 2, mor_b200.RSVD(), O.RSVD(), omega=omega)
    assert pb.renergy > 0.7                                           # DataReduction.jl:64-66
    pod = mor_b200.POD(X, min_renergy=0.1, min_nmodes=1, max_nmodes=2)
    mor_b200.reduce(pod, mor_b200.SVD())
    assert pod.nmodes == 1 and pod.renergy > 0.1                      # DataReduction.jl:45-48


@pytest.mark.parametrize("m,n,k,p", [(4000, 50, 6, 0), (4000, 50, 6, 4), (60_000, 200, 16, 0), (50_001, 512, 64, 0)])
def test_rsvd_shared_omega(gpu, m, n, k, p):
    rng = np.random.default_rng(10 + p)
    X = graded(m, n, 3, 20 + p)
    X[:, :k] *= 30.0                                   # a dominant rank-k part, like config 4
    omega = rng.standard_normal((n, k + p))
    pb, po = check_against_oracle(X, k, mor_b200.RSVD(p), O.RSVD(p), omega=omega)
    assert pb.spectrum.shape == (k,)
    # library-drawn Omega (seeded Philox): same subspace quality, deterministic
    p1 = mor_b200.POD(X, k); mor_b200.reduce(p1, mor_b200.RSVD(p), seed=5)
    p2 = mor_b200.POD(X, k); mor_b200.reduce(p2, mor_b200.RSVD(p), seed=5)
    assert np.array_equal(p1.spectrum, p2.spectrum) and np.array_equal(p1.rbasis, p2.rbasis)
    # a different Omega gives a different sketch (p = 0, no power iterations: no tight bound),
    # but sigma_i(Q'X) <= sigma_i(X) always and the dominant value is captured
    s_true = sla.svd(X, compute_uv=False, lapack_driver="gesdd")[:k]
    assert np.all(p1.spectrum <= s_true * (1 + 1e-12)) and p1.spectrum[0] >= 0.5 * s_true[0]
    assert np.max(np.abs(p1.rbasis.T @ p1.rbasis - np.eye(k))) < 1e-9


@pytest.mark.parametrize("kappa", [1e5, 1e8, 1e12])
def test_ill_conditioned_takes_more_passes(gpu, kappa):
    """Correlated columns (not a column scaling): CholeskyQR2 / shifted CholeskyQR3 path."""
    rng = np.random.default_rng(0)
    m, n = 20_000, 48
    U, _ = np.linalg.qr(rng.standard_normal((m, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    X = np.asfortranarray((U * np.logspace(0, -np.log10(kappa), n)) @ V.T)
    f = _factor(X, _lib.MOR_SVD, 8)
    passes, sweeps = f.info()
    assert passes >= 2
    s_ref = sla.svd(X, compute_uv=False, lapack_driver="gesdd")
    assert O.sigma_error(s_ref, f.s) <= SIGMA_TOL
    Uk = f.basis(8)
    Uo = O._svd(X)[0][:, :8]
    assert O.subspace_sine(Uo, Uk) <= ANGLE_TOL
    f.close()


def test_rank_deficient_and_zero_columns(gpu):
    rng = np.random.default_rng(1)
    X = np.asfortranarray(rng.standard_normal((3000, 12)))
    X[:, 5] = 0.0                      # an all-zero snapshot
    X[:, 9] = X[:, 2]                  # a repeated snapshot
    pb, po = check_against_oracle(X, 4, mor_b200.SVD(), O.SVD())
    assert pb.spectrum[-1] <= 1e-10 * pb.spectrum[0] and pb.spectrum[-2] <= 1e-10 * pb.spectrum[0]


def test_nonfinite_input_is_an_argument_error(gpu):
    X = graded(1000, 8, 1, 3); X[17, 3] = np.nan
    with pytest.raises(ValueError):
        mor_b200.reduce(mor_b200.POD(X, 2), mor_b200.SVD())
    X[17, 3] = np.inf
    with pytest.raises(ValueError):
        mor_b200.reduce(mor_b200.POD(X, 2), mor_b200.RSVD())


def test_right_vectors_and_sign_convention(gpu):
    X = graded(5000, 20, 2, 11)
    f = _factor(X, _lib.MOR_SVD, 20)
    U, V = f.basis(20, want_v=True)
    assert np.linalg.norm(X - (U * f.s) @ V.T) <= 1e-12 * f.s[0]
    j = np.argmax(np.abs(V), axis=0)
    assert np.all(V[j, np.arange(20)] > 0)            # largest-|entry| of every right vector positive
    f.close()


def test_device_resident_synthetic_config2_shape(gpu):
    """Device-generated graded matrix (the bench generator): downloaded copy vs oracle, and
    orthonormality / reconstruction of the device-side basis."""
    m, n, k = 524_288, 256, 32
    X = DeviceMatrix.alloc(m, n).synth(_lib.SYNTH_GRADED, seed=2, param=3.0)
    f = pod_factor_dev(X, _lib.MOR_SVD, k)
    passes, sweeps = f.info()
    assert passes == 1                                  # column-scaled Gaussian: one Gram pass suffices
    Xh = X.to_host()
    s_ref = sla.svd(Xh, compute_uv=False, lapack_driver="gesdd")
    assert O.sigma_error(s_ref, f.s) <= SIGMA_TOL
    U = f.basis(k, DeviceMatrix.alloc(m, k)).to_host()
    assert O.subspace_sine(O._svd(Xh)[0][:, :k], U) <= ANGLE_TOL
    # forcing the two-pass path gives the same answer
    f2 = pod_factor_dev(X, _lib.MOR_SVD, k, flags=_lib.POD_FORCE_TWO_PASS)
    assert f2.info()[0] == 2
    assert O.sigma_error(s_ref, f2.s) <= SIGMA_TOL
    assert np.array_equal(X.to_host(), Xh)              # without MAY_OVERWRITE the input is untouched
    t = _lib.last_timings()
    assert t["gram"] > 0 and t["total"] >= t["gram"]


def test_pipelined_host_ingest_chunks(gpu, monkeypatch):
    """Host-pointer ingest streams row chunks on a copy stream while the Gram matrix of the chunks
    that already landed accumulates; forced to 13 ragged chunks here.  Same result as one shot."""
    X = graded(50_003, 96, 3, 21)
    monkeypatch.setenv("MOR_B200_INGEST_CHUNK_ROWS", "4000")
    pb, _ = check_against_oracle(X, 10, mor_b200.SVD(), O.SVD())
    big = np.zeros((50_100, 96), order="F"); big[:50_003] = X          # ld > m view
    f = _factor(big[:50_003], _lib.MOR_TSVD, 10)
    assert O.sigma_error(pb.spectrum[:10], f.s) <= 1e-12
    f.close()


def test_fitzhugh_nagumo_pod5_deim5(gpu):
    """BASELINE config 1: FitzHugh-Nagumo MOL N=15 (32-dim), POD5-DEIM5 as deim(sys, snapshot, 5;
    deim_dim=5) does numerically (deim.jl:223-225,245-249,147): TSVD POD of the 32 x nt snapshots,
    TSVD POD of the nonlinear snapshots, DEIM indices of the latter -- wide matrices throughout."""
    import fhn_fixture as F
    X = F.snapshots()
    assert X.shape == (32, 1401)
    pb, po = check_against_oracle(X, 5, mor_b200.TSVD(), O.TSVD())          # V = pod_reducer.rbasis
    assert pb.rbasis.shape == (32, 5)
    NL = F.nonlinear_snapshots(X)[:16]                                        # the 16 v-equations carry f(v)
    ub, uo = check_against_oracle(NL, 5, mor_b200.TSVD(), O.TSVD())          # U = deim_reducer.rbasis
    idx = mor_b200.deim_interpolation_indices(ub.rbasis)
    assert idx.tolist() == O.deim_interpolation_indices(ub.rbasis).tolist()
    assert idx.tolist() == O.deim_interpolation_indices(uo.rbasis * np.sign((uo.rbasis * ub.rbasis).sum(0))).tolist()
    assert len(set(idx.tolist())) == 5 and 1 <= idx.min() and idx.max() <= 16


def test_known_answer_matrix(gpu):
    """X = H_u [S V'; 0] generated on the device (Householder reflectors: singular values and left
    vectors known in closed form at ANY size -- SURVEY 8d 'KAT').  First the generator itself against
    LAPACK at a small size, then the library at 2M rows, where no CPU oracle is practical."""
    n = 24
    Xs = DeviceMatrix.alloc(1001, n).synth(_lib.SYNTH_KAT, seed=9, param=3.0)
    s_exact = 10.0 ** (-3.0 * np.arange(n) / (n - 1))
    assert np.max(np.abs(sla.svd(Xs.to_host(), compute_uv=False) - s_exact)) < 1e-14
    m, n, k = 2_000_003, 256, 32
    X = DeviceMatrix.alloc(m, n).synth(_lib.SYNTH_KAT, seed=9, param=3.0)
    f = pod_factor_dev(X, _lib.MOR_SVD, k, flags=_lib.POD_MAY_OVERWRITE)
    s_exact = 10.0 ** (-3.0 * np.arange(n) / (n - 1))
    assert f.info()[0] >= 2                              # not a column scaling: CholeskyQR2 path
    assert O.sigma_error(s_exact, f.s) <= SIGMA_TOL
    U = f.basis(k, DeviceMatrix.alloc(m, k))
    assert U.kat_check(k, n, seed=9, param=3.0) <= ANGLE_TOL


def test_abi_error_paths(gpu):
    """Status codes of the C ABI (include/mor_b200.h): bad arguments are reported, never crash."""
    import ctypes as C
    lib = gpu
    X = graded(500, 6, 1, 0)
    s = np.empty(6); nS = C.c_int64(0); h = C.c_void_p()
    args = (X.ctypes.data, 500, 6, 500)
    assert lib.mor_b200_pod_factor(*args, _lib.MOR_TSVD, 0, 0, None, 0, C.byref(h), s.ctypes.data, C.byref(nS)) == _lib.ERR_ARG
    assert lib.mor_b200_pod_factor(*args, _lib.MOR_TSVD, 7, 0, None, 0, C.byref(h), s.ctypes.data, C.byref(nS)) == _lib.ERR_ARG
    assert lib.mor_b200_pod_factor(*args, _lib.MOR_RSVD, 2, -1, None, 0, C.byref(h), s.ctypes.data, C.byref(nS)) == _lib.ERR_ARG
    assert lib.mor_b200_pod_factor(*args, 7, 2, 0, None, 0, C.byref(h), s.ctypes.data, C.byref(nS)) == _lib.ERR_ARG
    assert lib.mor_b200_pod_factor(None, 500, 6, 500, _lib.MOR_SVD, 2, 0, None, 0, C.byref(h), s.ctypes.data, C.byref(nS)) == _lib.ERR_ARG
    assert lib.mor_b200_pod_factor(X.ctypes.data, 500, 6, 499, _lib.MOR_SVD, 2, 0, None, 0, C.byref(h), s.ctypes.data, C.byref(nS)) == _lib.ERR_ARG
    assert b"" != lib.mor_b200_last_error()
    f = _factor(X, _lib.MOR_TSVD, 3)
    U = np.empty((500, 6), order="F")
    assert lib.mor_b200_pod_basis(f.handle, 7, U.ctypes.data, 500, None, 0) == _lib.ERR_ARG     # only 6 triplets exist
    assert lib.mor_b200_pod_basis(f.handle, 3, U.ctypes.data, 499, None, 0) == _lib.ERR_ARG     # ldu < m
    assert lib.mor_b200_pod_basis(f.handle, 3, U.ctypes.data, 500, None, 0) == _lib.OK
    f.close()
    assert lib.mor_b200_pod_free(None) == _lib.OK
    # a borrowed device matrix with an odd leading dimension cannot feed the TMA kernels: reported
    d = DeviceMatrix.alloc(11, 4)
    _, _, _, ptr = d.info
    w = DeviceMatrix.wrap(ptr, 5, 4, 11)
    with pytest.raises(ValueError):
        pod_factor_dev(w, _lib.MOR_SVD, 2)
    # a wide matrix and an empty one
    with pytest.raises(ValueError):
        mor_b200.deim_interpolation_indices(np.zeros((4, 0), order="F"))
    idx = np.zeros(1025, dtype=np.int64)
    B = np.zeros((2000, 1025), order="F")
    assert lib.mor_b200_deim_indices(B.ctypes.data, 2000, 1025, 2000, idx.ctypes.data) == _lib.ERR_ARG


def test_mpmath_goldens(gpu, mp_golden):
    """The independent pin: singular values / energies / leading subspace of the reference's sin(i*j/n) fixture
    (src/precompile.jl:5-12) against 50-digit mpmath values (oracle/make_golden_mpmath.py)."""
    f2 = mp_golden["F2"]
    S = sin_fixture()
    pod = mor_b200.POD(S, 3)
    mor_b200.reduce(pod, mor_b200.SVD())
    assert O.sigma_error(np.array(f2["sigma"]), pod.spectrum) <= SIGMA_TOL
    assert pod.renergy == pytest.approx(f2["renergy_svd_k3"], rel=1e-10)
    assert O.subspace_sine(f2["U"][:, :3], pod.rbasis) <= ANGLE_TOL
    pod = mor_b200.POD(S, 3)
    mor_b200.reduce(pod, mor_b200.TSVD())
    assert pod.renergy == pytest.approx(f2["renergy_tsvd_k3"], rel=1e-10)
    np.testing.assert_allclose(pod.spectrum, f2["sigma"][:3], rtol=1e-12)


def test_config2_exact_through_host_abi(gpu):
    """BASELINE config 2 exactly: synthetic 1,048,576 x 256 FP64 snapshot matrix (N(0,1) * 10^(-3j/(n-1)), Philox
    seed 2), reduce!(POD(X, 32), SVD()) through the host-pointer ABI (what the Julia shim calls), against dgesdd on
    the same matrix: all 256 singular values, the 32 modes, the energy."""
    m, n, k = 1_048_576, 256, 32
    Xh = DeviceMatrix.alloc(m, n).synth(_lib.SYNTH_GRADED, seed=2, param=3.0).to_host()
    pod = mor_b200.POD(Xh, k)
    assert mor_b200.reduce(pod, mor_b200.SVD()) is None
    u_ref, s_ref, _ = O._svd(Xh)
    assert pod.spectrum.shape == (n,) and pod.rbasis.shape == (m, k) and pod.nmodes == k
    assert O.sigma_error(s_ref, pod.spectrum) <= SIGMA_TOL
    assert O.subspace_sine(u_ref[:, :k], pod.rbasis) <= ANGLE_TOL
    ref = O.POD(Xh, k)
    ref.nmodes, ref.renergy = O.determine_truncation(s_ref, k, k, 0.0)
    assert pod.renergy == pytest.approx(ref.renergy, rel=1e-10)
    assert np.abs(pod.rbasis.T @ pod.rbasis - np.eye(k)).max() <= 1e-11
