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


def test_deim_projector_vs_oracle(gpu):
    """deim.jl:150: projector = ((U[idx,:])' \\ (U'V))' -- next row of the hot path."""
    import fhn_fixture as F
    rng = np.random.Generator(np.random.PCG64(77))
    for (n, kd, kp) in [(40, 3, 5), (5000, 24, 10), (120_001, 64, 64), (30_000, 130, 70)]:
        U, _ = np.linalg.qr(rng.standard_normal((n, kd)))
        V, _ = np.linalg.qr(rng.standard_normal((n, kp)))
        idx = O.deim_interpolation_indices(U)
        P = mor_b200.deim_projector(U, V, idx)
        Pref = O.deim_projector(U, V, idx)
        assert P.shape == (kp, kd)
        assert np.max(np.abs(P - Pref)) <= 1e-10 * np.max(np.abs(Pref))
    # the README example end to end: POD5 / DEIM5 of FitzHugh-Nagumo, the v-block of the state basis
    X = F.snapshots()
    pv = mor_b200.POD(X[:16], 5); mor_b200.reduce(pv, mor_b200.TSVD())
    pu = mor_b200.POD(F.nonlinear_snapshots(X)[:16], 5); mor_b200.reduce(pu, mor_b200.TSVD())
    idx = mor_b200.deim_interpolation_indices(pu.rbasis)
    P = mor_b200.deim_projector(pu.rbasis, pv.rbasis, idx)
    Pref = O.deim_projector(pu.rbasis, pv.rbasis, idx)
    assert np.max(np.abs(P - Pref)) <= 1e-10 * np.max(np.abs(Pref))
    # a singular selection is reported like the reference's SingularException
    with pytest.raises(mor_b200.SingularError):
        mor_b200.deim_projector(U, V, np.full(U.shape[1], 1, dtype=np.int64))


def test_galerkin_projection_of_the_linear_part(gpu):
    """deim.jl:154-155,261: A_hat = V'AV (A sparse CSC), g_hat = V'g, y_hat(0) = V'snapshot[:,1]."""
    import scipy.sparse as sp
    import fhn_fixture as F
    rng = np.random.Generator(np.random.PCG64(5))
    for (m, k, dens) in [(32, 5, 0.2), (2000, 12, 0.01), (150_001, 40, 2e-5)]:
        nnz = int(dens * m * m)        # scattered entries + a second-difference stencil (a MOL operator)
        A = sp.coo_matrix((rng.standard_normal(nnz), (rng.integers(0, m, nnz), rng.integers(0, m, nnz))),
                          shape=(m, m)).tocsc() + \
            sp.diags([np.full(m - 1, 1.0), np.full(m, -2.0), np.full(m - 1, 1.0)], [-1, 0, 1], format="csc")
        V, _ = np.linalg.qr(rng.standard_normal((m, k)))
        ref = V.T @ (A @ V)
        got = mor_b200.galerkin_project(A, V)
        assert got.shape == (k, k)
        assert np.max(np.abs(got - ref)) <= 1e-12 * max(np.max(np.abs(ref)), 1.0)
        g = rng.standard_normal(m)
        assert np.max(np.abs(mor_b200.project(V, g) - V.T @ g)) <= 1e-12 * np.linalg.norm(g)
        G2 = rng.standard_normal((m, 3))
        assert np.max(np.abs(mor_b200.project(V, G2) - V.T @ G2)) <= 1e-12 * np.linalg.norm(G2)
    X = F.snapshots()
    pod = mor_b200.POD(X, 5); mor_b200.reduce(pod, mor_b200.TSVD())
    y0 = mor_b200.project(pod.rbasis, X[:, 0])                       # deim.jl:261
    assert np.max(np.abs(y0 - pod.rbasis.T @ X[:, 0])) <= 1e-13 * max(np.linalg.norm(X[:, 0]), 1.0)
    with pytest.raises(ValueError):
        mor_b200.galerkin_project(sp.eye(7, format="csc"), np.ones((8, 2)))


def test_calls_from_several_host_threads_serialise(gpu):
    """The library is one context per process; concurrent calls from several host threads (ctypes
    drops the GIL, like Julia tasks on several threads) must serialise on it and all be correct."""
    import threading
    rng = np.random.Generator(np.random.PCG64(3))
    mats = [np.asfortranarray(np.linalg.qr(rng.standard_normal((20_000 + 37 * i, 12 + i)))[0]) for i in range(6)]
    want = [O.deim_interpolation_indices(Q).tolist() for Q in mats]
    got = [None] * len(mats)
    pods = [None] * len(mats)

    def work(i):
        got[i] = mor_b200.deim_interpolation_indices(mats[i]).tolist()
        p = mor_b200.POD(mats[i], 3)
        mor_b200.reduce(p, mor_b200.SVD())
        pods[i] = p.spectrum
    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(mats))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert got == want
    for s in pods:                      # orthonormal columns: all singular values are 1
        assert np.max(np.abs(s - 1.0)) < 1e-12


def test_burgers_snapshots_generator_vs_closed_form(gpu):
    """MOR_SYNTH_BURGERS (BASELINE config 5 input) against the oracle's restatement of the same Cole-Hopf
    closed form: entry by entry, on ragged shards (row0 not a multiple of the grid width)."""
    nx, n, nu = 96, 37, 1.5e-3
    m = nx * 80 + 13
    ref = O.burgers_snapshots(m, n, nx, nu, seed=9)
    X = DeviceMatrix.alloc(m, n).synth(_lib.SYNTH_BURGERS, seed=9, param=nu, iparam=nx).to_host()
    assert np.abs(X - ref).max() <= 1e-11
    for row0, rows in ((0, 1), (101, 3001), (m - 7, 7)):
        part = DeviceMatrix.alloc(rows, n).synth(_lib.SYNTH_BURGERS, seed=9, global_row0=row0, param=nu, iparam=nx)
        assert np.array_equal(part.to_host(), X[row0:row0 + rows])     # shard invariant, bit for bit
    with pytest.raises(ValueError):
        DeviceMatrix.alloc(8, 2).synth(_lib.SYNTH_BURGERS, seed=0, param=0.0, iparam=4)


def test_burgers_pod_basis_deim_config5(gpu):
    """Config 5 end to end at oracle size: Burgers snapshots -> POD basis (library) -> DEIM indices,
    indices bit-exact against the oracle on the same basis; sigma / subspace against dgesdd."""
    from mor_b200.device import pod_factor_dev
    nx, n, k = 256, 64, 64
    m = nx * nx
    S = DeviceMatrix.alloc(m, n).synth(_lib.SYNTH_BURGERS, seed=1, param=1e-3, iparam=nx)
    Sh = S.to_host()
    f = pod_factor_dev(S, _lib.MOR_SVD, k)
    u_ref, s_ref, _ = O._svd(Sh)
    assert O.sigma_error(s_ref, f.s) <= 1e-10
    U = f.basis(k, DeviceMatrix.alloc(m, k))
    Uh = U.to_host()
    assert O.subspace_sine(u_ref[:, :16], Uh[:, :16]) <= 1e-8
    assert np.abs(Uh.T @ Uh - np.eye(k)).max() <= 1e-10
    idx = deim_indices_dev(U)
    assert idx.tolist() == O.deim_interpolation_indices(Uh).tolist()
    assert len(set(idx.tolist())) == k
    # DEIM points of a front-dominated field sit on the swept region, not on the plateaus
    swept = np.ptp(Sh, axis=1) > 0.1
    assert swept[idx - 1].mean() > 0.9


@pytest.mark.parametrize("m,k,seed", [(4097, 33, 1), (50_000, 17, 2), (70_001, 40, 3), (131_072, 64, 4), (20_000, 130, 5)])
def test_blocked_and_streaming_paths_agree(gpu, monkeypatch, m, k, seed):
    """k > 16 takes the blocked path (16-column residual panels, one pass over the earlier columns per block);
    MOR_B200_DEIM_BLOCK=0 forces the streaming path.  Both must give the oracle's indices."""
    rng = np.random.Generator(np.random.PCG64(500 + seed))
    Q, _ = np.linalg.qr(rng.standard_normal((m, k)))
    Q = np.asfortranarray(Q)
    want = O.deim_interpolation_indices(Q).tolist()
    monkeypatch.delenv("MOR_B200_DEIM_BLOCK", raising=False)
    assert mor_b200.deim_interpolation_indices(Q).tolist() == want
    assert _lib.last_timings()["deim_blocked"] == 1.0 and _lib.last_timings()["deim_panel"] > 0
    monkeypatch.setenv("MOR_B200_DEIM_BLOCK", "0")
    assert mor_b200.deim_interpolation_indices(Q).tolist() == want
    assert _lib.last_timings()["deim_blocked"] == 0.0


def test_blocked_path_exact_ties_across_blocks(gpu, monkeypatch):
    """Duplicated rows tie exactly in every residual; the lowest global index must win in every block."""
    monkeypatch.delenv("MOR_B200_DEIM_BLOCK", raising=False)
    rng = np.random.Generator(np.random.PCG64(77))
    m, k = 30_000, 40
    A = rng.standard_normal((m, k))
    A[20_000:30_000] = A[5_000:15_000]            # every row of the second range duplicates one of the first
    idx = mor_b200.deim_interpolation_indices(np.asfortranarray(A))
    assert idx.tolist() == O.deim_interpolation_indices(A).tolist()
    assert not np.any((idx > 20_000))             # the duplicate with the higher index never wins


def test_mpmath_goldens(gpu, mp_golden):
    """The independent pin: indices from the 50-digit evaluation of deim.jl:13-24 (oracle/make_golden_mpmath.py),
    every step decided by a margin > 1e-3."""
    f2 = mp_golden["F2"]
    for basis, chain in ((f2["U"], f2["deim_U10"]), (f2["Q5"], f2["deim_Q5"])) + \
            tuple((c["basis"], c["chain"]) for c in mp_golden["deim_random"]):
        assert mor_b200.deim_interpolation_indices(np.asfortranarray(basis)).tolist() == chain["indices"]


def _first_difference_report(got, want, margins):
    t = int(np.flatnonzero(np.asarray(got) != np.asarray(want))[0])
    return f"first difference at step {t + 1}: got row {got[t]}, oracle row {want[t]}, oracle margin {margins[t]:.3e}"


@pytest.mark.parametrize("kind", ["gauss", "burgers"])
def test_headline_k256_vs_oracle(gpu, monkeypatch, kind):
    """k = 256, the headline DEIM dimension (BASELINE config 5), at the largest size the oracle finishes in
    seconds (262,144 rows): the blocked path, the streaming path and the host-pointer entry (pipelined column
    groups, deferred columns) against deim.jl:9-28 restated, Int64 equality."""
    from mor_b200.device import orthonormalize_dev, pod_factor_dev
    m, k = 262_144, 256
    monkeypatch.delenv("MOR_B200_DEIM_BLOCK", raising=False)
    if kind == "gauss":
        B = DeviceMatrix.alloc(m, k).synth(_lib.SYNTH_GAUSS, seed=5, param=1.0 / np.sqrt(m))
        orthonormalize_dev(B)
    else:        # POD basis (this library, SVD(), all 256 modes) of Burgers snapshots on a 512 x 512 grid
        S = DeviceMatrix.alloc(m, k).synth(_lib.SYNTH_BURGERS, seed=5, param=2e-3, iparam=512)
        f = pod_factor_dev(S, _lib.MOR_SVD, k, flags=_lib.POD_MAY_OVERWRITE)
        B = f.basis(k, DeviceMatrix.alloc(m, k))
        f.free(); S.free()
    Bh = B.to_host()
    assert np.abs(Bh.T @ Bh - np.eye(k)).max() < 1e-10
    want, margins = O.deim_interpolation_indices(Bh, return_margins=True)
    want = want.tolist()
    got = deim_indices_dev(B).tolist()
    t = _lib.last_timings()
    assert t["deim_blocked"] == 1.0 or t["deim_fallback"] == 1.0
    assert got == want, "blocked: " + _first_difference_report(got, want, margins)
    # the library's own margins (from the blocked path's residuals) against the oracle's
    mg, mn, at = mor_b200.deim_last_margins()
    assert mg.shape == (k,) and 0.0 <= mn <= 1.0 and 2 <= at <= k
    if t["deim_fallback"] == 0.0:
        np.testing.assert_allclose(mg[1:], margins[1:], rtol=1e-6, atol=1e-9)
    got = mor_b200.deim_interpolation_indices(Bh).tolist()          # host pointers: pipelined ingest
    assert got == want, "host entry: " + _first_difference_report(got, want, margins)
    monkeypatch.setenv("MOR_B200_DEIM_BLOCK", "0")
    got = deim_indices_dev(B).tolist()
    assert _lib.last_timings()["deim_blocked"] == 0.0
    assert got == want, "streaming: " + _first_difference_report(got, want, margins)
    assert len(set(want)) == k


def test_margins_and_tie_fallback(gpu, monkeypatch):
    """mor_b200_deim_last_margins: the relative gap to the runner-up at every greedy step; a near-tie at a step >= 2 of
    the blocked path repeats the call in the reference's formulation (streaming path)."""
    monkeypatch.delenv("MOR_B200_DEIM_BLOCK", raising=False)
    monkeypatch.delenv("MOR_B200_DEIM_TIE_FALLBACK", raising=False)
    rng = np.random.Generator(np.random.PCG64(9))
    m, k = 40_000, 40
    Q = np.asfortranarray(np.linalg.qr(rng.standard_normal((m, k)))[0])
    want, margins = O.deim_interpolation_indices(Q, return_margins=True)
    assert mor_b200.deim_interpolation_indices(Q).tolist() == want.tolist()
    t = _lib.last_timings()
    assert t["deim_blocked"] == 1.0 and t["deim_fallback"] == 0.0
    mg, mn, at = mor_b200.deim_last_margins()
    np.testing.assert_allclose(mg, margins, rtol=1e-7, atol=1e-10)
    assert mn == pytest.approx(margins[1:].min(), rel=1e-7) and at == int(np.argmin(margins[1:])) + 2
    mg2, _, _, vals = mor_b200.deim_last_margins(with_values=True)
    assert np.array_equal(mg2, mg) and vals.shape == (k,) and np.all(vals > 0)
    assert vals[0] == np.abs(Q[:, 0]).max()
    # duplicated rows: every step is an exact tie (margin 0) -> fallback to the streaming path, same indices
    A = rng.standard_normal((m, k))
    A[30_000:] = A[:10_000]
    A = np.asfortranarray(A)
    idx = mor_b200.deim_interpolation_indices(A)
    t = _lib.last_timings()
    assert t["deim_fallback"] == 1.0 and t["deim_blocked"] == 0.0
    assert idx.tolist() == O.deim_interpolation_indices(A).tolist()
    assert mor_b200.deim_last_margins()[1] == 0.0
    monkeypatch.setenv("MOR_B200_DEIM_TIE_FALLBACK", "0")
    idx2 = mor_b200.deim_interpolation_indices(A)
    t = _lib.last_timings()
    assert t["deim_fallback"] == 0.0 and t["deim_blocked"] == 1.0
    assert idx2.tolist() == idx.tolist()                # exact ties resolve to the lowest index on both paths
    # streaming path reports margins too
    monkeypatch.setenv("MOR_B200_DEIM_BLOCK", "0")
    mor_b200.deim_interpolation_indices(Q)
    np.testing.assert_allclose(mor_b200.deim_last_margins()[0], margins, rtol=1e-7, atol=1e-10)
