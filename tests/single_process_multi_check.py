"""Single-process multi-GPU mode (mor_b200_init_multi / MOR_B200_NGPUS) against the unsharded CPU oracle: ONE
process, the host-pointer entry points exactly as the Julia shim calls them, the library shards the rows over all
visible GPUs.  Also: pageable host memory (numpy arrays = what a Julia Matrix is) through the staging ring vs
page-locked buffers, wide / tiny inputs on GPU 0, the device API on GPU 0, shutdown and return to single-GPU mode.
Run:  python tests/single_process_multi_check.py [ngpus]      (not a pytest file; test_multi_gpu.py wraps it)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "modelorderreduction.jl_b200", "python"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def main():
    import torch
    import mor_b200
    import mor_oracle as O
    from mor_b200 import _lib
    from mor_b200 import dist as mdist
    from mor_b200.device import DeviceMatrix, deim_indices_dev

    n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    lib = _lib.load()
    if os.environ.get("MOR_B200_NGPUS"):
        # lazy start-up: the first library call of a process that never called mor_b200_init
        X0 = np.asfortranarray(np.random.default_rng(0).standard_normal((50, 4)))
        p0 = mor_b200.POD(X0, 2); mor_b200.reduce(p0, mor_b200.SVD())
        assert mdist.multi_info() == min(int(os.environ["MOR_B200_NGPUS"]), torch.cuda.device_count()), mdist.multi_info()
        n_gpus = mdist.multi_info()
    else:
        mdist.init_multi(n_gpus)
    assert mdist.multi_info() == n_gpus
    rng = np.random.default_rng(7)

    # ---- POD SVD(): sharded over the GPUs, pageable input -----------------------------------------
    m, n, k = 400_003, 96, 12
    X = np.asfortranarray(rng.standard_normal((m, n)) * np.logspace(0, -3, n))
    pod = mor_b200.POD(X, k); assert mor_b200.reduce(pod, mor_b200.SVD()) is None
    ref = O.POD(X, k); O.reduce(ref, O.SVD())
    assert O.sigma_error(ref.spectrum, pod.spectrum) <= 1e-10
    assert O.subspace_sine(ref.rbasis, pod.rbasis) <= 1e-8
    assert pod.rbasis.shape == (m, k) and pod.spectrum.shape == (n,)
    print(f"POD SVD() on {n_gpus} GPU(s) in one process: sigma err {O.sigma_error(ref.spectrum, pod.spectrum):.2e} ok", flush=True)
    # correlated columns: two passes, in place on the library's own copy
    Xc = np.asfortranarray(X @ (np.eye(n) + 0.5 * rng.standard_normal((n, n))))
    pc = mor_b200.POD(Xc, k); mor_b200.reduce(pc, mor_b200.SVD())
    rc = O.POD(Xc, k); O.reduce(rc, O.SVD())
    assert O.sigma_error(rc.spectrum, pc.spectrum) <= 1e-10 and O.subspace_sine(rc.rbasis, pc.rbasis) <= 1e-8
    # vector of vectors, TSVD, RSVD with a shared Omega
    vov = [np.ascontiguousarray(X[:, j]) for j in range(n)]
    pv = mor_b200.POD(vov, k); mor_b200.reduce(pv, mor_b200.TSVD())
    rv = O.POD(vov, k); O.reduce(rv, O.TSVD())
    assert O.sigma_error(rv.spectrum, pv.spectrum) <= 1e-10 and pv.renergy == rv.renergy or abs(pv.renergy - rv.renergy) < 1e-12
    omega = rng.standard_normal((n, k + 2))
    pr = mor_b200.POD(X, k); mor_b200.reduce(pr, mor_b200.RSVD(2), omega=omega)
    rr = O.POD(X, k); O.reduce(rr, O.RSVD(2), omega=omega)
    assert O.sigma_error(rr.spectrum, pr.spectrum) <= 1e-10 and O.subspace_sine(rr.rbasis, pr.rbasis) <= 1e-8
    print("VoV / TSVD / RSVD ok", flush=True)
    # page-locked input and output give the same bits as the staged (pageable) path
    Xp, nbytes = C.c_void_p(), m * n * 8
    _lib.check(lib.mor_b200_host_alloc(nbytes, C.byref(Xp)))
    Xpin = np.frombuffer((C.c_char * nbytes).from_address(Xp.value), dtype=np.float64).reshape((m, n), order="F")
    Xpin[:] = X
    pp = mor_b200.POD(Xpin, k); mor_b200.reduce(pp, mor_b200.SVD())
    assert np.array_equal(pp.spectrum, pod.spectrum) and np.array_equal(pp.rbasis, pod.rbasis)
    del Xpin
    lib.mor_b200_host_free(Xp)
    # wide and tiny inputs run on GPU 0
    Xw = np.asfortranarray(rng.standard_normal((32, 500)))
    pw = mor_b200.POD(Xw, 5); mor_b200.reduce(pw, mor_b200.TSVD())
    rw = O.POD(Xw, 5); O.reduce(rw, O.TSVD())
    assert O.sigma_error(rw.spectrum, pw.spectrum) <= 1e-10 and O.subspace_sine(rw.rbasis, pw.rbasis) <= 1e-8
    print("pinned == pageable, wide input ok", flush=True)

    # ---- DEIM: sharded, exact ties across shard boundaries, k = 256 -------------------------------------
    Q, _ = np.linalg.qr(rng.standard_normal((300_001, 40)))
    Q[250_000] = Q[7]
    Q = np.asfortranarray(Q)
    assert mor_b200.deim_interpolation_indices(Q).tolist() == O.deim_interpolation_indices(Q).tolist()
    Q2 = np.asfortranarray(np.linalg.qr(rng.standard_normal((131_072, 256)))[0])
    want, margins = O.deim_interpolation_indices(Q2, return_margins=True)
    got = mor_b200.deim_interpolation_indices(Q2)
    t = _lib.last_timings()
    assert got.tolist() == want.tolist()
    assert t["deim_exchange"] == (1.0 if n_gpus > 1 else 0.0), t["deim_exchange"]
    if t["deim_fallback"] == 0.0:
        np.testing.assert_allclose(mor_b200.deim_last_margins()[0][1:], margins[1:], rtol=1e-6, atol=1e-9)
    # projector and projections
    V = np.asfortranarray(np.linalg.qr(rng.standard_normal((131_072, 20)))[0])
    P = mor_b200.deim_projector(Q2[:, :64].copy(order="F"), V, O.deim_interpolation_indices(Q2[:, :64]))
    Pref = O.deim_projector(Q2[:, :64], V, O.deim_interpolation_indices(Q2[:, :64]))
    assert np.max(np.abs(P - Pref)) <= 1e-10 * np.max(np.abs(Pref))
    g = rng.standard_normal(131_072)
    assert np.max(np.abs(mor_b200.project(V, g) - V.T @ g)) <= 1e-12 * np.linalg.norm(g)
    # Galerkin projection V'AV of a sparse MOL operator: the COLUMNS of A are sharded, every GPU takes the rows of V its
    # columns touch from the host matrix (stencil halo), the k x k partials are all-reduced (deim.jl:154)
    import scipy.sparse as sp
    mg = 131_072
    A = (sp.diags([np.full(mg - 1, 1.0), np.full(mg, -2.0), np.full(mg - 1, 1.0), np.full(mg - 512, 0.5), np.full(mg - 512, 0.5)],
                  [-1, 0, 1, -512, 512], format="csc")
         + sp.coo_matrix((rng.standard_normal(2000), (rng.integers(0, mg, 2000), rng.integers(0, mg, 2000))), shape=(mg, mg)).tocsc())
    Ah = mor_b200.galerkin_project(A, V)
    ref = V.T @ (A @ V)
    assert np.max(np.abs(Ah - ref)) <= 1e-12 * max(np.max(np.abs(ref)), 1.0)
    assert _lib.last_timings()["spmm"] > 0
    print(f"DEIM (k=256, peer-mailbox exchange inside one process), projector, projections, sharded Galerkin product ok on {n_gpus} GPU(s)", flush=True)

    # ---- device API = GPU 0; shutdown; back to the one-GPU mode -----------------------------------------
    B = DeviceMatrix.alloc(50_000, 24).synth(_lib.SYNTH_GAUSS, seed=5, param=0.01)
    assert deim_indices_dev(B).tolist() == O.deim_interpolation_indices(B.to_host()).tolist()
    B.free()
    _lib.check(lib.mor_b200_shutdown())
    assert mdist.multi_info() == 0
    mdist.init(0)
    assert mor_b200.deim_interpolation_indices(Q).tolist() == O.deim_interpolation_indices(Q).tolist()
    print(f"single-process multi-GPU check ok on {n_gpus} GPU(s)", flush=True)


if __name__ == "__main__":
    main()
