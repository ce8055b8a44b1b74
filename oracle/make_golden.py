"""Regenerates tests/golden/*.json from the oracle (run: python oracle/make_golden.py).

The reference pins no numeric value on this path, so these vectors
are the oracle's own output on the reference's fixtures; the analytic entries (F1, F5) are
additionally asserted in tests/test_oracle.py.  Fixtures:
  F1  [3 0; 0 1], k=1                      test/interface.jl:4-5
  F2  sin(i*j/n), 20 x 10                  src/precompile.jl:5-12, test/qa/jet_tests.jl:7-14
  F5  determine_truncation quirk cases     src/DataReduction/POD.jl:121-132
  R*  seeded random cases (numpy PCG64) for DEIM / POD parity at sizes the oracle runs fast
"""
import json
import os
import sys

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mor_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def sin_fixture():
    n, m = 20, 10
    return np.array([[np.sin(i * j / n) for j in range(1, m + 1)] for i in range(1, n + 1)], order="F")


def main():
    os.makedirs(OUT, exist_ok=True)
    g = {}
    # F1
    X = np.array([[3.0, 0.0], [0.0, 1.0]], order="F")
    f1 = {}
    for name, alg in (("svd", O.SVD()), ("tsvd", O.TSVD()), ("rsvd", O.RSVD())):
        pod = O.POD(X, 1)
        O.reduce(pod, alg, omega=np.array([[1.0], [0.5]]))
        f1[name] = dict(spectrum=pod.spectrum.tolist(), renergy=pod.renergy, nmodes=pod.nmodes,
                        rbasis_abs=np.abs(pod.rbasis).tolist())
    g["F1"] = f1
    # F2
    S = sin_fixture()
    pod = O.POD(S, 3)
    O.reduce(pod, O.SVD())
    f2 = dict(spectrum=pod.spectrum.tolist(), renergy_svd_k3=pod.renergy)
    pod = O.POD(S, 3)
    O.reduce(pod, O.TSVD())
    f2["renergy_tsvd_k3"] = pod.renergy
    Q, _ = sla.qr(S, mode="economic")
    f2["deim_Q5"] = O.deim_interpolation_indices(Q[:, :5]).tolist()
    U, _, _ = O._svd(S)
    f2["deim_U10"] = O.deim_interpolation_indices(U[:, :10]).tolist()
    g["F2"] = f2
    # F5
    s = np.array([4.0, 3.0, 2.0, 1.0])
    g["F5"] = [dict(args=[1, 4, 0.5], out=list(O.determine_truncation(s, 1, 4, 0.5))),
               dict(args=[1, 2, 0.1], out=list(O.determine_truncation(s, 1, 2, 0.1))),
               dict(args=[2, 3, 0.95], out=list(O.determine_truncation(s, 2, 3, 0.95)))]
    # seeded random DEIM cases: (m, k, seed) -> indices
    cases = []
    for (m, k, seed) in [(64, 8, 1), (1000, 24, 2), (4099, 33, 3), (20000, 64, 4)]:
        rng = np.random.Generator(np.random.PCG64(seed))
        A = rng.standard_normal((m, k))
        Qr, _ = np.linalg.qr(A)
        cases.append(dict(m=m, k=k, seed=seed, indices=O.deim_interpolation_indices(Qr).tolist()))
    g["deim_random"] = cases
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", os.path.join(OUT, "golden.json"))
    print(json.dumps(g["F2"], indent=1))


if __name__ == "__main__":
    main()
