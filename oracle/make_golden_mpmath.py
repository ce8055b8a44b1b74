"""Independent pin of the oracle: 50-digit (mpmath) answers for the reference's own fixture, written to
tests/golden/mpmath_golden.json (run: python oracle/make_golden_mpmath.py).

TEST INFRASTRUCTURE ONLY (see oracle/mor_oracle.py).  Julia cannot run in the build image and the reference's
tests pin no numeric value on this path, so the oracle's own LAPACK-based goldens
(tests/golden/golden.json) would otherwise be pinned by nothing but themselves.  This script recomputes, with
NO LAPACK/BLAS/numpy arithmetic on the path (pure mpmath at 50 significant digits):

  F2  sin(i*j/n), 20 x 10                  /root/reference/src/precompile.jl:5-12, test/qa/jet_tests.jl:7-14
      * the 10 singular values (what LinearAlgebra.svd must return to ~eps*sigma_1, POD.jl:13,164)
      * renergy of reduce!(POD(S, 3), SVD()) and of the TSVD/RSVD formula (POD.jl:121-132, 198)
      * the left singular vectors, rounded to double
  DEIM chains, /root/reference/src/deim.jl:13-24, evaluated in 50-digit arithmetic ON THE DOUBLE-PRECISION
  BASIS GIVEN (the basis entries are exact binary rationals, so the chain is a pure function of them):
      per step l: c = (P'U)\\(P'u_l), r = |u_l - U c|, argmax (first maximal index), the winning residual and the
      runner-up residual -> the relative margin by which the argmax is decided.
      bases: Q[:, 1:5] of the fixture's Householder QR (the precompile workload's DEIM basis, precompile.jl:9-12),
      the 10 left singular vectors, and two seeded random orthonormal bases.
  A margin of 1e-3 .. 1e-1 against rounding errors of 1e-13 means ANY correct FP64 implementation -- Julia's
  dgetrf/dgemv chain, the oracle's, this library's streaming and blocked paths -- must return these indices.
"""
import json
import os
import struct
import sys

import mpmath as mp
import numpy as np
import scipy.linalg as sla

mp.mp.dps = 50
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "mpmath_golden.json")


def hexd(x: float) -> str:
    return struct.pack(">d", float(x)).hex()


def sin_fixture_mp():
    n, m = 20, 10
    return mp.matrix([[mp.sin(mp.mpf(i * j) / n) for j in range(1, m + 1)] for i in range(1, n + 1)])


def deim_chain_mp(B: np.ndarray):
    """deim.jl:13-24 in 50-digit arithmetic on the exact values of the double matrix B (m x k)."""
    m, k = B.shape
    U = mp.matrix(m, k)
    for i in range(m):
        for j in range(k):
            U[i, j] = mp.mpf(float(B[i, j]))
    idx, win, second = [], [], []

    def pick(r):
        order = sorted(range(m), key=lambda i: (-r[i], i))      # first maximal index
        return order[0], r[order[0]], r[order[1]] if m > 1 else mp.mpf(0)

    r = [abs(U[i, 0]) for i in range(m)]
    p, a, b = pick(r)
    idx.append(p); win.append(a); second.append(b)
    for l in range(1, k):
        PtU = mp.matrix(l, l)
        rhs = mp.matrix(l, 1)
        for a_ in range(l):
            for b_ in range(l):
                PtU[a_, b_] = U[idx[a_], b_]
            rhs[a_] = U[idx[a_], l]
        c = mp.lu_solve(PtU, rhs)
        r = []
        for i in range(m):
            s = mp.mpf(0)
            for j in range(l):
                s += U[i, j] * c[j]
            r.append(abs(U[i, l] - s))
        p, a, b = pick(r)
        idx.append(p); win.append(a); second.append(b)
    margins = [float((a - b) / a) if a != 0 else 0.0 for a, b in zip(win, second)]
    return {"indices": [i + 1 for i in idx], "winning_residual": [float(a) for a in win],
            "runner_up_residual": [float(b) for b in second], "relative_margin": margins,
            "min_relative_margin": min(margins)}


def main():
    g = {"digits": mp.mp.dps, "generator": "oracle/make_golden_mpmath.py (mpmath %s)" % mp.__version__}
    S = sin_fixture_mp()
    U, s, V = mp.svd_r(S, full_matrices=False, compute_uv=True)
    order = sorted(range(len(s)), key=lambda i: -s[i])
    sig = [s[i] for i in order]
    total = sum(sig)
    f2 = {"sigma_str": [mp.nstr(x, 40) for x in sig], "sigma": [float(x) for x in sig]}
    # reduce!(POD(S, 3), SVD()): determine_truncation with min = max = 3, min_renergy = 0 -> energy of the first 3 (plain sigma)
    f2["renergy_svd_k3"] = float(sum(sig[:3]) / total)
    # TSVD / RSVD formula, POD.jl:198,234: sum(s[1:k]) / (sum(s[1:k]) + (n_max - k) * s[k]), n_max = 10
    f2["renergy_tsvd_k3"] = float(sum(sig[:3]) / (sum(sig[:3]) + (10 - 3) * sig[2]))
    Ud = np.zeros((20, 10))
    for jn, j in enumerate(order):
        col = [U[i, j] for i in range(20)]
        big = max(range(20), key=lambda i: abs(col[i]))
        sgn = 1 if col[big] > 0 else -1                          # largest-|entry| positive (sign convention of the test)
        for i in range(20):
            Ud[i, jn] = float(sgn * col[i])
    f2["U_hex"] = [[hexd(v) for v in row] for row in Ud]
    f2["deim_U10"] = deim_chain_mp(Ud)
    # the precompile workload's DEIM basis: Q[:, 1:5] of qr(S) -- Householder QR in double (scipy = LAPACK dgeqrf,
    # what Julia's qr calls); the matrix is stored so the chain is pinned on exactly these doubles
    Sd = np.array([[np.sin(i * j / 20) for j in range(1, 11)] for i in range(1, 21)], order="F")
    Q, _ = sla.qr(Sd, mode="economic")
    Q5 = np.asfortranarray(Q[:, :5])
    f2["Q5_hex"] = [[hexd(v) for v in row] for row in Q5]
    f2["deim_Q5"] = deim_chain_mp(Q5)
    g["F2"] = f2
    cases = []
    for (m, k, seed) in [(64, 8, 1), (300, 16, 7)]:
        rng = np.random.Generator(np.random.PCG64(seed))
        Qr, _ = np.linalg.qr(rng.standard_normal((m, k)))
        Qr = np.asfortranarray(Qr)
        cases.append({"m": m, "k": k, "seed": seed, "basis_hex": [[hexd(v) for v in row] for row in Qr],
                      "chain": deim_chain_mp(Qr)})
    g["deim_random"] = cases
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", OUT)
    print("sigma", f2["sigma_str"])
    print("deim_U10", f2["deim_U10"]["indices"], "min margin %.3e" % f2["deim_U10"]["min_relative_margin"])
    print("deim_Q5 ", f2["deim_Q5"]["indices"], "min margin %.3e" % f2["deim_Q5"]["min_relative_margin"])
    for c in cases:
        print("random", c["m"], c["k"], c["chain"]["indices"], "min margin %.3e" % c["chain"]["min_relative_margin"])


if __name__ == "__main__":
    sys.exit(main())
