"""Config 1 of BASELINE.json: the README's FitzHugh-Nagumo example (README.md:41-61,
docs/src/tutorials/deim_FitzHugh-Nagumo.md:50-85) discretised by hand with the method of lines
(N = 15 intervals -> 16 nodes x {v, w} = 32 states, second-order centred differences, Neumann
boundary conditions through ghost nodes) and integrated with scipy.  MethodOfLines.jl/Tsit5 would
give a slightly different trajectory; only its *shape* (32 x nt, wide) and character matter here:
the fixture feeds both the oracle and the library, which must agree with each other."""
import numpy as np
from scipy.integrate import solve_ivp

L, EPS, B, GAMMA, C = 1.0, 0.015, 0.5, 2.0, 0.05
N = 15
DX = L / N


def f_nl(v):
    return v * (v - 0.1) * (1.0 - v)


def i0(t):
    return 50000.0 * t**3 * np.exp(-15.0 * t)


def rhs(t, y):
    v, w = y[:N + 1], y[N + 1:]
    vl = v[1] + 2 * DX * i0(t)           # Dx(v(0,t)) = -i0(t)  -> ghost node on the left
    vr = v[-2]                           # Dx(v(L,t)) = 0
    vpad = np.concatenate(([vl], v, [vr]))
    vxx = (vpad[2:] - 2 * vpad[1:-1] + vpad[:-2]) / DX**2
    dv = (EPS**2 * vxx + f_nl(v) - w + C) / EPS
    dw = B * v - GAMMA * w + C
    return np.concatenate((dv, dw))


def snapshots(nt=1401):
    """32 x nt snapshot matrix (column = state at one time), t in [0, 14]."""
    sol = solve_ivp(rhs, (0.0, 14.0), np.zeros(2 * (N + 1)), method="LSODA", rtol=1e-9, atol=1e-11,
                    t_eval=np.linspace(0.0, 14.0, nt))
    assert sol.success
    return np.asfortranarray(sol.y)


def nonlinear_snapshots(X):
    """What deim(sys, ...) builds at deim.jl:239-243: the nonlinear term evaluated on every snapshot
    (here f(v)/eps on the v block; the w block has no nonlinearity)."""
    F = np.zeros_like(X)
    F[:N + 1] = f_nl(X[:N + 1]) / EPS
    return np.asfortranarray(F)
