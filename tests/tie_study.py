"""How often do the library's two DEIM paths and the oracle part ways on smooth PDE data, and by what margin?
(VERDICT r1 weak #3.)  Not a pytest file: run on the GPU box,
    python tests/tie_study.py > gpurun_out/tie_study.json
The first Burgers generator of round 1 used few-digit DECIMAL front directions; on a grid they produce distant
rows whose histories agree to the last bits, i.e. argmax decisions at rounding level.  For each case this script
builds the snapshots on the CPU (oracle restatement of the closed form), takes the POD basis from the library,
and compares three evaluations of deim.jl:9-28 on that same basis: the oracle (dgetrf/dgemv), the library's
streaming path and the library's blocked path.  For every disagreement it reports the step, both rows, the
oracle's relative margin at that step, and the library's own reported margin (mor_b200_deim_last_margin)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "modelorderreduction.jl_b200", "python"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import mor_b200  # noqa: E402
import mor_oracle as O  # noqa: E402
from mor_b200 import _lib  # noqa: E402


def lib_indices(B, blocked):
    if blocked:
        os.environ.pop("MOR_B200_DEIM_BLOCK", None)
    else:
        os.environ["MOR_B200_DEIM_BLOCK"] = "0"
    os.environ["MOR_B200_DEIM_TIE_FALLBACK"] = "0"          # study the raw paths
    try:
        idx = mor_b200.deim_interpolation_indices(B)
        margin = mor_b200.deim_last_margin() if hasattr(mor_b200, "deim_last_margin") else None
    finally:
        os.environ.pop("MOR_B200_DEIM_BLOCK", None)
        os.environ.pop("MOR_B200_DEIM_TIE_FALLBACK", None)
    return idx, margin


def first_diff(a, b):
    d = np.flatnonzero(np.asarray(a) != np.asarray(b))
    return None if d.size == 0 else int(d[0])


def main():
    _lib.check(_lib.load().mor_b200_init(0))
    cases = []
    for decimal in (True, False):
        for (nx, n, nu, seed) in [(256, 64, 1e-3, 1), (256, 64, 2e-3, 2), (256, 64, 5e-4, 3), (256, 64, 1e-3, 4),
                                  (256, 64, 3e-3, 5), (256, 64, 1.5e-3, 6), (512, 128, 1e-3, 7), (512, 128, 5e-4, 8),
                                  (1024, 128, 5e-4, 9), (1024, 160, 2.5e-4, 10)]:
            m = nx * nx
            S = O.burgers_snapshots(m, n, nx, nu, seed=seed, decimal=decimal)
            pod = mor_b200.POD(S, n)
            mor_b200.reduce(pod, mor_b200.SVD())
            B = np.asfortranarray(pod.rbasis)
            want, margins = O.deim_interpolation_indices(B, return_margins=True)
            got_b, mb = lib_indices(B, True)
            got_s, ms = lib_indices(B, False)
            case = {"directions": "decimal" if decimal else "irrational", "grid": nx, "snapshots": n, "nu": nu, "seed": seed,
                    "oracle_min_margin": float(margins.min()), "oracle_min_margin_step": int(np.argmin(margins)) + 1,
                    # step 1 is argmax |u_1| (no arithmetic, identical in every implementation): excluded below
                    "oracle_min_margin_steps_ge_2": float(margins[1:].min()),
                    "oracle_min_margin_step_ge_2": int(np.argmin(margins[1:])) + 2,
                    "steps_ge_2_with_margin_below_1e-9": int(np.sum(margins[1:] < 1e-9)),
                    "library_margin_blocked": mb, "library_margin_streaming": ms}
            for name, got in (("blocked", got_b), ("streaming", got_s)):
                t = first_diff(got, want)
                case[name] = {"equal_to_oracle": t is None}
                if t is not None:
                    case[name].update({"first_difference_step": t + 1, "library_row": int(got[t]), "oracle_row": int(want[t]),
                                       "oracle_margin_at_step": float(margins[t]),
                                       "positions_differing": int(np.sum(np.asarray(got) != np.asarray(want)))})
            t = first_diff(got_b, got_s)
            case["blocked_equals_streaming"] = t is None
            cases.append(case)
    summary = {}
    for d in ("decimal", "irrational"):
        cs = [c for c in cases if c["directions"] == d]
        summary[d] = {"cases": len(cs), "blocked_differs_from_oracle": sum(not c["blocked"]["equal_to_oracle"] for c in cs),
                      "streaming_differs_from_oracle": sum(not c["streaming"]["equal_to_oracle"] for c in cs),
                      "largest_oracle_margin_at_a_difference": max(
                          [c[p]["oracle_margin_at_step"] for c in cs for p in ("blocked", "streaming")
                           if not c[p]["equal_to_oracle"]] or [0.0])}
    print(json.dumps({"summary": summary, "cases": cases}, indent=1))


if __name__ == "__main__":
    main()
