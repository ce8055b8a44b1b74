"""bench.py's own checkers, exercised on the CPU against the oracle (the bench line's `deim.independent_check` and
the reference arm's bookkeeping must themselves be right before they are allowed to certify anything)."""
import json
import os
import subprocess
import sys

import numpy as np
import torch

import mor_oracle as O
from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_independent_deim_check_agrees_with_the_oracle():
    rng = np.random.default_rng(0)
    Q, _ = np.linalg.qr(rng.standard_normal((6000, 48)))
    idx, margins = O.deim_interpolation_indices(Q, return_margins=True)
    bt = torch.from_numpy(np.ascontiguousarray(Q.T))          # (k, ld) view: row j = basis column j
    out = bench.deim_independent_check(torch, None, bt, 6000, 0, idx, 1, list(range(48)), lib_values=None)
    assert out["all_argmax_match"] and out["mismatches"] == []
    assert abs(out["min_relative_margin_steps_ge_2"] / float(margins[1:].min()) - 1.0) < 1e-6   # two fp64 evaluations of one margin
    assert out["min_margin_step_ge_2"] == int(np.argmin(margins[1:])) + 2
    wrong = idx.copy()
    wrong[9] = wrong[9] % 6000 + 1                              # a neighbouring row: not the argmax
    bad = bench.deim_independent_check(torch, None, bt, 6000, 0, wrong, 1, list(range(12)))
    assert not bad["all_argmax_match"] and bad["mismatches"][0]["step"] == 10
    assert bad["mismatches"][0]["independent_argmax_row"] == int(idx[9])
    # with the library's winning residuals supplied, the rounding-level discrepancy is reported
    r = np.zeros(48)
    r[0] = np.abs(Q[:, 0]).max()
    for l in range(1, 48):
        c = np.linalg.solve(Q[idx[:l] - 1, :l], Q[idx[:l] - 1, l])
        r[l] = np.abs(Q[:, l] - Q[:, :l] @ c).max()
    out = bench.deim_independent_check(torch, None, bt, 6000, 0, idx, 1, list(range(48)), lib_values=r)
    assert out["residual_discrepancy"]["max_rel"] < 1e-10


def test_reference_arm_line_is_labelled_as_a_sample():
    env = dict(os.environ, OMP_NUM_THREADS="1")               # what torch.distributed.run exports to its workers
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-rows", "2048", "--cpu-rows-big", "4096"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port"
    assert line["config"]["sample_rows"] == 2048 and line["config"]["extrapolated"] is True
    assert line["cpu_baseline"]["linearity"]["rows"] == [2048, 4096]
    cores = len(os.sched_getaffinity(0))
    assert line["cpu_baseline"]["cores"] == cores             # all host cores despite OMP_NUM_THREADS=1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] == line["cpu_baseline"]["value"]


def test_pod_workload_law_of_the_cpu_sample():
    X = bench.cpu_pod_matrix(4096)
    assert X.shape == (4096, bench.N_COLS) and X.flags.f_contiguous
    d = np.sqrt((X * X).mean(axis=0))
    want = 10.0 ** (-3.0 * np.arange(bench.N_COLS) / (bench.N_COLS - 1))
    assert np.max(np.abs(d / want - 1.0)) < 0.1                 # column scale 10^(-3j/(n-1)), N(0,1) entries
