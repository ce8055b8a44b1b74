"""Quick DEIM timing on device-resident synthetic bases (development aid)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "python"))
from mor_b200 import _lib  # noqa: E402
from mor_b200.device import DeviceMatrix, deim_indices_dev  # noqa: E402

_lib.check(_lib.load().mor_b200_init(0))
for (m, k) in [(1 << 20, 64), (1 << 22, 128), (1 << 24, 256)]:
    B = DeviceMatrix.alloc(m, k).synth(_lib.SYNTH_GAUSS, seed=5, param=1.0 / np.sqrt(m))
    ref = None
    for path in ("blocked", "streaming"):
        if path == "streaming":
            os.environ["MOR_B200_DEIM_BLOCK"] = "0"
        else:
            os.environ.pop("MOR_B200_DEIM_BLOCK", None)
        for rep in range(2):
            t0 = time.time()
            idx = deim_indices_dev(B)
            wall = time.time() - t0
            t = _lib.last_timings()
            bytes_alg = 4.0 * m * k * (k + 1)
            print(f"m={m} k={k} {path}({int(t['deim_blocked'])}) rep={rep} wall={wall*1e3:.1f}ms total={t['total']:.2f}ms "
                  f"panel={t['deim_panel']:.2f}ms step={t['deim_step']:.2f}ms tail={t['deim_tail']:.2f}ms  "
                  f"alg GB/s (total)={bytes_alg/t['total']*1e-6:.0f}  distinct={len(set(idx.tolist()))}", flush=True)
        if ref is None:
            ref = idx
        else:
            print("  same indices on both paths:", bool(np.array_equal(ref, idx)), flush=True)
    os.environ.pop("MOR_B200_DEIM_BLOCK", None)
    B.free()
