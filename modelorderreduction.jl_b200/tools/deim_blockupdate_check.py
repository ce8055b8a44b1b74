"""One-shot check of the EXPERIMENTAL per-panel update of the k x k DEIM factors (MOR_B200_DEIM_BLOCKUPDATE=1):
indices against the default path on a few shapes, timing at 2M x 256 (one GPU's share of config 5 at N = 8)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "python"))
from mor_b200 import _lib  # noqa: E402
from mor_b200.device import DeviceMatrix, deim_indices_dev  # noqa: E402

_lib.check(_lib.load().mor_b200_init(0))
ok = True
for (m, k) in [(70_001, 40), (20_000, 130), (300_000, 17), (1 << 21, 256)]:
    B = DeviceMatrix.alloc(m, k).synth(_lib.SYNTH_GAUSS, seed=5, param=1.0 / np.sqrt(m))
    os.environ.pop("MOR_B200_DEIM_BLOCKUPDATE", None)
    ref = deim_indices_dev(B); ref = deim_indices_dev(B)
    t0 = _lib.last_timings()
    os.environ["MOR_B200_DEIM_BLOCKUPDATE"] = "1"
    try:
        got = deim_indices_dev(B); got = deim_indices_dev(B)
        t1 = _lib.last_timings()
        same = bool(np.array_equal(ref, got))
    except Exception as ex:  # noqa: BLE001
        same, t1 = False, {"total": float("nan"), "deim_tail": float("nan"), "deim_panel": float("nan")}
        print("  exception:", repr(ex))
    ok = ok and same
    print(f"m={m} k={k} same={same} default total={t0['total']:.2f}ms tail={t0['deim_tail']:.2f} panel={t0['deim_panel']:.2f} | "
          f"blockupdate total={t1['total']:.2f}ms tail={t1['deim_tail']:.2f} panel={t1['deim_panel']:.2f}", flush=True)
    B.free()
print("BLOCKUPDATE", "OK" if ok else "MISMATCH")
