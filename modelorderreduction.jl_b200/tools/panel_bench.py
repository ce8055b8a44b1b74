"""Blocked DEIM at 16M x 256 for every tile shape of deim_panel_kernel (development aid)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "python"))
from mor_b200 import _lib  # noqa: E402
from mor_b200.device import DeviceMatrix, deim_indices_dev  # noqa: E402

_lib.check(_lib.load().mor_b200_init(0))
m, k = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24, 256
B = DeviceMatrix.alloc(m, k).synth(_lib.SYNTH_GAUSS, seed=5, param=1.0 / np.sqrt(m))
ref = None
for tile in sys.argv[2:] or ["128", "256", "512"]:
    os.environ["MOR_B200_PANEL_TILE"] = tile
    for rep in range(2):
        idx = deim_indices_dev(B)
        t = _lib.last_timings()
        print(f"tile={tile} rep={rep} total={t['total']:.2f}ms panel={t['deim_panel']:.2f}ms step={t['deim_step']:.2f}ms "
              f"tail={t['deim_tail']:.2f}ms", flush=True)
    if ref is None:
        ref = idx
    else:
        print("  same indices:", bool(np.array_equal(ref, idx)), flush=True)
