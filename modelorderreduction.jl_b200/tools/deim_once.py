"""One DEIM run on a device-resident synthetic basis (profiling target for ncu)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "python"))
from mor_b200 import _lib  # noqa: E402
from mor_b200.device import DeviceMatrix, deim_indices_dev  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
k = int(sys.argv[2]) if len(sys.argv) > 2 else 256
_lib.check(_lib.load().mor_b200_init(0))
B = DeviceMatrix.alloc(m, k).synth(_lib.SYNTH_GAUSS, seed=5, param=1.0 / np.sqrt(m))
idx = deim_indices_dev(B)
t = _lib.last_timings()
print(f"m={m} k={k} total={t['total']:.2f} ms step={t['deim_step']:.2f} ms tail={t['deim_tail']:.2f} ms "
      f"alg GB/s={4.0 * m * k * (k + 1) / t['total'] * 1e-6:.0f} distinct={len(set(idx.tolist()))}")
