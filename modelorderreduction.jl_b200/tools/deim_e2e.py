"""DEIM through the host-pointer entry point (pinned host basis), pipelined ingest (development aid)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "python"))
from mor_b200 import _lib  # noqa: E402
from mor_b200.device import DeviceMatrix, deim_indices_dev  # noqa: E402

lib = _lib.load()
_lib.check(lib.mor_b200_init(0))
m, k = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24, 256
B = DeviceMatrix.alloc(m, k).synth(_lib.SYNTH_GAUSS, seed=5, param=1.0 / np.sqrt(m))
ref = deim_indices_dev(B)
print("device-resident ms", _lib.last_timings()["total"], flush=True)
Bh = C.c_void_p()
_lib.check(lib.mor_b200_host_alloc(m * k * 8, C.byref(Bh)))
_lib.check(lib.mor_b200_mat_download(B.handle, Bh, m))
B.free()
idx = np.empty(k, dtype=np.int64)
for rep in range(3):
    t0 = time.perf_counter()
    _lib.check(lib.mor_b200_deim_indices(Bh, m, k, m, idx.ctypes.data))
    dt = (time.perf_counter() - t0) * 1e3
    t = _lib.last_timings()
    print(f"host entry rep={rep} wall={dt:.1f}ms h2d={t['h2d']:.1f}ms device total={t['total']:.1f}ms panel={t['deim_panel']:.1f} "
          f"step={t['deim_step']:.1f} tail={t['deim_tail']:.1f} same={bool(np.array_equal(idx, ref))} "
          f"PCIe {m*k*8/dt*1e-6:.1f} GB/s", flush=True)
lib.mor_b200_host_free(Bh)
