"""One symmetric Gram (through the POD factor entry) on a device-resident graded matrix, for kernel timelines."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "python"))
from mor_b200 import _lib  # noqa: E402
from mor_b200.device import DeviceMatrix, pod_factor_dev  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
_lib.check(_lib.load().mor_b200_init(0))
X = DeviceMatrix.alloc(m, n).synth(_lib.SYNTH_GRADED, seed=3, param=3.0)
for rep in range(2):
    f = pod_factor_dev(X, _lib.MOR_SVD, 64)
    print("gram ms", _lib.last_timings()["gram"], file=sys.stderr)
    f.free()
