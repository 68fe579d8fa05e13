"""Driver for profiling the classical checkerboard kernel (config #2): python tools/prof_cls.py [R] [sweeps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isingmontecarlo_b200 import lattices  # noqa: E402
from isingmontecarlo_b200.classical import GraphState  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
L = int(os.environ.get("PROF_L", "1024"))
g = GraphState(lattices.square_periodic(L, -1.0), np.zeros(L * L), 0xB2000000 + np.arange(R, dtype=np.uint64), 0.44068679350977147)
g.do_time_step(20)
t0 = time.perf_counter()
g.do_time_step(sweeps)
dt = time.perf_counter() - t0
print(f"{sweeps} sweeps: {dt * 1e3:.2f} ms, {R * L * L * sweeps / dt:.3e} flips/s, e={g.get_energy().mean() / (L * L):.4f}")
