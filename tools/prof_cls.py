"""Small driver for profiling the classical kernels: config #2, R replicas, a few fused launches.
Usage: python tools/prof_cls.py [R] [sweeps per launch] [launches]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isingmontecarlo_b200 import lattices  # noqa: E402
from isingmontecarlo_b200.classical import GraphState  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
spl = int(sys.argv[2]) if len(sys.argv) > 2 else 20
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L = 1024
g = GraphState(lattices.square_periodic(L, -1.0), np.zeros(L * L), 0xB2000000 + np.arange(R, dtype=np.uint64), 0.44068679350977147)
if os.environ.get("PROF_UNFUSED"):
    g.set_option("fused", 0)
g.sweeps(spl)
for k in range(launches):
    t0 = time.perf_counter()
    g.sweeps(spl)
    dt = time.perf_counter() - t0
    print(f"launch {k}: {dt * 1e3:.2f} ms, {R * L * L * spl / dt:.3e} flip attempts/s")
