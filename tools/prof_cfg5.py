"""BASELINE config #5 at full size: frustrated triangular L=48 TFIM with longitudinal field (J=+1, Gamma=1.0,
h=0.2, beta=32, 1024 replicas).  Thermalises, verifies every 128th replica with the reference invariant, times sweeps.
Usage: python tools/prof_cfg5.py [R] [therm] [sweeps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isingmontecarlo_b200 import MODE_FAST, lattices  # noqa: E402
from isingmontecarlo_b200.sse import QmcIsingGraph  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
therm = int(sys.argv[2]) if len(sys.argv) > 2 else 60
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L, beta = 48, 32.0
edges = lattices.triangular_periodic(L, 1.0)
g = QmcIsingGraph(edges, 1.0, 0.2, L * L, 0xC5000000 + np.arange(R, dtype=np.uint64), beta, mode=MODE_FAST)
t0 = time.perf_counter()
e = g.timesteps(therm, beta)
print(f"therm {therm} sweeps: {time.perf_counter() - t0:.2f} s, <n>={g.get_n().mean():.0f} <M>={g.get_cutoff().mean():.0f} cap={g.get_capacity()} "
      f"E/N={e.mean() / (L * L):.4f}")
assert all(g.verify(r) for r in range(0, R, 128))
for k in range(sweeps):
    v0 = g.total_vertex_updates()
    t0 = time.perf_counter()
    g.enqueue_sweeps(1)
    g.synchronize()
    dt = time.perf_counter() - t0
    print(f"sweep {k}: {dt * 1e3:.2f} ms, {(g.total_vertex_updates() - v0) / dt:.3e} vertex updates/s")
