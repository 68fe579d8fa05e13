"""Cost of the RVB update (sse_rvb.cu) next to the plain sweep: triangular lattice L x L, J = 1, Gamma = 1, R replicas.
usage: [PROF_H=h] python tools/prof_rvb.py [L] [beta] [R] [sweeps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from isingmontecarlo_b200 import MODE_COUNTER, lattices  # noqa: E402
from isingmontecarlo_b200.sse import QmcIsingGraph  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 12
beta = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
R = int(sys.argv[3]) if len(sys.argv) > 3 else 256
sweeps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
edges = lattices.triangular_periodic(L, 1.0)
h = float(os.environ.get("PROF_H", "0"))
g = QmcIsingGraph(edges, 1.0, h, L * L, 0xB0B0000 + np.arange(R, dtype=np.uint64), beta, mode=MODE_COUNTER)
g.timesteps(60, beta)
for rvb in (False, True):
    g.set_run_rvb(rvb)
    g.timesteps(2, beta)
    t0 = time.perf_counter()
    g.timesteps(sweeps, beta)
    dt = (time.perf_counter() - t0) / sweeps
    print(f"L={L} beta={beta} R={R} rvb={rvb}: {dt * 1e3:.2f} ms per sweep, mean n {g.get_n().mean():.0f}, mean cutoff {g.get_cutoff().mean():.0f}"
          + (f", success rate {g.rvb_success_rate().mean():.3f}" if rvb else ""), flush=True)
assert g.verify()
