"""Soak run: many sweeps at full size through the fast kernels, then the reference invariant verify() on a sample of
replicas, the device status word, and the mean energy of the Metropolis and heat-bath rules against each other.
Usage: python tools/soak.py [R] [sweeps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isingmontecarlo_b200 import MODE_FAST, lattices  # noqa: E402
from isingmontecarlo_b200.sse import QmcIsingGraph  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
edges = lattices.square_periodic(32, -1.0)
res = {}
for hb in (False, True):
    g = QmcIsingGraph(edges, 3.04, 0.0, 1024, 0x50A70000 + 100000 * hb + np.arange(R, dtype=np.uint64), 16.0, mode=MODE_FAST)
    g.set_enable_heatbath(hb)
    t0 = time.perf_counter()
    g.timesteps(150, 16.0)
    e = g.timesteps(sweeps, 16.0)
    dt = time.perf_counter() - t0
    ok = all(g.verify(r) for r in range(0, R, max(1, R // 64)))
    res[hb] = (e.mean() / 1024, e.std(ddof=1) / np.sqrt(R) / 1024)
    print(f"heatbath={hb}: {150 + sweeps} sweeps of {R} replicas in {dt:.1f} s, verify={ok}, E/N = {res[hb][0]:.6f} +- {res[hb][1]:.6f}, "
          f"<n>={g.get_n().mean():.0f}, cap={g.get_capacity()}")
    assert ok
    g.close()
d = abs(res[False][0] - res[True][0])
err = np.hypot(res[False][1], res[True][1])
print(f"difference {d:.2e} = {d / err:.2f} sigma")
assert d < 4 * err
