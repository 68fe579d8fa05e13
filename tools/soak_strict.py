"""STRICT (reference cluster order) against FAST at config #3: thermalise with the fast kernels, then measure the energy
over the same number of sweeps in each order from independent streams.  Usage: python tools/soak_strict.py [R] [sweeps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isingmontecarlo_b200 import MODE_FAST, MODE_STRICT, lattices  # noqa: E402
from isingmontecarlo_b200.sse import QmcIsingGraph  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
edges = lattices.square_periodic(32, -1.0)
res = {}
for mode, name in ((MODE_FAST, "FAST"), (MODE_STRICT, "STRICT")):
    g = QmcIsingGraph(edges, 3.04, 0.0, 1024, 0x57A10000 + 100000 * mode + np.arange(R, dtype=np.uint64), 16.0, mode=MODE_FAST)
    g.timesteps(150, 16.0)
    g.set_mode(mode)
    t0 = time.perf_counter()
    e = g.timesteps(sweeps, 16.0)
    dt = time.perf_counter() - t0
    ok = all(g.verify(r) for r in range(0, R, max(1, R // 32)))
    res[name] = (e.mean() / 1024, e.std(ddof=1) / np.sqrt(R) / 1024)
    print(f"{name}: {sweeps} sweeps of {R} replicas in {dt:.1f} s, verify={ok}, E/N = {res[name][0]:.6f} +- {res[name][1]:.6f}")
    assert ok
    g.close()
d = abs(res["FAST"][0] - res["STRICT"][0])
err = np.hypot(res["FAST"][1], res["STRICT"][1])
print(f"difference {d:.2e} = {d / err:.2f} sigma")
assert d < 4 * err
