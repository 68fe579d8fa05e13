"""Soak run across the three modes at production shapes: config #3 with 4096 replicas (one wave, one warp per replica) and
with 1024 replicas (three warps per replica in COUNTER mode, two in FAST mode): verify() on a sample of replicas after many
sweeps and the mean energies of the modes against each other (they sample the same ensemble).
Usage: python tools/soak_modes.py [sweeps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isingmontecarlo_b200 import MODE_COUNTER, MODE_FAST, MODE_STRICT, lattices  # noqa: E402
from isingmontecarlo_b200.sse import QmcIsingGraph  # noqa: E402

sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
edges = lattices.square_periodic(32, -1.0)
res = {}
for name, mode, R, nsw in (("COUNTER 4096", MODE_COUNTER, 4096, sweeps), ("COUNTER 1024 (three warps per replica)", MODE_COUNTER, 1024, sweeps),
                           ("FAST 4096", MODE_FAST, 4096, sweeps), ("STRICT 4096", MODE_STRICT, 4096, max(40, sweeps // 5))):
    g = QmcIsingGraph(edges, 3.04, 0.0, 1024, 0x50A80000 + 7919 * len(res) + np.arange(R, dtype=np.uint64), 16.0, mode=MODE_COUNTER)
    g.timesteps(150, 16.0)  # thermalise in the fastest mode
    g.set_mode(mode)
    t0 = time.perf_counter()
    g.timesteps(20, 16.0)
    e = g.timesteps(nsw, 16.0)
    dt = time.perf_counter() - t0
    ok = all(g.verify(r) for r in range(0, R, max(1, R // 48)))
    res[name] = (e.mean() / 1024, e.std(ddof=1) / np.sqrt(R) / 1024)
    print(f"{name}: {20 + nsw} sweeps of {R} replicas in {dt:.1f} s, verify={ok}, E/N = {res[name][0]:.6f} +- {res[name][1]:.6f}, <n>={g.get_n().mean():.0f}")
    assert ok
    g.close()
names = list(res)
for a in names[1:]:
    d = abs(res[names[0]][0] - res[a][0])
    err = np.hypot(res[names[0]][1], res[a][1])
    print(f"{names[0]} vs {a}: difference {d:.2e} = {d / err:.2f} sigma")
    assert d < 4 * err
