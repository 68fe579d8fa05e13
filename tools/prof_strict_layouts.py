"""STRICT sweep of config #3 on both workspace layouts, same thermalised batch.  Usage: python tools/prof_strict_layouts.py [R] [therm] [sweeps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isingmontecarlo_b200 import MODE_COUNTER, MODE_STRICT, lattices  # noqa: E402
from isingmontecarlo_b200.sse import QmcIsingGraph  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
therm = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L, beta = 32, 16.0
g = QmcIsingGraph(lattices.square_periodic(L, -1.0), 3.04, 0.0, L * L, 0x55E00000 + np.arange(R, dtype=np.uint64), beta, mode=MODE_COUNTER)
g.timesteps(therm, beta)
print(f"<n>={g.get_n().mean():.0f} <M>={g.get_cutoff().mean():.0f}")
g.set_mode(MODE_STRICT)
DBG = bool(os.environ.get("PROF_DBG"))
if DBG:
    g.set_option("debug_counters", 1)
last = g.debug_counters().astype(np.float64) if DBG else None
plan = [(int(x), 0) for x in os.environ.get("PROF_LAYOUTS", "1,5,0,1,5").split(",")]
for layout, gran in plan:
    g.set_option("strict_layout", layout)
    if gran:
        g.set_option("l2_fetch_granularity", gran)
    print("l2 fetch granularity", gran or "default")
    for k in range(sweeps):
        v0 = g.total_vertex_updates()
        t0 = time.perf_counter()
        g.enqueue_sweeps(1)
        g.synchronize()
        dt = time.perf_counter() - t0
        print(f"layout {layout} sweep {k}: {dt * 1e3:.2f} ms, {(g.total_vertex_updates() - v0) / dt:.3e} vertex updates/s")
    if DBG:
        c = g.debug_counters().astype(np.float64)
        d = (c - last) / R / sweeps
        last = c
        names = ["links", "label walk", "flip bits", "apply", "free spins"]
        print("   phase Mcycles per replica-sweep: " + ", ".join(f"{n} {d[48 + i] / 1e6:.2f}" for i, n in enumerate(names)))
        cn = ["interior pops", "site arrivals", "interior ops labelled", "wraps", "frontier pops"]
        print("   events per replica-sweep: " + ", ".join(f"{n} {d[56 + i]:.0f}" for i, n in enumerate(cn)))
assert g.verify(0) if hasattr(g, "verify") else True
