"""Small driver for profiling the SSE kernels: thermalise R replicas of config #3, then run a few
timed sweeps.  Usage: python tools/prof_sse.py [R] [therm] [sweeps] [mode]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isingmontecarlo_b200 import MODE_COUNTER, MODE_FAST, MODE_STRICT, lattices  # noqa: E402
from isingmontecarlo_b200.sse import QmcIsingGraph  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
therm = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = {"strict": MODE_STRICT, "counter": MODE_COUNTER}.get(sys.argv[4] if len(sys.argv) > 4 else "", MODE_FAST)
L = int(os.environ.get("PROF_L", "32"))
beta = float(os.environ.get("PROF_BETA", "16"))
edges = lattices.square_periodic(L, -1.0)
g = QmcIsingGraph(edges, 3.04, 0.0, L * L, 0x55E00000 + np.arange(R, dtype=np.uint64), beta, mode=MODE_COUNTER if mode == MODE_COUNTER else MODE_FAST)
t0 = time.perf_counter()
g.timesteps(therm, beta)
print(f"therm {therm} sweeps: {time.perf_counter() - t0:.3f} s, <n>={g.get_n().mean():.0f} <M>={g.get_cutoff().mean():.0f}")
g.set_mode(mode)
if os.environ.get('PROF_MINB'):
    g.set_option('minblocks', int(os.environ['PROF_MINB']))
for knob in ('smem_pad', 'smem_carveout', 'pipeline'):
    if os.environ.get('PROF_' + knob.upper()):
        g.set_option(knob, int(os.environ['PROF_' + knob.upper()]))
if os.environ.get('PROF_EPK'):
    g.set_option('shared_edge_table', int(os.environ['PROF_EPK']))
if os.environ.get('PROF_DBG'):
    g.set_option('debug_counters', 1)
for k in range(sweeps):
    v0 = g.total_vertex_updates()
    t0 = time.perf_counter()
    g.enqueue_sweeps(1)
    g.synchronize()
    dt = time.perf_counter() - t0
    print(f"sweep {k}: {dt * 1e3:.2f} ms, {(g.total_vertex_updates() - v0) / dt:.3e} vertex updates/s")

if os.environ.get('PROF_DBG'):
    c = g.debug_counters()
    print('dbg: chunks', c[0], 'fast', c[1], 'fast_fail_first', c[2], 'slow', c[3], 'why[delta|empty<<1|amb<<2]', list(c[4:12]))
    names = ["between steps", "load+classify", "philox window", "thresholds+decode", "table+walk", "decisions+commit", "store+flips",
             "rest of segment step", "closure", "P2", "P3", "final-op decode", "tb/rep/match", "unions"]
    t = c[32:32 + len(names)].astype(float)
    if t.sum() == 0:
        print('(build with EXTRA=-DQMCB_PHASE_TIMERS to get phase timers)')
        t[:] = 1
    print("phase cycles per replica-sweep (lane 0 clock64, summed over replicas / R / sweeps):")
    for nm, v in zip(names, t):
        print(f"  {nm:22s} {v / R / sweeps / 1e6:8.2f} Mcycles  {100 * v / t.sum():5.1f}%")
if os.environ.get('PROF_DBG') and mode == MODE_COUNTER:
    c = g.debug_counters().astype(np.float64)
    names = ["set-up", "P1 loop", "closure", "P2", "P3", "tail"]
    for role in range(3):
        v = c[32 + 8 * role:32 + 8 * role + 6] / R / sweeps
        if v.sum() > 0:
            print(f"role {'ABC'[role]} Mcycles per replica-sweep: " + ", ".join(f"{n} {x / 1e6:.2f}" for n, x in zip(names, v))
                  + f"; waiting at the step barrier {c[56 + role] / R / sweeps / 1e6:.2f}")
